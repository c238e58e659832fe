#!/bin/bash
# ncu source-level capture of the persistent residual-VQ kernel at the stage-2 shape (run under gpurun)
set -e
mkdir -p gpurun_out/$1
python profiles/prof_rvq_trace.py > gpurun_out/$1/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:rvq_fused -s 3 -c 1 -f -o gpurun_out/$1/prof_rvq python profiles/prof_rvq_trace.py > gpurun_out/$1/ncu.log 2>&1
