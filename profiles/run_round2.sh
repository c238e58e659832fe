#!/bin/bash
# Round-2 evidence pass on one B200 (under gpurun): every bench workload, launch lists of the default bench (c3) and of
# the stage-2 workloads, one full ncu capture of the persistent residual-VQ kernel.  Numbers printed under ncu are never
# bench values.   bash profiles/run_round2.sh <out-subdir>
set -u
O=gpurun_out/${1:-r02}
mkdir -p $O
python bench.py --workload c2 --steps 50 --warmup 10 > $O/bench_c2.json 2> $O/bench_c2.err
python bench.py --workload rvq --steps 50 --warmup 10 > $O/bench_rvq.json 2> $O/bench_rvq.err
python bench.py --workload rvq --steps 50 --warmup 10 --graph > $O/bench_rvq_graph.json 2>> $O/bench_rvq.err
python bench.py --workload c4 --steps 3 --warmup 3 > $O/bench_c4.json 2> $O/bench_c4.err
python bench.py --workload c3 --mode bf16_input --steps 5 --warmup 3 > $O/bench_c3_bf16.json 2> $O/bench_c3_bf16.err
python bench.py --steps 2 --warmup 3 > $O/plain_c3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c3.csv python bench.py --steps 2 --warmup 3 > $O/ncu_c3.log 2>&1
python profiles/prof_rvq_trace.py > $O/trace_rvq.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:rvq_fused -s 3 -c 1 -f -o $O/prof_rvq_fused python profiles/prof_rvq_trace.py > $O/ncu_rvq_full.log 2>&1
for w in c2 rvq rvq_graph c4 c3_bf16; do python - <<PY
import json
try:
    d = json.loads(open("$O/bench_$w.json").read().strip().splitlines()[-1])
    print("$w", round(d["ms_per_step"], 4), "ms", round(d["value"] / 1e6, 2), "M/s e2e", round(d["e2e"]["value"] / 1e6, 2), "M/s",
          d["roofline"]["bound"], round(d["roofline"]["frac"], 3), "step", round(d["roofline"]["step_frac"], 3), "launches", d["gpu_launches"])
except Exception as e:
    print("$w failed", e)
PY
done
