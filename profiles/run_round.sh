#!/bin/bash
# One GPU-box pass: tests, smoke, every bench workload, launch lists, full ncu captures of the two hot kernels.
# Usage (from the repo root, under gpurun):  bash profiles/run_round.sh
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -2 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
python bench.py > $O/bench_c2.json 2> $O/bench_c2.err
python bench.py --workload c3 --steps 5 --warmup 3 > $O/bench_c3.json 2> $O/bench_c3.err
python bench.py --workload c3 --mode bf16_input --steps 5 --warmup 3 > $O/bench_c3_bf16.json 2> $O/bench_c3_bf16.err
python bench.py --workload rvq > $O/bench_rvq.json 2> $O/bench_rvq.err
python bench.py --workload rvq --graph > $O/bench_rvq_graph.json 2> $O/bench_rvq_graph.err
python bench.py --workload c5 > $O/bench_c5.json 2> $O/bench_c5.err
python bench.py --workload c4 --steps 3 --warmup 3 > $O/bench_c4.json 2> $O/bench_c4.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref_c2.json 2> $O/bench_ref_c2.err
for w in c2 c3 rvq c5 c4; do python - <<PY
import json
try:
    d = json.load(open("$O/bench_$w.json"))
    print("$w", round(d["ms_per_step"], 4), "ms", round(d["value"] / 1e6, 2), "M/s e2e", round(d["e2e"]["value"] / 1e6, 2), "M/s",
          d["roofline"]["bound"], round(d["roofline"]["frac"], 3), "launches", d["gpu_launches"])
except Exception as e:
    print("$w failed", e, open("$O/bench_$w.err").read()[-600:])
PY
done
# launch lists (per-launch durations; numbers printed under ncu are never bench values)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c2.csv python bench.py --steps 3 --warmup 3 > $O/ncu_c2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c3.csv python bench.py --workload c3 --steps 2 --warmup 3 > $O/ncu_c3.log 2>&1
# full captures of the dominant kernels
ncu --set full --clock-control none --import-source on -k regex:quantize_fused_kernel -s 2 -c 1 -f -o $O/prof_c2_fused_r2 python profiles/prof_forward.py 512 64 1048576 fp32 4 > $O/ncu_full_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:search_tc2_kernel -s 1 -c 1 -f -o $O/prof_c3_tc2_r2 python profiles/prof_search.py 8192 256 1048576 fp32 3 > $O/ncu_full_c3.log 2>&1
ls -la $O/*.ncu-rep | tail -3
