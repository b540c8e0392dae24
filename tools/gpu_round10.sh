#!/bin/bash
mkdir -p gpurun_out
run() {
env $1 timeout 600 python bench.py --steps 3 --warmup 3 --no-baselines > gpurun_out/bench_r10.json 2> gpurun_out/bench_r10.err; echo "bench [$1] rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_r10.json').read().strip().splitlines()[-1])
    print(round(d["value"],1), round(d["ms_per_frame"],4), "e2e", round(d["e2e"]["value"],1), d["sections_us"])
except Exception as e: print("parse failed", e); print(open('gpurun_out/bench_r10.err').read()[-2500:])
PY
}
run "MEMOTR_PIPE_SPLIT=3.5"
run "MEMOTR_PIPE_SPLIT=4"
run "MEMOTR_PIPE_SPLIT=4.5"
run "MEMOTR_PIPE_SPLIT=4.5 MEMOTR_PIPE_RESERVE=28"
