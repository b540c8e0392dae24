#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tracker_gpu.py -m gpu -q --tb=short -x -k "pipelined" > gpurun_out/pytest_r8.log 2>&1
echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|Error|assert" gpurun_out/pytest_r8.log | tail -15
for flag in ""; do
timeout 600 python bench.py --steps 3 --warmup 3 --no-baselines $flag > gpurun_out/bench_r8.json 2> gpurun_out/bench_r8.err; echo "bench [$flag] rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_r8.json').read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_frame","sections_us")}, d["e2e"]["value"], d["roofline"]["duration_us"], d.get("exact_two_phase",{}).get("value"))
except Exception as e: print("parse failed", e); print(open('gpurun_out/bench_r8.err').read()[-2500:])
PY
done
