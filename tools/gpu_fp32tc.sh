#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q --tb=short -x -s -k "f32x3 or fp32tc" > gpurun_out/pytest_tc3.log 2>&1
echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_tc3.log | tail -3
for m in fp32tc fp32; do
timeout 900 python bench.py --mode $m --steps 2 --warmup 3 --clip-frames 16 --no-baselines > gpurun_out/bench_$m.json 2> gpurun_out/bench_$m.err; echo "bench $m rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_$m.json').read().strip().splitlines()[-1])
    print("$m", round(d["value"],1), round(d["ms_per_frame"],4), "e2e", round(d["e2e"]["value"],1), d["sections_us"])
except Exception as e: print("parse failed", e); print(open('gpurun_out/bench_$m.err').read()[-1500:])
PY
done
