#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_tracker_gpu.py -m gpu -q --tb=short -x -k "pos_embed or position_maps or pipelined or bf16_clip or reference_init" > gpurun_out/pytest_r11.log 2>&1
echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_r11.log | tail -8
for i in 1 2; do
timeout 600 python bench.py --steps 3 --warmup 3 --no-baselines > gpurun_out/bench_r11.json 2> gpurun_out/bench_r11.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_r11.json').read().strip().splitlines()[-1])
    print(round(d["value"],1), round(d["ms_per_frame"],4), "e2e", round(d["e2e"]["value"],1), d["sections_us"])
except Exception as e: print("parse failed", e); print(open('gpurun_out/bench_r11.err').read()[-2500:])
PY
done
