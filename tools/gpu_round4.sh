#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/micro_dense.py 2>&1 | tail -2
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q --tb=short -x -k "linear256 or mlp2 or dense_block or reference_init" > gpurun_out/pytest_r4.log 2>&1
echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_r4.log | tail -15
for v in 1 0; do
MEMOTR_FUSE_OUTLN=$v timeout 600 python bench.py --steps 3 --warmup 3 --no-baselines > gpurun_out/bench_ol$v.json 2> gpurun_out/bench_ol$v.err; echo "bench outln=$v rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_ol$v.json').read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_frame","sections_us")}, d["e2e"]["value"], d["roofline"]["duration_us"])
except Exception as e: print("parse failed", e); print(open('gpurun_out/bench_ol$v.err').read()[-1500:])
PY
done
