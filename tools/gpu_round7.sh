#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/gemm_phases.py 2>&1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
for k, v in d.items():
    print(k, v['launch_us'], v['cta_total_us_median'], v['cta_total_us_max'], [(t['ctas'], t['wait_staging_us'], t['wait_acc_us'], t['epilogue_us']) for t in v['tiles']], v['final_store_us'])
"
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q --tb=short -x > gpurun_out/pytest_r7.log 2>&1
echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_r7.log | tail -15
timeout 600 python bench.py --steps 3 --warmup 3 --no-baselines > gpurun_out/bench_r7.json 2> gpurun_out/bench_r7.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_r7.json').read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_frame","sections_us")}, d["e2e"]["value"], d["roofline"]["duration_us"])
except Exception as e: print("parse failed", e); print(open('gpurun_out/bench_r7.err').read()[-1500:])
PY
