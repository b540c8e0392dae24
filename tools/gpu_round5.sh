#!/bin/bash
mkdir -p gpurun_out
for d in 0 2; do MEMOTR_WINDOW_DEBUG=$d timeout 300 python tools/micro_msda.py --quick 2>&1 | tail -2 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('debug', '$d', d['case'], 'eq', d['bit_equal'], d['generic_bit_equal'], 'win', round(d['window_us'],1), 'generic', round(d['window_generic_us'],1), 'global', round(d['global_us'],1))
"; done
timeout 900 python -m pytest tests/test_msda_window_gpu.py -m gpu -q --tb=short -x > gpurun_out/pytest_r5.log 2>&1
echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_r5.log | tail -5
timeout 600 python bench.py --steps 3 --warmup 3 --no-baselines > gpurun_out/bench_r5.json 2> gpurun_out/bench_r5.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_r5.json').read().strip().splitlines()[-1])
    print(round(d["value"],1), round(d["ms_per_frame"],4), "e2e", round(d["e2e"]["value"],1), d["sections_us"], round(d["roofline"]["duration_us"],1))
except Exception as e: print("parse failed", e); print(open('gpurun_out/bench_r5.err').read()[-2500:])
PY
