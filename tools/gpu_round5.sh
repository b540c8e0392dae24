#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/micro_msda.py --quick 2>&1 | tail -2 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['case'], 'eq', d['bit_equal'], d['generic_bit_equal'], 'win', round(d['window_us'],1), 'generic', round(d['window_generic_us'],1), 'global', round(d['global_us'],1), 'warm', round(d['window_l2warm_us'],1))
"
MEMOTR_WINDOW_LEAD=0 timeout 300 python tools/micro_msda.py --quick 2>&1 | tail -2 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('lead0', d['case'], 'eq', d['bit_equal'], 'win', round(d['window_us'],1))
"
timeout 900 python -m pytest tests/test_msda_window_gpu.py -m gpu -q --tb=short -x > gpurun_out/pytest_r5.log 2>&1
echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_r5.log | tail -15
