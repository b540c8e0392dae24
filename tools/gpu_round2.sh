#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_msda_window_gpu.py tests/test_msda_gpu.py -m gpu -q --tb=short -x > gpurun_out/pytest_msda.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_msda.log
timeout 600 python tools/micro_msda.py > gpurun_out/micro_msda.log 2>&1; echo "micro rc=$?"
python - <<'PY'
import json
for r in json.load(open('gpurun_out/micro_msda.json')):
    print(r['case'], 'eq', r['bit_equal'], 'win %.1f us (%.3f) l2warm %.1f | glob %.1f us (%.3f)' % (r['window_us'], r['window_frac'], r['window_l2warm_us'], r['global_us'], r['global_frac']), 'staged', r['staged_points'], 'glob', r['global_points_in_staged_units'])
PY
timeout 600 ncu --set full --import-source on --clock-control none -k regex:msda_window_kernel -s 2 -c 1 -o gpurun_out/r02_msda_window_v2 -f python tools/micro_msda.py --ncu > gpurun_out/ncu_window.log 2>&1; echo "ncu rc=$?"
