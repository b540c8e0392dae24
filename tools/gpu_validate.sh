#!/bin/bash
# full GPU suite, smoke, default bench, reference arm, launch list of one eager step (no kernel-name filter)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -m gpu -q --tb=short > gpurun_out/pytest_gpu_full.log 2>&1
echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu_full.log | tail -8
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","sections_us","gpu_launches","clocks")}, d["e2e"], d["roofline"]["duration_us"], d["roofline"]["frac"], d.get("cpu_baseline"), d.get("gpu_reference"), d.get("exact_two_phase"))
except Exception as e: print("parse failed", e); print(open('gpurun_out/bench_default.err').read()[-2500:])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"; tail -c 600 gpurun_out/bench_reference.json
MEMOTR_NONCOOP=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_step.csv python tools/prof_step.py bf16 1 > gpurun_out/launches_step.log 2>&1; echo "ncu launches rc=$?"; tail -2 gpurun_out/launches_step.log
