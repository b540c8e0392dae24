#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -q --tb=short -s -k "mlp2 or dense_block or reference_init or teacher_forced or sine_embed" > gpurun_out/pytest_dense.log 2>&1
echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_dense.log | tail -15
grep -E "rel err|engine vs" gpurun_out/pytest_dense.log | cut -c1-400 | tail -6
for fb in 1 0; do
MEMOTR_FUSE_BLOCK=$fb timeout 600 python bench.py --steps 3 --warmup 3 --no-baselines > gpurun_out/bench_fb$fb.json 2> gpurun_out/bench_fb$fb.err; echo "bench fuse_block=$fb rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_fb$fb.json').read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_frame","sections_us")}, d["e2e"]["value"], d["roofline"]["duration_us"])
except Exception as e: print("parse failed", e); print(open('gpurun_out/bench_fb$fb.err').read()[-1500:])
PY
done
MEMOTR_FUSE_BLOCK=0 MEMOTR_FUSE_LN2=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-baselines > gpurun_out/bench_nofuse.json 2> gpurun_out/bench_nofuse.err; echo "bench nofuse rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_nofuse.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_frame","sections_us")})
PY
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_step.csv python tools/prof_step.py bf16 1 > gpurun_out/launches_step.log 2>&1; echo "ncu launches rc=$?"
