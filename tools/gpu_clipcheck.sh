#!/bin/bash
# sub-clip lengths a rank sees at N = 1 / 8 / 21 on a 64-frame clip, run on one GPU (same code path minus NCCL)
timeout 600 python -m pytest tests/test_tracker_gpu.py -m gpu -q --tb=short -x -k "pipelined" 2>&1 | tail -3
for cf in 64 8 3; do
timeout 600 python bench.py --steps 4 --warmup 3 --no-baselines --clip-frames $cf > gpurun_out/bench_cf.json 2> gpurun_out/bench_cf.err; echo "clip-frames $cf rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_cf.json').read().strip().splitlines()[-1])
    print(round(d["value"],1), round(d["ms_per_frame"],4), "e2e", round(d["e2e"]["value"],1), d["config"]["frame_pipelining"][:12])
except Exception as e: print("parse failed", e); print(open('gpurun_out/bench_cf.err').read()[-1500:])
PY
done
