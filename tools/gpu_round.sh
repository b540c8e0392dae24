#!/bin/bash
# one gpurun call: the whole GPU suite + the bench.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | tail -25
grep -E "rel err|engine vs|thresholds" gpurun_out/pytest_gpu.log | cut -c1-700 | tail -12
timeout 900 python bench.py --steps 4 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -5 gpurun_out/bench.err | cut -c1-400
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
    keep={k:d.get(k) for k in ("value","ms_per_step","ms_per_frame","e2e","sections_us","clocks","gpu_launches","exact_two_phase","cpu_baseline","gpu_reference","msda_backward")}
    print(json.dumps(keep)[:3000]); print(json.dumps(d.get("roofline"))[:1200]); print(json.dumps(d.get("roofline_tensor"))[:600]); print(json.dumps(d.get("msda_sweep"))[:2500])
except Exception as e: print("bench parse failed", e)
PY
# ncu: launch list of a short bench (no kernel-name filter) + full captures of the decoder / updater cluster kernels and the FFN
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 330 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --clip-frames 2 --no-baselines > gpurun_out/launches_bench.log 2>&1; echo "ncu launches rc=$?"
MEMOTR_NONCOOP=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:decoder_cluster_kernel -s 2 -c 1 -o gpurun_out/r02_decoder_cluster -f python tools/prof_decoder.py > gpurun_out/ncu_decoder.log 2>&1; echo "ncu decoder rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:mlp2_tc_kernel -s 4 -c 1 -o gpurun_out/r02_mlp2_lnout -f python tools/prof_decoder.py > gpurun_out/ncu_mlp2.log 2>&1; echo "ncu mlp2 rc=$?"
tail -2 gpurun_out/ncu_decoder.log gpurun_out/ncu_mlp2.log
