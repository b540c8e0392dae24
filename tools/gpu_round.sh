#!/bin/bash
# one gpurun call: tests, micro-benchmark, bench, ncu of the gather.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -s -x --deselect tests/test_engine_gpu.py::test_engine_bf16_white_noise_weights_vs_matched_rounding_oracle > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python -m pytest tests/test_engine_gpu.py -m gpu -q --tb=short -s -k "teacher_forced or reference_init or matched_rounding" > gpurun_out/pytest_numerics.log 2>&1
echo "numerics rc=$?" | tee -a gpurun_out/pytest_numerics.log
grep -E "rel err|engine vs|passed|failed|Error|assert" gpurun_out/pytest_numerics.log | cut -c1-1500 | tail -20
timeout 600 python tools/micro_msda.py > gpurun_out/micro_msda.log 2>&1; echo "micro rc=$?"
cut -c1-600 gpurun_out/micro_msda.log | tail -14
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cut -c1-3000 gpurun_out/bench.json | tail -3; tail -3 gpurun_out/bench.err
timeout 600 ncu --set full --import-source on --clock-control none -k regex:msda_window_kernel -s 2 -c 1 -o gpurun_out/r02_msda_window -f python tools/micro_msda.py --ncu > gpurun_out/ncu_window.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_window.log
