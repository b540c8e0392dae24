#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi -L | wc -l
MEMOTR_BENCH_TIME_EXCHANGE=${2:-0} timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 4 --warmup 3 --no-baselines > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?"
tail -1 gpurun_out/bench_n$N.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['rank_ms'], d.get('exchange_us'))
"
grep -iE "error|Traceback|NCCL WARN" gpurun_out/bench_n$N.err | head -5
