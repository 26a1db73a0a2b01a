#!/bin/bash
# bench.py under torchrun on N GPUs of one box: bash scripts/gpu_r2_n8.sh N
N=${1:-8}
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_fp16x3_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench n$N exit $?"
tail -3 gpurun_out/r02_bench_n$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench_fp16x3_n$N.json'))
for k in ('value','ms_per_step','e2e','n_gpus','clocks'): print(k, d[k])
for k in ('train','train_weak','sharded','similarity','scaling_extras'): print(k, json.dumps(d.get(k))[:1800])
PY
