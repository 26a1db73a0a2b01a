#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s > gpurun_out/r02_pytest_multi_n2.log 2>&1; echo "multi exit $?"; tail -3 gpurun_out/r02_pytest_multi_n2.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench n2 exit $?"
tail -5 gpurun_out/r02_bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n2.json'))
for k in ('value','ms_per_step','e2e','n_gpus'): print(k, d[k])
for k in ('train','train_weak','sharded','similarity','scaling_extras'): print(k, json.dumps(d.get(k))[:1500])
PY
