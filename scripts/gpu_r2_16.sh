#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tower.py tests/test_gpu_serving.py tests/test_gpu_model.py -m gpu -q -x > gpurun_out/r02_pytest_tower.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/r02_pytest_tower.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_quick.json 2> gpurun_out/r02_bench_quick.err; echo "bench exit $?"; tail -2 gpurun_out/r02_bench_quick.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_quick.json'))
for k in ('value','ms_per_step','e2e','gpu_launches','clocks','request_latency'): print(k, d.get(k))
print('train', {k:d['train'][k] for k in ('ms_per_step','value','gpu_launches_per_step')}, d['train']['cuda_graph'])
PY
