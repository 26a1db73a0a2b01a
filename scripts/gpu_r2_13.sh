#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tower.py tests/test_gpu_model.py tests/test_gpu_serving.py -m gpu -q -x -k "not large_batch" > gpurun_out/r02_pytest_tower.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/r02_pytest_tower.log
timeout 300 python scripts/tower_probe.py 4194304 > gpurun_out/r02_tower_probe.log 2>&1; tail -6 gpurun_out/r02_tower_probe.log
timeout 600 python bench.py --steps 5 --warmup 3 --skip-extras > gpurun_out/r02_bench_quick.json 2> gpurun_out/r02_bench_quick.err; echo "bench exit $?"; tail -2 gpurun_out/r02_bench_quick.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_quick.json'))
for k in ('value','ms_per_step','e2e','gpu_launches','clocks'): print(k, d[k])
r=d['roofline']; print({k:r[k] for k in ('achieved','frac','frac_vs_split_ceiling','launches','avg_launch_us','share_of_step','isolated_tflops')})
PY
