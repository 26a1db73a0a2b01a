#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_all.log 2>&1; echo "all gpu tests exit $?"
tail -6 gpurun_out/r02_pytest_gpu_all.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r02_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench exit $?"
tail -3 gpurun_out/r02_bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n1.json'))
for k in ('value','ms_per_step','e2e','gpu_launches','clocks'): print(k, d[k])
r=d['roofline']; print({k:r[k] for k in ('achieved','frac','frac_vs_split_ceiling','launches','avg_launch_us','share_of_step','isolated_tflops','kernel')})
for k in ('train','sharded','similarity','mmr','cpu_baseline'): print(k, json.dumps(d.get(k))[:900])
print('kernels', json.dumps(d.get('kernels'))[:600]); print('kernels4m', json.dumps(d.get('kernels_4m_rows'))[:600])
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref exit $?"; cut -c1-400 gpurun_out/r02_bench_ref.json
