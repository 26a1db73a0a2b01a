#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -20
numactl -H 2>/dev/null | head -5; lscpu | grep -i "numa\|socket\|model name" | head
timeout 600 python bench.py --steps 5 --warmup 3 --skip-extras > gpurun_out/r02_bench_quick.json 2> gpurun_out/r02_bench_quick.err; echo "bench exit $?"; tail -2 gpurun_out/r02_bench_quick.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_quick.json'))
for k in ('value','ms_per_step','e2e'): print(k, d[k])
print(d['config'])
PY
