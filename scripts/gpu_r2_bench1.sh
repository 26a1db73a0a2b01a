#!/bin/bash
O=gpurun_out; TAG=r02
mkdir -p $O
timeout 400 python bench.py --impl reference --steps 5 --warmup 3 > $O/${TAG}_bench_reference_n1.json 2> $O/${TAG}_bench_ref.err; echo "bench ref exit $?"
timeout 900 python bench.py > $O/${TAG}_bench_fp16x3_n1.json 2> $O/${TAG}_bench.err; echo "bench exit $?"; tail -c 300 $O/${TAG}_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_fp16x3_n1.json'))
for k in ('value','ms_per_step','e2e','gpu_launches','clocks','request_latency'): print(k, d.get(k))
print('roofline', {k:d['roofline'][k] for k in ('achieved','frac','traffic','frac_vs_split_ceiling','share_of_step','isolated_tflops')})
print('train', {k:d['train'][k] for k in ('ms_per_step','value','gpu_launches_per_step')}, d['train']['cuda_graph'])
print('similarity', json.dumps(d['similarity'])[:1500])
print('kernels', json.dumps(d['kernels'])[:900])
print('cpu', json.dumps(d['cpu_baseline'])[:600])
PY
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
