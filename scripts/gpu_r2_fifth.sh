#!/bin/bash
mkdir -p gpurun_out
DCNR_SPLITK2=1 PARITY_QUICK=2 PARITY_OUT=r02_parity_65536_k2.md timeout 600 python scripts/parity_report.py tf32x3 > /dev/null 2> gpurun_out/p3.err; echo "parity exit $?"
paste -d'|' <(grep "^| " gpurun_out/r02_parity_65536_k2.md | cut -d'|' -f2,3,4)
tail -3 gpurun_out/p3.err
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu_all.log 2>&1; echo "all gpu tests exit $?"
tail -12 gpurun_out/r02_pytest_gpu_all.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench exit $?"
tail -5 gpurun_out/r02_bench_n1.err
cat gpurun_out/r02_bench_n1.json | cut -c1-3000
