#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_knn.py -m gpu -q -x > gpurun_out/r02_pytest_knn.log 2>&1; echo "tests exit $?"; tail -5 gpurun_out/r02_pytest_knn.log
for Q in 8 32 100 256 1024; do timeout 300 python scripts/knn_scan_probe.py 10000000 16 $Q 2>&1 | tail -1; done
timeout 300 python scripts/knn_scan_probe.py 10000000 64 1024 2>&1 | tail -1
