#!/bin/bash
mkdir -p gpurun_out
for Q in 8 32 64 100 1024; do timeout 300 python scripts/knn_scan_probe.py 10000000 16 $Q 2>&1 | tail -1; done
timeout 300 python scripts/knn_scan_probe.py 10000000 64 32 2>&1 | tail -1
timeout 300 python scripts/knn_scan_probe.py 1250000 16 32 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_knn.py tests/test_gpu_serving.py -m gpu -q -x > gpurun_out/r02_pytest_knn.log 2>&1; echo "tests exit $?"; tail -5 gpurun_out/r02_pytest_knn.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/knn_launches_32.csv python scripts/knn_scan_probe.py 10000000 16 32 > gpurun_out/ncu_knn.log 2>&1
grep -E "k_knn_tc" gpurun_out/knn_launches_32.csv | awk -F'","' '{print substr($5,1,24), $NF}' | tail -5
