#!/bin/bash
L=hybrid-hotel-recommendation-system-based-on-friends-recommendations_b200/lib/libdcnr_sm100a.so
for round in 1 2; do
for v in a b; do
cp build/ab/lib_$v.so $L
echo "== $v"
timeout 300 python scripts/train_probe.py tf32x3 30
done
done
cp build/ab/lib_b.so $L
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_gemm.py tests/test_gpu_cross_v2.py -m gpu -q -x 2>&1 | tail -2
