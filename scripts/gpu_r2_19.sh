#!/bin/bash
L=hybrid-hotel-recommendation-system-based-on-friends-recommendations_b200/lib/libdcnr_sm100a.so
for round in 1 2; do
for v in epi4 epi8; do
cp build/ab/lib_$v.so $L
echo "== $v"
timeout 300 python scripts/train_probe.py tf32x3 30
timeout 300 python scripts/gemm_probe.py tf32x3 65536 30 2>&1 | tail -2
done
done
