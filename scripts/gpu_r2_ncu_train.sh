#!/bin/bash
# ncu --set full captures of the training step's two tensor-core kernels and the radix sort (after the plain run exited 0)
mkdir -p gpurun_out
python scripts/train_probe.py tf32x3 3 > gpurun_out/r02_train_probe.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_wgrad_tc" -s 10 -c 2 -f -o gpurun_out/prof_wgrad_r02 \
    python scripts/train_probe.py tf32x3 3 > gpurun_out/ncu_wgrad.log 2>&1
echo "ncu wgrad exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_tc" -s 22 -c 2 -f -o gpurun_out/prof_gemm_r02 \
    python scripts/train_probe.py tf32x3 3 > gpurun_out/ncu_gemm.log 2>&1
echo "ncu gemm exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_radix_sort_coop" -s 3 -c 1 -f -o gpurun_out/prof_sort_r02 \
    python scripts/train_probe.py tf32x3 3 > gpurun_out/ncu_sort.log 2>&1
echo "ncu sort exit $?"
cat gpurun_out/r02_train_probe.log
