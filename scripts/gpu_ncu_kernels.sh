#!/bin/bash
# ncu --set full captures of the non-GEMM kernels (one launch each), after the plain runs exited 0.
mkdir -p gpurun_out
python scripts/kernel_probe.py 1048576 > gpurun_out/kp_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_embed_cross_fwd|k_scatter_window|k_bn_stats_partial|k_bn_act_fwd" -s 40 -c 8 -f -o gpurun_out/prof_kernels_r01 \
    python scripts/kernel_probe.py 1048576 > gpurun_out/ncu_kernels.log 2>&1
echo "ncu kernels exit $?"
python scripts/knn_probe.py 10000000 16 201 1 > gpurun_out/knn_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_knn_stream" -s 3 -c 2 -f -o gpurun_out/prof_knn_r01 \
    python scripts/knn_probe.py 10000000 16 201 1 > gpurun_out/ncu_knn.log 2>&1
echo "ncu knn exit $?"
python scripts/train_probe.py tf32x3 3 > gpurun_out/train_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_wgrad_tc" -s 10 -c 2 -f -o gpurun_out/prof_wgrad_r01 \
    python scripts/train_probe.py tf32x3 3 > gpurun_out/ncu_wgrad.log 2>&1
echo "ncu wgrad exit $?"
