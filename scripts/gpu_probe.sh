mkdir -p gpurun_out
for p in ${PRECS:-tf32 tf32x3}; do python scripts/gemm_probe.py $p; done 2>&1 | tee gpurun_out/gemm_probe.log
