mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
timeout 300 python -m pytest tests/test_gpu_gemm.py -x -q > gpurun_out/pytest_gemm.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gemm.log
tail -15 gpurun_out/pytest_gemm.log
fi
for p in ${PRECS:-tf32 tf32x3}; do timeout 120 python scripts/gemm_probe.py $p; done 2>&1 | tee gpurun_out/gemm_probe.log
for d in ${DEBUGS:-}; do DCNR_GEMM_DEBUG=$d timeout 120 python scripts/gemm_probe.py tf32x3 2>&1 | sed "s/^/debug$d /" | tee -a gpurun_out/gemm_probe.log; done
