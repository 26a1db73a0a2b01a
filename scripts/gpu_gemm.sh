mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py -m gpu -q --maxfail=100 -p no:cacheprovider > gpurun_out/pytest_gemm.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gemm.log
tail -40 gpurun_out/pytest_gemm.log
