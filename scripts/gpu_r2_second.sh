#!/bin/bash
# round 2, second GPU call: fused tower parity + timing, tightened gradient parity
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tower.py -m gpu -x -q > gpurun_out/r02_pytest_tower.log 2>&1; echo "tower tests exit $?"
tail -25 gpurun_out/r02_pytest_tower.log
timeout 300 python scripts/tower_probe.py 4194304 > gpurun_out/r02_tower_probe.log 2>&1; echo "tower probe exit $?"
cat gpurun_out/r02_tower_probe.log | tail -12
timeout 1500 python -m pytest tests/test_gpu_model.py tests/test_gpu_gemm.py -m gpu -q > gpurun_out/r02_pytest_model.log 2>&1; echo "model tests exit $?"
tail -30 gpurun_out/r02_pytest_model.log
