#!/bin/bash
mkdir -p gpurun_out
PARITY_QUICK=2 PARITY_OUT=r02_parity_65536_centred.md timeout 900 python scripts/parity_report.py tf32x3 > gpurun_out/parity.log 2>&1; echo "parity exit $?"; grep -E "initial_deep|layer1.weight|layer2.weight|logits|embedding" gpurun_out/r02_parity_65536_centred.md
PARITY_QUICK=1 PARITY_OUT=r02_parity_4096_centred.md timeout 900 python scripts/parity_report.py tf32x3 > gpurun_out/parity1.log 2>&1; echo "parity exit $?"; grep -E "initial_deep|layer1.weight|layer2.weight|logits|embedding" gpurun_out/r02_parity_4096_centred.md
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_gemm.py -m gpu -q -x > gpurun_out/r02_pytest_model.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/r02_pytest_model.log
timeout 300 python scripts/train_probe.py tf32x3 20
