#!/bin/bash
# round 2, first GPU call (2 GPUs): the N > 1 parity test with its log kept, the accumulator-rounding probe,
# and the tf32x3 per-tensor gradient table with and without a separate accumulator for the lo terms
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s > gpurun_out/r02_pytest_multi_n2.log 2>&1; echo "multi exit $?"
tail -5 gpurun_out/r02_pytest_multi_n2.log
timeout 300 python scripts/acc_probe.py > gpurun_out/r02_acc_probe.log 2>&1; echo "probe exit $?"
DCNR_GEMM_CORRSEP=1 timeout 300 python scripts/acc_probe.py > gpurun_out/r02_acc_probe_corrsep.log 2>&1; echo "probe2 exit $?"
PARITY_QUICK=1 PARITY_OUT=r02_parity_tf32x3_base.md timeout 600 python scripts/parity_report.py tf32x3 > /dev/null 2>gpurun_out/parity_base.err; echo "parity exit $?"
DCNR_GEMM_CORRSEP=1 PARITY_QUICK=1 PARITY_OUT=r02_parity_tf32x3_corrsep.md timeout 600 python scripts/parity_report.py tf32x3 > /dev/null 2>gpurun_out/parity_corr.err; echo "parity2 exit $?"
cat gpurun_out/r02_acc_probe.log
