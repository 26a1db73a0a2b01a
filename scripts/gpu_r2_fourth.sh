#!/bin/bash
mkdir -p gpurun_out
{
for args in "300 0x101" "300 0x200" "2961 0x101" "2961 0x200" "2961 0x301" "37001 0x1" "37001 0x0" "1000000 0x1" "1000000 0x0"; do
  timeout 120 python scripts/tower_debug.py $args 2>&1 | tail -1
done
} > gpurun_out/r02_tower_debug.log 2>&1
cat gpurun_out/r02_tower_debug.log
timeout 900 python -m pytest tests/test_gpu_tower.py -m gpu -x -q > gpurun_out/r02_pytest_tower.log 2>&1; echo "tower tests exit $?"
tail -8 gpurun_out/r02_pytest_tower.log
timeout 300 python scripts/tower_probe.py 4194304 > gpurun_out/r02_tower_probe.log 2>&1; echo "tower probe exit $?"
tail -12 gpurun_out/r02_tower_probe.log
PARITY_QUICK=2 PARITY_OUT=r02_parity_65536_tc.md timeout 600 python scripts/parity_report.py tf32x3 > /dev/null 2> gpurun_out/p1.err; echo "parity exit $?"
DCNR_W0_FP32=1 PARITY_QUICK=2 PARITY_OUT=r02_parity_65536_w0fp32.md timeout 600 python scripts/parity_report.py tf32x3 > /dev/null 2> gpurun_out/p2.err; echo "parity exit $?"
paste -d'|' <(grep "^| " gpurun_out/r02_parity_65536_tc.md | cut -d'|' -f2,3,4) <(grep "^| " gpurun_out/r02_parity_65536_w0fp32.md | cut -d'|' -f3)
