#!/bin/bash
mkdir -p gpurun_out
DCNR_L0_FP32=1 PARITY_QUICK=2 PARITY_OUT=r02_parity_65536_l0.md timeout 600 python scripts/parity_report.py tf32x3 > /dev/null 2> gpurun_out/p4.err; echo "parity exit $?"
DCNR_WG_FP32=1 PARITY_QUICK=2 PARITY_OUT=r02_parity_65536_wg.md timeout 600 python scripts/parity_report.py tf32x3 > /dev/null 2> gpurun_out/p5.err; echo "parity exit $?"
PARITY_QUICK=2 PARITY_OUT=r02_parity_65536_fp32.md timeout 600 python scripts/parity_report.py fp32 > /dev/null 2> gpurun_out/p6.err; echo "parity exit $?"
paste -d'|' <(grep "^| " gpurun_out/r02_parity_65536_l0.md | cut -d'|' -f2,4,3) <(grep "^| " gpurun_out/r02_parity_65536_wg.md | cut -d'|' -f3) <(grep "^| " gpurun_out/r02_parity_65536_fp32.md | cut -d'|' -f3)
