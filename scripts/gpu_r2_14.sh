#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/wgrad_real_probe.py 65536 > gpurun_out/r02_wgrad_real_probe.log 2>&1; tail -12 gpurun_out/r02_wgrad_real_probe.log
