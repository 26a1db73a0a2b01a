#!/bin/bash
mkdir -p gpurun_out
for args in "300 0x100" "2961 0x100" "100000 0" "100000 0 bf16"; do timeout 300 python scripts/tower_debug.py $args 2>&1 | tail -1; done
timeout 900 python -m pytest tests/test_gpu_tower.py tests/test_gpu_serving.py -m gpu -q -x > gpurun_out/r02_pytest_tower.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/r02_pytest_tower.log
timeout 300 python scripts/tower_probe.py 4194304 > gpurun_out/r02_tower_probe.log 2>&1; tail -6 gpurun_out/r02_tower_probe.log
