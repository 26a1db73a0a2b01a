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
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -k "large_batch" > gpurun_out/r02_pytest_model.log 2>&1; echo "model tests exit $?"
tail -8 gpurun_out/r02_pytest_model.log
