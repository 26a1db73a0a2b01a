#!/bin/bash
mkdir -p gpurun_out
for args in "300 0x100" "2961 0x100" "100000 0" "100000 0 bf16"; do timeout 300 python scripts/tower_debug.py $args 2>&1 | tail -1; done
timeout 900 python -m pytest tests/test_gpu_tower.py -m gpu -q -x > gpurun_out/r02_pytest_tower.log 2>&1; echo "tower tests exit $?"; tail -3 gpurun_out/r02_pytest_tower.log
timeout 300 python scripts/tower_probe.py 4194304 > gpurun_out/r02_tower_probe.log 2>&1; tail -6 gpurun_out/r02_tower_probe.log
timeout 300 python scripts/graph_debug.py > gpurun_out/r02_graph_debug.log 2>&1; tail -16 gpurun_out/r02_graph_debug.log
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -x -k "scatter or golden or determin" > gpurun_out/r02_pytest_k7.log 2>&1; echo "k7 tests exit $?"; tail -3 gpurun_out/r02_pytest_k7.log
python - <<'PY' > gpurun_out/r02_kernel_probe.json 2> gpurun_out/r02_kernel_probe.err
import sys, json; sys.path.insert(0, "scripts"); sys.path.insert(0, ".")
import kernel_probe, bench
pk = bench.peaks()
for B in (65536, 1 << 22):
    r = kernel_probe.probe(B, pk["hbm"])
    print(B, json.dumps({k: {kk: vv for kk, vv in v.items() if kk != "note"} for k, v in r.items() if k.startswith(("K7",))}))
PY
cat gpurun_out/r02_kernel_probe.json; tail -2 gpurun_out/r02_kernel_probe.err
