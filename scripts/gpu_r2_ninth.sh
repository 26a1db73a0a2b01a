#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/tower_debug.py 300 0x100 2>&1 | tail -1
timeout 300 python scripts/tower_debug.py 100000 0 2>&1 | tail -1
timeout 300 python scripts/tower_debug.py 100000 1 2>&1 | tail -1
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu_all.log 2>&1; echo "tests exit $?"; tail -5 gpurun_out/r02_pytest_gpu_all.log
python - <<'PY' > gpurun_out/r02_kernel_probe.json 2> gpurun_out/r02_kernel_probe.err
import sys, json; sys.path.insert(0, "scripts"); sys.path.insert(0, ".")
import kernel_probe, bench
pk = bench.peaks()
for B in (65536, 1 << 22):
    r = kernel_probe.probe(B, pk["hbm"])
    print(B, json.dumps({k: {kk: vv for kk, vv in v.items() if kk != "note"} for k, v in r.items() if k.startswith(("K1", "K7"))}))
PY
cat gpurun_out/r02_kernel_probe.json; tail -2 gpurun_out/r02_kernel_probe.err
timeout 300 python scripts/tower_probe.py 4194304 > gpurun_out/r02_tower_probe.log 2>&1; tail -8 gpurun_out/r02_tower_probe.log
