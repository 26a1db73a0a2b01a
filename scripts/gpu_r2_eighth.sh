#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_serving.py -m gpu -q -x > gpurun_out/r02_pytest_k7.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/r02_pytest_k7.log
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
# ncu: launch list of the (shortened) bench command, after it has run once without ncu; then one full capture of the tower kernel
timeout 300 python bench.py --steps 2 --warmup 1 --skip-extras --requests 8192 > gpurun_out/profile_plain.json 2> gpurun_out/profile_plain.err \
 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_fp16x3.csv \
      python bench.py --steps 2 --warmup 1 --skip-extras --requests 8192 > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"
timeout 120 python scripts/tower_debug.py 1048576 0 > gpurun_out/tower_plain.log 2>&1 \
 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tower_eval -c 1 -f -o gpurun_out/r02_prof_tower \
      python scripts/tower_debug.py 1048576 0 > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; tail -3 gpurun_out/ncu_full.log
