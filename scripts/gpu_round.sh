#!/bin/bash
# Full GPU pass: parity tests, smoke, bench (both arms), ncu launch list + one full capture of the top kernel.
# Usage (from the repo root, under gpurun): bash scripts/gpu_round.sh [tag]
TAG=${1:-r01}
PREC=${PREC:-tf32x3}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/smoke.log
timeout 600 python bench.py --precision $PREC > gpurun_out/bench_${PREC}.json 2> gpurun_out/bench_${PREC}.err; echo "bench $PREC exit $?"
tail -c 600 gpurun_out/bench_${PREC}.err
if [ -z "$SKIP_FP32" ]; then
timeout 600 python bench.py --precision fp32 --skip-extras > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; echo "bench fp32 exit $?"
fi
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref exit $?"
if [ -z "$SKIP_NCU" ]; then
# launch list of the same (shortened) command, after it has run once without ncu
timeout 300 python bench.py --steps 2 --warmup 1 --skip-extras --precision $PREC --requests 8192 > gpurun_out/profile_plain_${PREC}.json 2> gpurun_out/profile_plain_${PREC}.err \
 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_${PREC}_${TAG}.csv \
      python bench.py --steps 2 --warmup 1 --skip-extras --precision $PREC --requests 8192 > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -s 6 -c 2 -f -o gpurun_out/prof_gemm_${PREC}_${TAG} \
      python scripts/gemm_probe.py $PREC 1048576 3 > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
fi
cat gpurun_out/bench_${PREC}.json
