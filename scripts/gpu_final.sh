#!/bin/bash
# Final evidence pass of a round (1 GPU): parity tests, smoke, the bench (both arms), the ncu launch list of the profiling
# form of the bench and full captures of the top kernel and of the gather.  Usage: bash scripts/gpu_final.sh [tag]
TAG=${1:-r01v6}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/smoke.log
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref exit $?"
timeout 600 python bench.py > gpurun_out/bench_tf32x3.json 2> gpurun_out/bench_tf32x3.err; echo "bench exit $?"
timeout 300 python bench.py --steps 2 --warmup 1 --skip-extras --precision tf32x3 --requests 8192 > gpurun_out/profile_plain.json 2> gpurun_out/profile_plain.err \
 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_tf32x3_${TAG}.csv \
      python bench.py --steps 2 --warmup 1 --skip-extras --precision tf32x3 --requests 8192 > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -s 6 -c 2 -f -o gpurun_out/prof_gemm_tf32x3_${TAG} \
      python scripts/gemm_probe.py tf32x3 1048576 3 > gpurun_out/ncu_full.log 2>&1
echo "ncu gemm exit $?"
(cd scripts && timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_embed_cross_fwd -c 2 -f -o ../gpurun_out/prof_k1_${TAG} \
      python k1_probe.py 4194304 2 > ../gpurun_out/ncu_k1.log 2>&1)
echo "ncu k1 exit $?"
tail -c 300 gpurun_out/bench_tf32x3.err
