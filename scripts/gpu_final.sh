#!/bin/bash
# Final evidence pass: launch list of the profiling form of the bench + N=1 bench (both arms).
mkdir -p gpurun_out
timeout 300 python bench.py --steps 2 --warmup 1 --skip-extras --precision tf32x3 --requests 8192 > gpurun_out/profile_plain.json 2> gpurun_out/profile_plain.err \
 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_tf32x3_final.csv \
      python bench.py --steps 2 --warmup 1 --skip-extras --precision tf32x3 --requests 8192 > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -s 6 -c 2 -f -o gpurun_out/prof_gemm_tf32x3_final \
      python scripts/gemm_probe.py tf32x3 1048576 3 > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
timeout 600 python bench.py > gpurun_out/bench_tf32x3.json 2> gpurun_out/bench_tf32x3.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref exit $?"
