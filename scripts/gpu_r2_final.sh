#!/bin/bash
# Round-2 evidence pass (1 GPU): parity tests, smoke, bench (both arms), ncu launch lists of the eval bench and of a training
# step, full captures of the fused tower and of the gather.  Each ncu pass follows the same command run plain (exit 0).
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > $O/${TAG}_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $O/${TAG}_pytest_gpu_n1.log 2>&1; echo "pytest exit $?" | tee -a $O/${TAG}_pytest_gpu_n1.log
tail -3 $O/${TAG}_pytest_gpu_n1.log
timeout 300 python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1; echo "smoke exit $?" | tee -a $O/${TAG}_smoke.log
timeout 400 python bench.py --impl reference --steps 5 --warmup 3 > $O/${TAG}_bench_reference_n1.json 2> $O/${TAG}_bench_ref.err; echo "bench ref exit $?"
timeout 900 python bench.py > $O/${TAG}_bench_fp16x3_n1.json 2> $O/${TAG}_bench.err; echo "bench exit $?"; tail -c 300 $O/${TAG}_bench.err
CMD="python bench.py --steps 2 --warmup 1 --skip-extras --requests 8192"
timeout 300 $CMD > $O/${TAG}_profile_plain.json 2> $O/${TAG}_profile_plain.err \
 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/${TAG}_launches_fp16x3.csv $CMD > $O/ncu_launches.log 2>&1
echo "ncu launches exit $?"
timeout 300 python scripts/train_probe.py tf32x3 1 > $O/${TAG}_train_probe.log 2>&1 \
 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/${TAG}_launches_train.csv python scripts/train_probe.py tf32x3 1 > $O/ncu_train.log 2>&1
echo "ncu train exit $?"; cat $O/${TAG}_train_probe.log
timeout 300 python scripts/tower_probe.py 1048576 > $O/${TAG}_tower_probe_1m.log 2>&1 \
 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tower_eval -s 3 -c 2 -f -o $O/prof_tower_${TAG} python scripts/tower_probe.py 1048576 > $O/ncu_tower.log 2>&1
echo "ncu tower exit $?"
(cd scripts && timeout 200 python k1_probe.py 4194304 2 > ../$O/${TAG}_k1_probe.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_embed_cross_fwd -c 2 -f -o ../$O/prof_k1_${TAG} python k1_probe.py 4194304 2 > ../$O/ncu_k1.log 2>&1)
echo "ncu k1 exit $?"
timeout 300 python scripts/knn_scan_probe.py 10000000 16 1024 > $O/${TAG}_knn_tc_probe_q1024.log 2>&1 \
 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_knn_tc_scan -s 3 -c 1 -f -o $O/prof_knn_tc_${TAG} python scripts/knn_scan_probe.py 10000000 16 1024 > $O/ncu_knn_tc.log 2>&1
echo "ncu knn exit $?"
timeout 600 python scripts/knn_tc_probe.py 10000000 > $O/${TAG}_knn_tc_probe.log 2>&1; tail -24 $O/${TAG}_knn_tc_probe.log
ls -la $O | tail -24
