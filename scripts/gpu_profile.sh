# ncu evidence for the headline bench command (one GPU).  Each ncu pass is preceded by the same
# command run plain (exit 0 required), as the profiling recipe demands.
mkdir -p gpurun_out
PREC=${PREC:-tf32x3}
CMD="python bench.py --steps 2 --warmup 1 --skip-extras --precision $PREC --requests 8192"
$CMD > gpurun_out/profile_plain_$PREC.json 2> gpurun_out/profile_plain_$PREC.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$PREC.csv $CMD > gpurun_out/ncu_launches_$PREC.log 2>&1
echo "launch list exit $?"
$CMD > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"${KREGEX:-k_gemm_tc}" -s 4 -c 3 -o gpurun_out/prof_$PREC -f $CMD > gpurun_out/ncu_full_$PREC.log 2>&1
echo "full capture exit $?"
ls -la gpurun_out | tail -12
