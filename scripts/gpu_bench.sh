mkdir -p gpurun_out
for prec in ${PRECS:-fp32 tf32x3 tf32}; do
timeout 900 python bench.py --steps 3 --warmup 3 --precision $prec > gpurun_out/bench_$prec.json 2> gpurun_out/bench_$prec.err; echo "bench $prec exit $?"
tail -3 gpurun_out/bench_$prec.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$prec.json"))
    print("$prec", "value %.3e e2e %.3e ms/step %.1f roofline %.1f TF/s frac %.3f launches %d train %.3e (%.2f ms) clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["frac"], d["gpu_launches"], d["train"]["value"], d["train"]["ms_per_step"], d["clocks"]))
    print(" similarity", d.get("similarity"))
    print(" cpu", d.get("cpu_baseline"))
except Exception as e: print("parse fail", e)
PY
done
