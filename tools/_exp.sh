python -m pytest tests -m gpu -x -q > gpurun_out/r02h_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02h_pytest.log; grep -n "^E  " gpurun_out/r02h_pytest.log | head -5
run() { # name, mode, env...
  name=$1; shift; mode=$1; shift
  env "$@" python bench.py --steps 5 --warmup 3 --mode $mode --no-cpu-baseline --no-gpu-eager --no-e2e > gpurun_out/r02h_$name.json 2> gpurun_out/r02h_$name.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r02h_$name.json"))
print("$name", round(d["ms_per_step"],2), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
for r in d["roofline_kernels"]:
    if r["kernel"] in ("fn.fc1+lif","fd.edgeconv(per-point P|Q)","fn.qkv+lif"): print("   %-45s %8.3f" % (r["kernel"], r["ms_per_step"]))
PY
}
run new tc A=1
run old tc SAPCU_TC_PQ_FP16X3=0 SAPCU_TC_FC1_TABLE=0
