python -m pytest tests -m gpu -x -q > gpurun_out/r02ac_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02ac_pytest.log
run() { # name, mode, env...
  name=$1; shift; mode=$1; shift
  env "$@" python bench.py --steps 5 --warmup 3 --mode $mode --no-cpu-baseline --no-gpu-eager --no-e2e > gpurun_out/r02ac_$name.json 2> gpurun_out/r02ac_$name.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r02ac_$name.json"))
print("$name", round(d["ms_per_step"],2), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
for r in d["roofline_kernels"][:12]: print("   %-45s %8.3f" % (r["kernel"], r["ms_per_step"]))
PY
}
run tc tc A=1
run fast fast A=1
