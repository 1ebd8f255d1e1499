run() { # name, mode, env...
  name=$1; shift; mode=$1; shift
  env "$@" python bench.py --steps 5 --warmup 3 --mode $mode --no-cpu-baseline --no-gpu-eager --no-e2e > gpurun_out/r02aa_$name.json 2> gpurun_out/r02aa_$name.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r02aa_$name.json"))
print("$name", round(d["ms_per_step"],2), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
for r in d["roofline_kernels"][:5]: print("   %-45s %8.3f" % (r["kernel"], r["ms_per_step"]))
PY
}
run small1 tc SAPCU_TC_SMALL_TAB=1
run small0 tc SAPCU_TC_SMALL_TAB=0
run small1b tc SAPCU_TC_SMALL_TAB=1
run small0b tc SAPCU_TC_SMALL_TAB=0
python -m pytest tests -m gpu -x -q -k "tensor_core or benchmarked or block_taps or chunk" > gpurun_out/r02aa_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02aa_pytest.log
