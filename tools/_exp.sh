run() { # name, mode, env...
  name=$1; shift; mode=$1; shift
  env "$@" python bench.py --steps 5 --warmup 3 --mode $mode --no-cpu-baseline --no-gpu-eager --no-e2e > gpurun_out/r02ad_$name.json 2> gpurun_out/r02ad_$name.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r02ad_$name.json"))
print("$name", round(d["ms_per_step"],2), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
for r in d["roofline_kernels"][:6]: print("   %-45s %8.3f" % (r["kernel"], r["ms_per_step"]))
PY
}
run copy1 tc SAPCU_TC_POS_COPY=1
run copy0 tc SAPCU_TC_POS_COPY=0
run copy1b tc SAPCU_TC_POS_COPY=1
run copy0b tc SAPCU_TC_POS_COPY=0
