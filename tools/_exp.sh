python -m pytest tests -m gpu -x -q > gpurun_out/r02x_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02x_pytest.log
for m in fast tc; do
  python bench.py --steps 5 --warmup 3 --mode $m --no-cpu-baseline --no-gpu-eager --no-e2e > gpurun_out/r02x_$m.json 2> gpurun_out/r02x_$m.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r02x_$m.json"))
print("$m", round(d["ms_per_step"],2), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
for r in d["roofline_kernels"][:16]: print("   %-45s %8.3f" % (r["kernel"], r["ms_per_step"]))
PY
done
