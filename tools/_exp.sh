python bench.py > gpurun_out/r02_bench_cfg1_tc_final.json 2> gpurun_out/r02i_tc.err; echo "bench tc rc=$?"
python bench.py --mode fast --no-gpu-eager > gpurun_out/r02_bench_cfg1_fast_final.json 2> gpurun_out/r02i_fast.err; echo "bench fast rc=$?"
python - <<'PY'
import json
for n in ("tc_final", "fast_final"):
    d = json.load(open("gpurun_out/r02_bench_cfg1_%s.json" % n))
    print(n, round(d["ms_per_step"], 2), round(d["value"]), (d.get("e2e") or {}).get("value"), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
