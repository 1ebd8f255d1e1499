#!/bin/bash
# Round-end evidence run on one B200 (gpurun): GPU tests with the parity archive, smoke(), DRAM traffic of the contraction
# family, the bench lines, ncu launch lists and `--set full` captures.  Everything lands in gpurun_out/ (copied to profiles/ here).
set -u
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-e2e"
SAPCU_PARITY_JSON=gpurun_out/r02_parity.json python -m pytest tests -m gpu -x -q > gpurun_out/r02f_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -2 gpurun_out/r02f_pytest.log
[ $rc -ne 0 ] && { grep -n "^E \|Error" gpurun_out/r02f_pytest.log | head -20; exit 1; }
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02f_smoke.log 2>&1; echo "smoke rc=$?"; grep smoke gpurun_out/r02f_smoke.log
for m in tc fast; do
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:"gemm_tc|edge_pos" --csv --log-file /tmp/traffic_$m.csv $B --mode $m > /dev/null 2>&1
  python tools/gemm_traffic.py /tmp/traffic_$m.csv profiles/r02_gemm_traffic_$m.json && cp profiles/r02_gemm_traffic_$m.json gpurun_out/
done
python bench.py > gpurun_out/r02_bench_cfg1_tc_final.json 2> gpurun_out/r02f_tc.err; echo "bench tc rc=$?"
python bench.py --mode fast > gpurun_out/r02_bench_cfg1_fast_final.json 2> gpurun_out/r02f_fast.err; echo "bench fast rc=$?"
python bench.py --init stress --no-cpu-baseline --no-gpu-eager > gpurun_out/r02_bench_cfg1_tc_stress_init.json 2>/dev/null
python bench.py --init stress --mode fast --no-cpu-baseline --no-gpu-eager > gpurun_out/r02_bench_cfg1_fast_stress_init.json 2>/dev/null
python - <<'PY'
import json
for n in ("tc_final", "fast_final", "tc_stress_init", "fast_stress_init"):
    try:
        d = json.load(open("gpurun_out/r02_bench_cfg1_%s.json" % n))
        print(n, round(d["ms_per_step"], 2), round(d["value"]), (d.get("e2e") or {}).get("value"), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    except Exception as e:
        print(n, "failed", e)
PY
for m in tc fast; do
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_$m.csv $B --mode $m > /dev/null 2>&1
  python tools/launch_summary.py gpurun_out/r02_launches_$m.csv --md > gpurun_out/r02_launches_${m}_body.md 2>&1
done
bash tools/ncu_capture.sh r02_ncu_tc_fn "gemm_tc|edge_pos" 60 python tools/profile_fn.py --S 1024 --mode tc
bash tools/ncu_capture.sh r02_ncu_fast_fn "gemm_tc|edge_pos" 60 python tools/profile_fn.py --S 1024 --mode fast
bash tools/ncu_capture.sh r02_ncu_fast_fd "." 40 python tools/profile_fn.py --S 1024 --mode fast --fd
