"""DRAM traffic of the contraction kernels over ONE pipeline pass, from an ncu metrics csv:
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:gemm_tc --csv \
        --log-file traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline
    python tools/gemm_traffic.py traffic.csv profiles/r01_gemm_traffic.json
Every pass (warm-up, timed, end-to-end) launches the same kernel sequence; the last period of it is summed.
"""
import csv
import json
import sys


def main():
    src, dst = sys.argv[1], sys.argv[2]
    lines = [l for l in open(src) if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    ids = sorted({int(r["ID"]) for r in rows})
    names = {int(r["ID"]): r["Kernel Name"] + r["Grid Size"] for r in rows}
    seq = [names[i] for i in ids]
    per = next(p for p in range(1, len(seq) + 1)            # launches of one pass = the period of the launch sequence
               if len(seq) % p == 0 and all(seq[i] == seq[i + p] for i in range(len(seq) - p)))
    last = set(ids[-per:])
    tot = {"dram__bytes_read.sum": 0.0, "dram__bytes_write.sum": 0.0, "gpu__time_duration.sum": 0.0}
    other = {"dram__bytes_read.sum": 0.0, "dram__bytes_write.sum": 0.0, "gpu__time_duration.sum": 0.0}
    for r in rows:
        if int(r["ID"]) in last and r["Metric Name"] in tot and "gemm_tc" not in r["Kernel Name"]:     # e.g. edge_pos_lif: reported beside the family
            v = float(r["Metric Value"].replace(",", ""))
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r["Metric Unit"])
            other[r["Metric Name"]] += v * (scale or 0.0)
            continue
        if int(r["ID"]) in last and r["Metric Name"] in tot:
            v = float(r["Metric Value"].replace(",", ""))
            u = r["Metric Unit"]
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(u)
            if scale is None:
                raise SystemExit("unexpected unit %r" % u)
            tot[r["Metric Name"]] += v * scale
    out = {
        "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:gemm_tc|edge_pos on "
                  "`python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-e2e [--mode ...]` (last pipeline pass); tools/gemm_traffic.py",
        "kernel_family": "gemm_tc_kernel + gemm_tc2_kernel",
        "launches_per_step": per,
        "dram_read_bytes_per_step": tot["dram__bytes_read.sum"],
        "dram_write_bytes_per_step": tot["dram__bytes_write.sum"],
        "traffic_bytes_per_step": tot["dram__bytes_read.sum"] + tot["dram__bytes_write.sum"],
        "kernel_ms_per_step_under_ncu": tot["gpu__time_duration.sum"],
        "other_captured_kernels": {"what": "edge_pos_lif (fc_delta K = 3 + LIF, not a tensor-core contraction)",
                                   "traffic_bytes_per_step": other["dram__bytes_read.sum"] + other["dram__bytes_write.sum"],
                                   "kernel_ms_per_step_under_ncu": other["gpu__time_duration.sum"]},
    }
    if len(sys.argv) > 3:
        out["mode"] = sys.argv[3]
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
