"""Markdown table of the counters that matter for the roofline discussion from an `ncu --set full` report:
    ncu -i rep.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_summary.py raw.csv [title] > profiles/rNN_ncu_*.md
One row per captured launch: duration, pipe utilisation (tensor / XU = MUFU / FMA / ALU / LSU, % of peak sustained active),
DRAM bytes read + written, registers per thread, shared-memory bank conflicts per shared wavefront."""
import csv
import re
import sys

COLS = [("gpu__time_duration.sum", "ms", "time"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %", 1),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU %", 1),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA %", 1),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU %", 1),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU %", 1),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %", 1),
        ("dram__bytes_read.sum", "DRAM rd GB", None), ("dram__bytes_write.sum", "DRAM wr GB", None),
        ("launch__registers_per_thread", "regs", 1), ("smsp__inst_executed.sum", "warp-instr (M)", 1e-6)]
UNIT = {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0, "Tbyte": 1e3}
TIME = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    title = sys.argv[2] if len(sys.argv) > 2 else "ncu --set full summary"
    h, units = rows[0], rows[1]
    ix = {name: h.index(name) for name, _, _ in COLS if name in h}
    kn = h.index("Kernel Name")
    bc = h.index("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum") if "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum" in h else None
    wf = h.index("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum") if "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum" in h else None
    print("# %s\n" % title)
    print("| kernel | " + " | ".join(label for name, label, _ in COLS if name in ix) + " | smem conflicts / wavefront |")
    print("|---|" + "---:|" * (len(ix) + 1))
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[kn]).replace("void ", "").replace("sapcu::", "").replace("(int)", "")
        cells = []
        for cname, _, scale in COLS:
            if cname not in ix:
                continue
            v = float(r[ix[cname]].replace(",", "") or 0)
            if scale is None:
                v *= UNIT.get(units[ix[cname]], 1e-9)
                cells.append("%.3f" % v)
            elif scale == "time":
                cells.append("%.3f" % (v * TIME.get(units[ix[cname]], 1e-6)))
            else:
                v *= scale
                cells.append(("%.2f" % v) if v < 100 else ("%.0f" % v))
        conf = ""
        if bc is not None and wf is not None:
            w = float(r[wf].replace(",", "") or 0)
            conf = "%.2f" % (float(r[bc].replace(",", "") or 0) / w) if w else ""
        print("| `%s` | %s | %s |" % (name, " | ".join(cells), conf))


if __name__ == "__main__":
    main()
