"""Summarise the source page of an ncu report: per kernel, total stall samples by reason and the hottest SASS lines.
    ncu -i rep.ncu-rep --page source --csv > src.csv ; python tools/ncu_stalls.py src.csv [kernel-substring] [top]
"""
import csv
import sys


def main():
    path = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    rows = list(csv.reader(open(path)))
    tables, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "data": []}
            tables.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and len(r) == len(cur["hdr"]):
            cur["data"].append(r)
    for t in tables:
        if want not in t["name"]:
            continue
        h, data = t["hdr"], t["data"]
        isrc, isamp = h.index("Source"), h.index("# Samples")
        stalls = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
        tot = sum(int(r[isamp]) for r in data)
        agg = {}
        for r in data:
            for i in stalls:
                agg[h[i]] = agg.get(h[i], 0) + int(r[i])
        print("==", t["name"][:90], "samples", tot, "instrs", len(data))
        print("  ", [(k, round(100.0 * v / max(tot, 1), 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]])
        top = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:top_n]
        for i in sorted(top):
            r = data[i]
            s = {h[j]: int(r[j]) for j in stalls if int(r[j]) > 0}
            print("   %4d %-58s %6s %s" % (i, r[isrc].strip()[:58], r[isamp], sorted(s.items(), key=lambda kv: -kv[1])[:2]))


if __name__ == "__main__":
    main()
