"""Per-kernel summary of the LAST pipeline pass in an ncu launch list (--metrics gpu__time_duration.sum --csv).
    python tools/launch_summary.py gpurun_out/launches.csv [--md]
"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    md = "--md" in sys.argv
    lines = [l for l in open(path) if l.startswith('"')]
    L = [(x["Kernel Name"], float(x["Metric Value"].replace(",", "")))
         for x in csv.DictReader(lines) if x["Metric Name"] == "gpu__time_duration.sum"]
    starts = [i for i, x in enumerate(L) if "knn_seed" in x[0]]
    P = L[starts[-1] - 2:]          # the pass opens with two cloud_to_f32 launches and the seed kNN
    if len(starts) > 1 and len(P) < starts[-1] - starts[-2]:      # the capture window ended inside the last pass: take the last complete one
        P = L[starts[-2] - 2: starts[-1] - 2]
    P = [x for x in P if "at::native" not in x[0] and "elementwise_kernel" not in x[0]]      # torch's own fills / copies between the calls
    agg = collections.OrderedDict()
    for n, t in P:
        k = re.sub(r"\(.*", "", n).replace("void ", "").replace("sapcu::", "")
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(v[1] for v in agg.values())
    print(("total %.1f ms over %d launches" % (tot / 1e6, len(P))))
    if md:
        print("\n| kernel | launches | ms | share |\n|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if md:
            print("| `%s` | %d | %.2f | %.1f%% |" % (k, v[0], v[1] / 1e6, 100 * v[1] / tot))
        else:
            print("%-48s %4d %8.2f %5.1f%%" % (k, v[0], v[1] / 1e6, 100 * v[1] / tot))


if __name__ == "__main__":
    main()
