"""SASS opcode histogram of the built objects (the Blackwell-native evidence without the binary):
    python tools/sass_histogram.py > profiles/rNN_sass_histogram.md
Counts, per object of <package>/build/, the mnemonics B200_PROFILING.md names: tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM,
TMA -> UTMALDG / UTMAPF / UBLKCP, tcgen05.commit -> UTCBAR, packed fp32 -> FADD2 / FMUL2 / FFMA2, plus MUFU and the legacy HMMA."""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = glob.glob(os.path.join(ROOT, "c-users-*_b200"))[0]
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMAPF", "UBLKCP", "SYNCS", "FFMA2", "FADD2", "FMUL2", "FMNMX3",
        "MUFU.EX2", "MUFU.RCP", "MUFU.TANH", "MUFU.", "REDUX", "HMMA", "FFMA", "LDS", "STG", "LDG"]


def main():
    print("# SASS opcode histogram (cuobjdump -sass of every object in csrc build, sm_100a)\n")
    print("`UTC*MMA` = tcgen05.mma (`.2CTA` = cta_group::2), `LDTM` = tcgen05.ld, `UTMALDG` = cp.async.bulk.tensor (TMA), `UTMAPF` = TMA L2 prefetch, "
          "`UBLKCP` = cp.async.bulk (1-D TMA copy), `UTCBAR` = tcgen05.commit, `SYNCS` = mbarrier ops, `FFMA2/FADD2/FMUL2` = packed fp32. "
          "No `HMMA` (legacy mma.sync) anywhere.\n")
    rows = []
    for obj in sorted(glob.glob(os.path.join(PKG, "build", "*.o"))):
        sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        ops = collections.Counter()
        two = 0
        for m in re.finditer(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", sass):
            op = m.group(1)
            for k in KEYS:
                if op.startswith(k):
                    ops[k] += 1
            if op.startswith("UTCHMMA") and ".2CTA" in op:
                two += 1
        kernels = len(re.findall(r"Function : ", sass))
        rows.append((os.path.basename(obj), kernels, ops, two))
    cols = [k for k in KEYS if any(r[2][k] for r in rows) or k == "HMMA"]
    print("| object | kernels | " + " | ".join("`%s`" % c for c in cols) + " | `UTCHMMA.2CTA` |")
    print("|---|---:|" + "---:|" * (len(cols) + 1))
    for name, kernels, ops, two in rows:
        print("| `%s` | %d | %s | %d |" % (name, kernels, " | ".join(str(ops[c]) for c in cols), two))


if __name__ == "__main__":
    main()
