"""Print the SASS of one kernel of the in-tree library (cuobjdump), instructions only.  usage: sass_loop.py <mangled-name-substring> [start end]"""
import subprocess, sys, re, glob
so = glob.glob("c-users*/libsapcu_b200.so")[0]
names = subprocess.run(["cuobjdump", "-elf", so], capture_output=True, text=True).stdout  # noqa
fun = sys.argv[1]
out = subprocess.run(["cuobjdump", "-sass", "-fun", fun, so], capture_output=True, text=True).stdout
ins = []
for l in out.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 30
for a, t in ins:
    if lo <= a <= hi:
        print("%05x  %s" % (a, t))
