"""Golden vectors for the seed generator ("next" row 1) from the REAL reference binary.

    make -C oracle ref && python oracle/make_golden_seeds.py

Runs oracle/_ref/dense (g++ -O2 of /root/reference/dense.cpp, built where it lies, never copied) on seeded
synthetic clouds written to test.xyz with full precision, and stores the emitted seeds (in file order) as int32
micro-units (the file holds 6 decimals) in tests/golden/seeds.npz.  The big configuration-1 case (N=2048,
cell=0.004, ~390k seeds) is stored as count + SHA-256 of the ordered micro-unit array + its first/last 512 rows.
"""
import hashlib
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
DENSE = os.path.join(ROOT, "oracle", "_ref", "dense")


def run_dense(cloud, cell):
    with tempfile.TemporaryDirectory() as tmp:
        np.savetxt(os.path.join(tmp, "test.xyz"), cloud, fmt="%.17g")
        subprocess.check_call([DENSE, repr(cell), str(cloud.shape[0])], cwd=tmp)
        out = np.loadtxt(os.path.join(tmp, "target.xyz")).reshape(-1, 3)
    return out


def micro(a):
    return np.rint(a * 1e6).astype(np.int32)


def main():
    import sapcu_b200.synthetic as syn
    cases = {
        "sphere256_c010": (syn.cloud(256, seed=5, shape="sphere"), 0.01),
        "boxes2048_c008": (syn.cloud(2048, seed=3, shape="boxes"), 0.008),
        "sphere2048_c004": (syn.cloud(2048, seed=0, shape="sphere"), 0.004),
    }
    store = {}
    for name, (cloud, cell) in cases.items():
        seeds = run_dense(cloud, cell)
        m = micro(seeds)
        assert np.abs(m / 1e6 - seeds).max() < 1e-12
        print(name, seeds.shape[0], "seeds")
        store[name + "_count"] = np.int64(seeds.shape[0])
        store[name + "_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(m).tobytes()).digest(), dtype=np.uint8)
        if seeds.shape[0] <= 60000:
            store[name] = m
        else:
            store[name + "_head"] = m[:512]
            store[name + "_tail"] = m[-512:]
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "seeds.npz"), **store)
    print("written", os.path.getsize(os.path.join(ROOT, "tests", "golden", "seeds.npz")), "bytes")


if __name__ == "__main__":
    main()
