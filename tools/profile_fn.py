"""Small driver for ncu captures: one fn (+ optionally fd) forward of S synthetic patches in a given mode.
    python tools/profile_fn.py --S 1024 --mode tc [--fd]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--S", type=int, default=1024)
    ap.add_argument("--mode", default="tc")
    ap.add_argument("--fd", action="store_true")
    ap.add_argument("--reps", type=int, default=1)
    a = ap.parse_args()
    import bench
    mfn, mfd, _, _ = bench.build_models(torch.device("cuda:0"))
    mfn.set_mode(a.mode), mfd.set_mode(a.mode)
    g = torch.Generator().manual_seed(0)
    p = (torch.randn(a.S, 100, 3, generator=g) * 0.03).to("cuda:0")
    for _ in range(a.reps):
        out = mfd(p) if a.fd else mfn(p)
    torch.cuda.synchronize()
    print("ok", tuple(out.shape), float(out.abs().mean()))


if __name__ == "__main__":
    main()
