"""Golden vectors for the outlier filter and FPS rows, from the REAL reference code.

    python oracle/make_golden_post.py

* outlier filter: the unmodified reference Generator3D6.generateiopoint is run with its default outlier_threshold = 1.5
  on a displaced point set that contains genuine outliers.  To keep the run short the two networks are replaced by
  stubs that return fixed normals/distances (the filter only sees the final points), seeds enter through target.xyz.
* FPS: generate.py's farthest_point_sample is executed from its own source text with the hard-coded device string
  'cuda' replaced by 'cpu' (nothing else touched).
Both are asserted equal to the oracle restatements and stored in tests/golden/post.npz.
"""
import os
import re
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    import sapcu_oracle as orc
    for m in ("trimesh", "h5py"):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.path.insert(0, REF)
    import generation as rgen

    rng = np.random.default_rng(21)
    # ---------------- outlier filter through the reference pipeline
    n_seed = 600
    cloud = rng.normal(size=(400, 3)); cloud = 0.5 * cloud / np.linalg.norm(cloud, axis=1, keepdims=True)
    seeds = cloud[rng.integers(0, 400, n_seed)] * 1.02
    normals = rng.normal(size=(n_seed, 3)).astype(np.float32)
    dists = rng.uniform(0.0, 0.01, n_seed).astype(np.float32)
    dists[::37] += 0.2                                   # genuine outliers

    class StubFn(torch.nn.Module):
        def forward(self, x): return torch.from_numpy(normals[: x.shape[0]])
    class StubFd(torch.nn.Module):
        def forward(self, x): return torch.from_numpy(dists[: x.shape[0]])
    gen = rgen.Generator3D6(StubFn(), StubFd(), torch.device("cpu"), k_neighbors=16, batch_size=10 ** 9)
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            open("dense", "w").write("#!/bin/sh\nexit 0\n"); os.chmod("dense", 0o755)
            np.savetxt("target.xyz", seeds, fmt="%.18e")
            kept_ref = np.asarray(gen.upsample(np.expand_dims(cloud, 0)))
            seeds_rt = np.loadtxt("target.xyz")
        finally:
            os.chdir(cwd)
    un = torch.nn.functional.normalize(torch.from_numpy(normals), dim=-1).numpy()
    pts = orc.displace(seeds_rt, un, dists)
    keep = orc.outlier_filter(pts, 1.5)
    assert np.array_equal(pts[keep], kept_ref), "oracle outlier filter != reference"
    print("outlier filter: kept", len(keep), "of", n_seed)

    # ---------------- FPS from the reference source text
    src = open(os.path.join(REF, "generate.py")).read()
    body = re.search(r"def farthest_point_sample\(.*?\n(?=\n\n# =)", src, re.S).group(0)
    assert "device = 'cuda'" in body
    ns = {"torch": torch, "np": np}
    exec(body.replace("device = 'cuda'", "device = 'cpu'"), ns)
    xyz = rng.normal(size=(5000, 3)) * np.array([1.0, 0.6, 0.3])
    idx_ref = ns["farthest_point_sample"](xyz, 512)
    assert np.array_equal(orc.fps(xyz, 512), idx_ref), "oracle fps != reference"
    grid = np.stack(np.meshgrid(np.arange(12), np.arange(12), np.arange(12), indexing="ij"), -1).reshape(-1, 3).astype(np.float64)
    idx_grid = ns["farthest_point_sample"](grid, 200)          # many exact ties
    assert np.array_equal(orc.fps(grid, 200), idx_grid)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "post.npz"), out_points=pts, out_keep=keep.astype(np.int32),
                        fps_xyz=xyz, fps_idx=idx_ref.astype(np.int32), fps_grid_idx=idx_grid.astype(np.int32))
    print("written tests/golden/post.npz")


if __name__ == "__main__":
    main()
