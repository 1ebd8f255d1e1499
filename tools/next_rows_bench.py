"""Timing of the SURVEY.md section 8(f) rows built around the hot path, device vs the reference's host implementation:
seed generator (csrc/seedgen.cu vs the compiled reference ./dense), outlier filter (self-kNN k=30 + mask vs sklearn
KDTree), farthest-point sampling (cooperative kernel vs the reference loop restated in numpy).  Results are also
checked for equality, so every timing is of a parity-green call.
    python tools/next_rows_bench.py > profiles/r01_next_rows.json
"""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402
import torch  # noqa: E402


def gpu_time(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best, out


def main():
    import sapcu_b200
    import sapcu_b200.synthetic as syn
    from sapcu_b200.generation import Generator3D6
    from sapcu_b200.generate import farthest_point_sample
    import sapcu_oracle as orc
    sapcu_b200.lib()
    gen = Generator3D6(torch.nn.Identity(), torch.nn.Identity(), torch.device("cuda:0"))
    res = {"note": "wall-clock through the Python entry points (host numpy in, host numpy out), best of 3; "
                   "host numbers on the GPU box's cores"}

    # ---- seed generator: 1,000-point cloud at cell 0.01 (the size ./dense handles in seconds) and the bench cloud
    cloud = syn.cloud(1000, seed=5, shape="sphere")
    gen.dense_spacing = 0.01
    t_gpu, seeds = gpu_time(lambda: gen.gpu_seeds(cloud))
    entry = {"cloud_points": 1000, "cell": 0.01, "seeds": int(seeds.shape[0]), "gpu_s": t_gpu}
    dense = os.path.join(ROOT, "oracle", "_ref", "dense")
    if os.path.exists(dense):
        with tempfile.TemporaryDirectory() as td:
            np.savetxt(os.path.join(td, "test.xyz"), cloud, fmt="%.17g")
            t0 = time.perf_counter()
            subprocess.check_call([dense, "0.01", "1000"], cwd=td, stdout=subprocess.DEVNULL)
            entry["reference_binary_s"] = time.perf_counter() - t0
            ref = np.loadtxt(os.path.join(td, "target.xyz")).reshape(-1, 3)
            entry["identical"] = bool(ref.shape == seeds.shape and np.array_equal(np.rint(ref * 1e6), np.rint(seeds * 1e6)))
    res["seed_generator"] = entry
    # the reference's default cell (0.004) on the 2,048-point bench cloud; ./dense gets 300 s
    cloud2 = syn.cloud(2048, seed=0, shape="sphere")
    gen.dense_spacing = 0.004
    t_gpu, seeds2 = gpu_time(lambda: gen.gpu_seeds(cloud2))
    entry2 = {"cloud_points": 2048, "cell": 0.004, "seeds": int(seeds2.shape[0]), "gpu_s": t_gpu}
    if os.path.exists(dense):
        with tempfile.TemporaryDirectory() as td:
            np.savetxt(os.path.join(td, "test.xyz"), cloud2, fmt="%.17g")
            t0 = time.perf_counter()
            try:
                subprocess.check_call([dense, "0.004", "2048"], cwd=td, stdout=subprocess.DEVNULL, timeout=300)
                entry2["reference_binary_s"] = time.perf_counter() - t0
                ref = np.loadtxt(os.path.join(td, "target.xyz")).reshape(-1, 3)
                entry2["identical"] = bool(ref.shape == seeds2.shape and np.array_equal(np.rint(ref * 1e6), np.rint(seeds2 * 1e6)))
            except subprocess.TimeoutExpired:
                entry2["reference_binary_s"] = "> 300"
    res["seed_generator_default_cell"] = entry2

    # ---- outlier filter on 32,768 upsampled-like points
    pts = syn.cloud(32768, seed=7, shape="sphere") + np.random.default_rng(0).normal(scale=2e-3, size=(32768, 3))
    t_gpu, kept = gpu_time(lambda: gen._outlier_filter(pts))
    t0 = time.perf_counter()
    keep_ref = orc.outlier_filter(pts, gen.outlier_threshold)
    t_cpu = time.perf_counter() - t0
    same = bool(kept.shape[0] == keep_ref.shape[0] and np.array_equal(kept, pts[keep_ref]))
    res["outlier_filter"] = {"points": 32768, "k": 30, "gpu_s": t_gpu, "host_kdtree_s": t_cpu, "identical": same}

    # ---- FPS 32,768 -> 8,192
    x = pts.astype(np.float32)
    t_gpu, idx = gpu_time(lambda: farthest_point_sample(x, 8192, device="cuda:0"))
    t0 = time.perf_counter()
    idx_ref = orc.fps(x, 8192)
    t_cpu = time.perf_counter() - t0
    res["fps"] = {"points": 32768, "npoint": 8192, "gpu_s": t_gpu, "host_numpy_s": t_cpu,
                  "identical": bool(np.array_equal(idx, idx_ref))}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
