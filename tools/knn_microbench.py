"""Stand-alone microbenchmark of K1 (seed -> cloud kNN), BASELINE.json configs[4] "standalone kNN microbench at K=48".
    python tools/knn_microbench.py [--N 2000000] [--S 65536] [--K 48]
Reports pair evaluations per second against the FP32-issue bound (8 lane-ops per pair, 148 SM x 128 lanes x clock)."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=2000000)
    ap.add_argument("--S", type=int, default=65536)
    ap.add_argument("--K", type=int, default=48)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--check", type=int, default=64, help="seeds verified against the fp64 oracle")
    ap.add_argument("--sorted", action="store_true", help="Morton-order the seeds first, as Generator3D6 does for clouds of >= 2^20 points")
    a = ap.parse_args()
    import sapcu_b200, sapcu_b200.synthetic as syn
    from sapcu_b200 import _native as N
    import oracle_c
    L = sapcu_b200.lib()
    cloud = syn.cloud(a.N, seed=0, shape="sphere")
    seeds = syn.seeds(cloud, a.S / a.N, seed=1)[: a.S]
    dc, ds = torch.from_numpy(cloud).cuda(), torch.from_numpy(seeds).cuda()
    if a.sorted:
        from sapcu_b200.generation import morton_order
        perm = morton_order(ds)
        ds = ds[perm].contiguous()
        seeds = seeds[perm.cpu().numpy()]
    idx = torch.empty(a.S, a.K, dtype=torch.int32, device="cuda")
    ws = torch.empty(L.sapcu_knn_workspace_bytes(a.N), dtype=torch.uint8, device="cuda")
    def run():
        N.check(L.sapcu_knn(N.ptr(dc), a.N, N.ptr(ds), a.S, a.K, N.ptr(idx), N.ptr(ws), ws.numel(), None))
    run(); torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.reps)]
    for s, e in ev:
        s.record(); run(); e.record()
    torch.cuda.synchronize()
    ms = min(s.elapsed_time(e) for s, e in ev)
    ok = bool(np.array_equal(idx[: a.check].cpu().numpy(), oracle_c.knn(cloud, seeds[: a.check], a.K)))
    pairs = a.N * a.S / (ms / 1e3)
    bound = 148 * 128 * 1.965e9 / 8
    print(json.dumps({"N": a.N, "S": a.S, "K": a.K, "ms": ms, "pairs_per_s": pairs, "fp32_issue_bound_pairs_per_s": bound,
                      "frac_of_bound": pairs / bound, "seeds_per_s": a.S / (ms / 1e3), "bit_exact_vs_fp64_oracle": ok,
                      "seed_order": "morton" if a.sorted else "as generated (random over the surface)"}))


if __name__ == "__main__":
    main()
