"""Generate tests/golden/*.npz from the REAL reference (run in the build container, where /root/reference exists).

    python oracle/make_golden.py            # writes tests/golden/, asserts oracle == reference on every case

For every case the unmodified reference modules (imported from /root/reference with the two unused top-level
imports `trimesh`/`h5py` stubbed) are run on seeded inputs with the seeded weights of
sapcu_b200.synthetic.init_weights; the oracle restatement must reproduce them, and inputs + reference outputs are
stored so tests can re-check the oracle (and the CUDA path) anywhere.  The fixtures hold no weights: they are
regenerated from the seed.
"""
import os
import subprocess
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("SAPCU_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLD = os.path.join(ROOT, "tests", "golden")


def load_reference():
    for m in ("trimesh", "h5py"):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.path.insert(0, REF)
    import fn.config as rfn      # noqa
    import fd.config as rfd      # noqa
    import generation as rgen    # noqa
    return rfn, rfd, rgen


def main():
    import sapcu_b200.synthetic as syn
    import sapcu_oracle as orc
    rfn, rfd, rgen = load_reference()
    import fn.snn_coder as rfn_coder
    import fd.snn_coder as rfd_coder
    os.makedirs(GOLD, exist_ok=True)
    torch.set_grad_enabled(False)
    cfg_fn = rfn.load_config(os.path.join(REF, "config/fn.yaml"))
    cfg_fd = rfd.load_config(os.path.join(REF, "config/fd.yaml"))
    mfn = rfn.get_model(cfg_fn, torch.device("cpu")).eval()
    mfd = rfd.get_model(cfg_fd, torch.device("cpu")).eval()

    def clear_knn(m):
        for b in (m.encoder.trans1, m.encoder.trans2, m.encoder.trans3):
            b.knn_cache.cache.clear()

    # ---- state_dict key/shape inventory (the drop-in boundary)
    inv = {"fn": {k: list(v.shape) for k, v in mfn.state_dict().items()},
           "fd": {k: list(v.shape) for k, v in mfd.state_dict().items()}}
    import json
    with open(os.path.join(GOLD, "state_dict_keys.json"), "w") as f:
        json.dump(inv, f, indent=0, sort_keys=True)

    cloud = syn.cloud(2048, seed=0, shape="sphere")
    seeds = syn.seeds(cloud, 4, seed=1)

    # ---- K1: KDTree indices
    from sklearn.neighbors import KDTree
    S_knn = 512
    tree = KDTree(cloud)
    _, idx = tree.query(seeds[:S_knn], 100)
    oidx = orc.knn_seed(cloud, seeds[:S_knn], 100)
    assert (oidx == idx).all(), "oracle kNN != KDTree"
    cloud_b = syn.cloud(1500, seed=3, shape="boxes")
    seeds_b = syn.seeds(cloud_b, 0.2, seed=4)
    _, idx_b = KDTree(cloud_b).query(seeds_b, 48)
    assert (orc.knn_seed(cloud_b, seeds_b, 48) == idx_b).all()
    np.savez_compressed(os.path.join(GOLD, "knn.npz"), idx_sphere=idx.astype(np.int32), idx_boxes=idx_b.astype(np.int32))

    # ---- neuron known-answer vectors (LIF + EIF, T steps fed back / external input)
    g = torch.Generator().manual_seed(7)
    C, R, T = 16, 48, 7
    x = torch.empty(R, C).normal_(0.3, 1.5, generator=g)
    x[0, :] = torch.linspace(-12, 12, C)            # exercise the +-10 clamp
    lif = rfd_coder.MultiTimeConstantLIFNeuron(C).eval()
    eif = rfd_coder.MultiTimeConstantEIFNeuron(C).eval()
    prm = {}
    for name, lo, hi in (("membrane_decay", 0.05, 1.05), ("threshold_adapt", 0.0, 0.12), ("refractory_decay", 0.05, 1.0),
                         ("threshold_base", 0.5, 1.5), ("delta_T", 0.05, 5.5), ("theta_rh", 0.05, 2.2)):
        prm[name] = torch.empty(C).uniform_(lo, hi, generator=g)      # deliberately beyond the clamp ranges
    for mod in (lif, eif):
        for k in prm:
            if hasattr(mod, k):
                getattr(mod, k).data.copy_(prm[k])
    outs = {}
    for tag, mod in (("lif", lif), ("eif", eif)):
        st, s, seq = [None, None, None], x.clone(), []
        for t in range(T):
            s, *st = mod(s, *st)
            seq.append(s.clone())
        outs[tag] = torch.stack(seq, 0)
        p = {k: v for k, v in prm.items() if tag == "eif" or k not in ("delta_T", "theta_rh")}
        mine = orc.lif_chain(x.clone(), p, T, all_steps=True)
        assert torch.equal(mine, outs[tag]), tag
    np.savez_compressed(os.path.join(GOLD, "neuron.npz"), x=x.numpy(), lif=outs["lif"].numpy(), eif=outs["eif"].numpy(),
                        **{"p_" + k: v.numpy() for k, v in prm.items()})

    # ---- Rodrigues + gather/centre/rotate + displacement
    rng = np.random.default_rng(11)
    nrm = rng.normal(size=(40, 3)).astype(np.float32)
    nrm[0] = [1, 0, 0]; nrm[1] = [-1, 0, 0]; nrm[2] = [0, 1, 0]; nrm[3] = [3e-4, 1, -2]   # noqa: E702
    Rm = np.stack([rgen.rotation_matrix_from_vectors(n, [1, 0, 0]) for n in nrm])
    assert all(np.array_equal(orc.rotation_to_x(n), r) for n, r in zip(nrm, Rm))
    S_g = 40
    pi = idx[:S_g]
    patch = cloud[pi] - np.tile(np.expand_dims(seeds[:S_g], 1), (1, 100, 1))
    rot = patch.copy()
    for j in range(S_g):
        rot[j] = (np.matmul(Rm[j], rot[j].T)).T
    dd = rng.uniform(0, 0.03, size=S_g).astype(np.float32)
    nn_ = (nrm / np.linalg.norm(nrm, axis=1, keepdims=True)).astype(np.float32)
    disp = seeds[:S_g] + nn_ * np.tile(np.expand_dims(dd, 1), (1, 3))
    assert np.array_equal(orc.gather_center(cloud, seeds[:S_g], pi, nrm), rot.astype(np.float32))
    assert np.array_equal(orc.displace(seeds[:S_g], nn_, dd), disp)
    np.savez_compressed(os.path.join(GOLD, "patch_ops.npz"), normals=nrm, R=Rm, patch=patch.astype(np.float32),
                        rotated=rot.astype(np.float32), unit_normals=nn_, dist=dd, displaced=disp)

    # ---- model forwards: default ("random") init and stress init
    B = 3
    p_fn = torch.from_numpy(orc.gather_center(cloud, seeds[:B], idx[:B]))
    store = {"patches": p_fn.numpy()}
    for tag, stress in (("default", False), ("stress", True)):
        syn.init_weights(mfn, seed=100, stress=stress)
        syn.init_weights(mfd, seed=200, stress=stress)
        caps = {}
        hooks = []
        for name, mod in list(mfn.named_modules()) + list(mfd.named_modules()):
            if "Neuron" in type(mod).__name__:
                key = ("fn." if mod in set(mfn.modules()) else "fd.") + name
                hooks.append(mod.register_forward_hook(lambda m, i, o, n=key: caps.setdefault(n, []).append(o[0].clone())))
        clear_knn(mfn)
        n_ref = mfn(p_fn)
        n_unit = torch.nn.functional.normalize(n_ref, dim=-1)
        p_fd = torch.from_numpy(orc.gather_center(cloud, seeds[:B], idx[:B], n_unit.numpy()))
        d_ref = mfd(p_fd)
        for h in hooks:
            h.remove()
        sd_fn = {k: v.clone() for k, v in mfn.state_dict().items()}
        sd_fd = {k: v.clone() for k, v in mfd.state_dict().items()}
        taps_fn, taps_fd = {}, {}
        n_or = orc.fn_forward(sd_fn, p_fn, taps=taps_fn)
        d_or = orc.fd_forward(sd_fd, p_fd, taps=taps_fd)
        d_dce = orc.fd_forward(sd_fd, p_fd, schedule="dce")
        print(tag, "fn max|oracle-ref| =", (n_or - n_ref).abs().max().item(), " fd =", (d_or - d_ref).abs().max().item(),
              " fd dce-faithful =", (d_dce - d_or).abs().max().item())
        assert torch.equal(n_or, n_ref) and torch.equal(d_or, d_ref) and torch.equal(d_dce, d_or)
        # spike-train pins: last step of a few fn layers, every step of the fd blocks (sub-sampled channels)
        assert torch.equal(taps_fn["snn_init"], caps["fn.encoder.snn_init"][-1])
        assert torch.equal(taps_fn["snn_final"], caps["fn.encoder.snn_final"][-1])
        assert torch.equal(taps_fn["encoder.trans3"]["a1"], caps["fn.encoder.trans3.snn_gamma"][-1])
        fd_sp = torch.stack([torch.cat([caps["fd.encoder.snn_blocks.%d" % b][t] for b in range(4)], 1) for t in range(7)], 0)
        assert torch.equal(taps_fd["spikes"], fd_sp)
        hard = {k: float((torch.stack(v) > 0.449471).float().mean()) for k, v in caps.items()}
        print("   hard-spike fraction min/max:", min(hard.values()), max(hard.values()))
        store.update({
            tag + "_normals": n_ref.numpy(), tag + "_patches_fd": p_fd.numpy(), tag + "_dist": d_ref.numpy(),
            tag + "_fn_snn_init": taps_fn["snn_init"].numpy()[:, ::4], tag + "_fn_snn_final": taps_fn["snn_final"].numpy()[:, ::16],
            tag + "_fn_fcat": taps_fn["fcat"].numpy(), tag + "_fn_gmax": taps_fn["gmax"].numpy(),
            tag + "_fd_spikes": fd_sp.numpy()[:, :, ::16], tag + "_fd_pool": taps_fd["pool"].numpy(), tag + "_fd_z": taps_fd["z"].numpy(),
        })
    np.savez_compressed(os.path.join(GOLD, "models.npz"), **store)

    # ---- the unmodified reference pipeline (Generator3D6.generateiopoint) with injected seeds
    S_p = 32      # >= 30: the reference's outlier step queries 30 neighbours
    syn.init_weights(mfn, seed=100, stress=True)
    syn.init_weights(mfd, seed=200, stress=True)
    clear_knn(mfn)
    gen = rgen.Generator3D6(mfn, mfd, torch.device("cpu"), k_neighbors=100, outlier_threshold=1e9, batch_size=400)
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            with open("dense", "w") as f:
                f.write("#!/bin/sh\nexit 0\n")      # seeds are injected through target.xyz; ./dense becomes a no-op
            os.chmod("dense", 0o755)
            np.savetxt("target.xyz", seeds[:S_p], fmt="%.18e")
            pts_ref = np.asarray(gen.upsample(np.expand_dims(cloud, 0)))
        finally:
            os.chdir(cwd)
    sd_fn = {k: v.clone() for k, v in mfn.state_dict().items()}
    sd_fd = {k: v.clone() for k, v in mfd.state_dict().items()}
    seeds_rt = np.loadtxt(__import__("io").StringIO("\n".join(" ".join("%.18e" % v for v in r) for r in seeds[:S_p])))
    pts_or, _, n_or, d_or = orc.pipeline(sd_fn, sd_fd, cloud, seeds_rt, K=100, batch=400)
    print("pipeline max|oracle-ref| =", np.abs(pts_or - pts_ref).max())
    assert np.array_equal(pts_or, pts_ref)
    np.savez_compressed(os.path.join(GOLD, "pipeline.npz"), seeds=seeds_rt, points=pts_ref, normals=n_or, dist=d_or)
    print("golden fixtures written to", GOLD)
    for f in sorted(os.listdir(GOLD)):
        print("  %-24s %8d bytes" % (f, os.path.getsize(os.path.join(GOLD, f))))


if __name__ == "__main__":
    main()
