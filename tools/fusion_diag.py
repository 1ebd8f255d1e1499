"""Diagnostic: fn normals in MODE_TC with the current env switches vs the fp32 oracle, on B stress-init patches.
    SAPCU_TC_FACTOR_ATTNIN=0 python tools/fusion_diag.py 48 out.npz
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import sapcu_b200  # noqa: E402
import sapcu_b200.synthetic as syn  # noqa: E402
from sapcu_b200.fn import config as fc  # noqa: E402
import sapcu_oracle as orc  # noqa: E402
import oracle_c  # noqa: E402


def angle(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.degrees(np.arctan2(np.linalg.norm(np.cross(a, b), axis=-1), (a * b).sum(-1)))


def main():
    B = int(sys.argv[1])
    cloud = syn.cloud(2048, seed=0, shape="sphere")
    seeds = syn.seeds(cloud, 4, seed=1)
    idx = oracle_c.knn(cloud, seeds[:B], 100)
    p = torch.from_numpy(orc.gather_center(cloud, seeds[:B], idx))
    mfn = fc.get_model(fc.load_config(os.path.join(sapcu_b200.CONFIG_DIR, "fn.yaml")))
    syn.init_weights(mfn, seed=100, stress=True)
    sd = {k: v.clone() for k, v in mfn.state_dict().items()}
    ref_path = sys.argv[2]
    if os.path.exists(ref_path):
        ref = np.load(ref_path)["ref"]
    else:
        with torch.no_grad():
            ref = orc.fn_forward(sd, p).numpy()
        np.savez(ref_path, ref=ref)
    mfn = mfn.to("cuda:0")
    out = {}
    for mode in ("fp32", "tc"):
        mfn.set_mode(mode)
        got = mfn(p.to("cuda:0")).cpu().numpy()
        ang = angle(got, ref)
        out[mode] = got
        print(mode, "vs oracle: max %.3e deg, mean %.3e deg" % (ang.max(), ang.mean()))
    print("tc vs fp32 mode: max %.3e deg" % angle(out["fp32"], out["tc"]).max())


if __name__ == "__main__":
    main()
