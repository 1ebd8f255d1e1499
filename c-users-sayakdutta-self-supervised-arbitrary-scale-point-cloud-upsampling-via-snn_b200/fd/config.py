"""Construction boundary of the fd model: `load_config` + `get_model` with the reference's signatures
(reference: fd/config.py:6-30, :89-155).  Keys consumed: k, emb_dims, time_steps_enc, time_steps_dec,
num_heads, dropout, use_snn_decoder, k_scales, type.  The reference's config/fd.yaml loads unchanged.
"""
import os

import yaml

from .snn_coder import EnhancedSNNDistanceEstimation


def _merge(dst, src):
    for key, val in src.items():
        if isinstance(val, dict):
            dst[key] = _merge(dst.get(key, {}) if isinstance(dst.get(key), dict) else {}, val)
        else:
            dst[key] = val
    return dst


def load_config(path, default_path=None):
    with open(path, "r") as f:
        special = yaml.safe_load(f) or {}
    parent = special.get("inherit_from")
    if parent is not None:
        cfg = load_config(parent, default_path)
    elif default_path is not None:
        with open(default_path, "r") as f:
            cfg = yaml.safe_load(f) or {}
    else:
        cfg = {}
    return _merge(cfg, special)


def get_model(cfg, device):
    m = cfg.get("model", {})
    if m.get("type", "enhanced") != "enhanced":
        raise ValueError("only model.type == 'enhanced' (EnhancedSNNDistanceEstimation) is on the hot path")
    model = EnhancedSNNDistanceEstimation(
        k=m.get("k", 20), emb_dims=m.get("emb_dims", 512), time_steps_enc=m.get("time_steps_enc", 5),
        time_steps_dec=m.get("time_steps_dec", 8), num_heads=m.get("num_heads", 4), dropout=m.get("dropout", 0.1),
        use_snn_decoder=m.get("use_snn_decoder", False), k_scales=m.get("k_scales", [10, 20, 40]))
    return model.to(device) if device is not None else model
