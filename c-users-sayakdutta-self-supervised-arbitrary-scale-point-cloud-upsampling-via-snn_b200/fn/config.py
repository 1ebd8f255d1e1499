"""Construction boundary of the fn model: `load_config` + `get_model` with the reference's signatures
(reference: fn/config.py:10-40 yaml loading with `inherit_from`, :183-231 model factory).

Only the keys `get_model` consumes matter on the hot path: k_values, emb_dims, time_steps_enc,
time_steps_dec, num_heads, use_snn_decoder, decoder_dropout.  The reference's own config/fn.yaml loads
unchanged (every other section is carried along untouched).
"""
import os

import yaml

from .snn_coder import ImprovedSNNNormalEstimation

_MODEL_DEFAULTS = {            # reference fn/config.py:58-73 (model part of set_default_config_values)
    "k_values": [20, 20, 16], "emb_dims": 1024, "time_steps_enc": 8, "time_steps_dec": 12, "num_heads": 4,
    "dropout": 0.1, "use_snn_decoder": False, "decoder_dropout": 0.1,
}


def _merge(dst, src):
    for key, val in src.items():
        if isinstance(val, dict):
            node = dst.get(key)
            if not isinstance(node, dict):
                node = dst[key] = {}
            _merge(node, val)
        else:
            dst[key] = val
    return dst


def load_config(path, default_path=None):
    if not os.path.exists(path):
        raise FileNotFoundError(f"Config file not found: {path}")
    with open(path, "r") as f:
        special = yaml.safe_load(f) or {}
    parent = special.get("inherit_from")
    if parent is not None:
        if not os.path.isabs(parent):
            parent = os.path.join(os.path.dirname(path), parent)
        cfg = load_config(parent, default_path)
    elif default_path is not None:
        with open(default_path, "r") as f:
            cfg = yaml.safe_load(f) or {}
    else:
        cfg = {}
    _merge(cfg, special)
    model = cfg.setdefault("model", {})
    for key, val in _MODEL_DEFAULTS.items():
        model.setdefault(key, val)
    cfg.setdefault("data", {})
    return cfg


def get_model(cfg, device=None):
    m = cfg["model"]
    for req in ("k_values", "emb_dims"):
        if req not in m:
            raise ValueError(f"Missing required model parameter: {req}")
    model = ImprovedSNNNormalEstimation(
        k_values=m["k_values"], emb_dims=m["emb_dims"], time_steps_enc=m["time_steps_enc"],
        time_steps_dec=m["time_steps_dec"], num_heads=m["num_heads"],
        use_snn_decoder=m.get("use_snn_decoder", False), decoder_dropout=m.get("decoder_dropout", 0.1))
    if device is not None:
        model = model.to(device)
    return model
