"""Checkpoint loading shared by the fn and fd shims.

The reference stores `{'model': state_dict, 'optimizer': ..., <scalars>}` with torch.save and restores every module
registered under a keyword (fn/checkpoints.py, fd/checkpoints.py); DataParallel checkpoints carry a `module.` key
prefix that is dropped on load.  Only that contract is reproduced here (construct, load, save, register)."""
import os

import torch


def _strip_data_parallel(sd):
    if any(k.startswith("module.") for k in sd):
        return {k[len("module."):] if k.startswith("module.") else k: v for k, v in sd.items()}
    return sd


def make_checkpoint_io(missing_exc):
    """Build a CheckpointIO class; `missing_exc` is what `load` raises for an absent file (the two reference
    modules differ: fn raises FileExistsError, fd FileNotFoundError)."""

    class CheckpointIO(object):
        def __init__(self, checkpoint_dir="./chkpts", **modules):
            self.module_dict = dict(modules)
            self.checkpoint_dir = checkpoint_dir
            os.makedirs(checkpoint_dir, exist_ok=True)

        def _path(self, name):
            return name if os.path.isabs(name) else os.path.join(self.checkpoint_dir, name)

        def register_modules(self, **modules):
            self.module_dict.update(modules)

        def save(self, filename, **scalars):
            payload = dict(scalars)
            payload.update({k: m.state_dict() for k, m in self.module_dict.items()})
            torch.save(payload, self._path(filename))

        def load(self, filename):
            path = self._path(filename)
            if not os.path.exists(path):
                raise missing_exc(path)
            return self.parse_state_dict(torch.load(path, map_location="cpu"))

        def parse_state_dict(self, blob):
            for key, module in self.module_dict.items():
                if key not in blob:
                    print("Warning: Could not find %s in checkpoint!" % key)
                    continue
                module.load_state_dict(_strip_data_parallel(blob[key]))
            return {k: v for k, v in blob.items() if k not in self.module_dict}

    return CheckpointIO
