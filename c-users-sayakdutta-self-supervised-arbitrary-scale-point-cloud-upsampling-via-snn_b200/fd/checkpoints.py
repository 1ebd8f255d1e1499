"""CheckpointIO for the fd model: loads the reference's checkpoints ({'model': state_dict, ...}) into the
shim modules (reference: fd/checkpoints.py).  Only what generate.py uses: construct + load()."""
import os

import torch


class CheckpointIO(object):
    def __init__(self, checkpoint_dir='./chkpts', **kwargs):
        self.module_dict = kwargs
        self.checkpoint_dir = checkpoint_dir
        if not os.path.exists(checkpoint_dir):
            os.makedirs(checkpoint_dir)

    def register_modules(self, **kwargs):
        self.module_dict.update(kwargs)

    def save(self, filename, **kwargs):
        if not os.path.isabs(filename):
            filename = os.path.join(self.checkpoint_dir, filename)
        out = dict(kwargs)
        for k, v in self.module_dict.items():
            out[k] = v.state_dict()
        torch.save(out, filename)

    def load(self, filename):
        if not os.path.isabs(filename):
            filename = os.path.join(self.checkpoint_dir, filename)
        if not os.path.exists(filename):
            raise FileNotFoundError(filename)
        return self.parse_state_dict(torch.load(filename, map_location='cpu'))

    def parse_state_dict(self, state_dict):
        for k, v in self.module_dict.items():
            if k in state_dict:
                sd = state_dict[k]
                if any(key.startswith('module.') for key in sd.keys()):      # nn.DataParallel checkpoints
                    sd = {key.replace('module.', '', 1): val for key, val in sd.items()}
                v.load_state_dict(sd)
            else:
                print('Warning: Could not find %s in checkpoint!' % k)
        return {k: v for k, v in state_dict.items() if k not in self.module_dict}
