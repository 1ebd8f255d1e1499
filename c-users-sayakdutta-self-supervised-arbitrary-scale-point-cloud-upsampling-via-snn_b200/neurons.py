"""Parameter holders for the two spiking neuron types on the hot path.

These mirror the *names and shapes* of the reference neuron modules so that
reference checkpoints load unchanged (state_dict keys):

  LIF: membrane_decay, threshold_adapt, refractory_decay, threshold_base   [C]
       (reference: fn/snn_coder.py:63-85, fd/snn_coder.py:70-92)
  EIF: the four above + delta_T, theta_rh                                  [C]
       (reference: fd/snn_coder.py:158-196)

They carry no arithmetic: the membrane recurrence itself runs inside the CUDA
library (csrc/neuron.cuh), and its CPU restatement lives in oracle/.
"""
import torch
import torch.nn as nn


class MultiTimeConstantLIFNeuron(nn.Module):
    def __init__(self, layer_size, membrane_decay_init=0.9, threshold_adapt_init=0.01,
                 refractory_decay_init=0.5, grad_width=10.0):
        super().__init__()
        self.layer_size = layer_size
        self.grad_width = grad_width
        self.membrane_decay = nn.Parameter(torch.full((layer_size,), float(membrane_decay_init)))
        self.threshold_adapt = nn.Parameter(torch.full((layer_size,), float(threshold_adapt_init)))
        self.refractory_decay = nn.Parameter(torch.full((layer_size,), float(refractory_decay_init)))
        self.threshold_base = nn.Parameter(torch.ones(layer_size))

    def forward(self, *a, **k):  # pragma: no cover - never called on the product path
        raise RuntimeError("neuron recurrences run inside the sapcu_b200 CUDA library, not in Python")


class MultiTimeConstantEIFNeuron(MultiTimeConstantLIFNeuron):
    def __init__(self, layer_size, delta_T_init=1.0, theta_rh_init=0.8, **kw):
        super().__init__(layer_size, **kw)
        self.delta_T = nn.Parameter(torch.full((layer_size,), float(delta_T_init)))
        self.theta_rh = nn.Parameter(torch.full((layer_size,), float(theta_rh_init)))
