"""Shared machinery of the two model shims: a torch nn.Module that owns the reference-shaped parameters
(so state_dict()/load_state_dict()/CheckpointIO keep working) and forwards through the C ABI.

torch is plumbing here: parameter containers, device memory for inputs/outputs/workspace, the stream.
"""
import ctypes

import torch
import torch.nn as nn

from . import _native as N


class NativeModel(nn.Module):
    KIND = None            # N.MODEL_FN / N.MODEL_FD
    #: bytes of workspace the shim is willing to allocate per forward (chunks internally below that)
    WORKSPACE_CAP = int(float(__import__('os').environ.get('SAPCU_WS_CAP_GB', '24')) * (1 << 30))

    def __init__(self):
        super().__init__()
        self._handle = None
        self._handle_version = None
        self._ws = None
        self.mode = N.MODE_FP32

    # ------------------------------------------------------------------ handle management
    def _cfg_ints(self):
        raise NotImplementedError

    def _state_version(self):
        # parameters/buffers are re-uploaded when any of them was modified in place or replaced
        return tuple((k, v.data_ptr(), v._version) for k, v in self.state_dict(keep_vars=True).items())

    def _ensure_handle(self, device=None):
        """The native handle for the module's current parameters on `device` (the weights are uploaded to the CUDA device
        that is current at finalize time, so the device is part of the version key; callers hold torch.cuda.device(device))."""
        dev_index = torch.device(device).index if device is not None else (torch.cuda.current_device() if torch.cuda.is_available() else -1)
        ver = (dev_index,) + self._state_version()
        if self._handle is not None and ver == self._handle_version:
            return self._handle
        self._drop_handle()
        L = N.lib()
        cfg = (ctypes.c_int32 * len(self._cfg_ints()))(*self._cfg_ints())
        h = L.sapcu_model_create(self.KIND, cfg, len(cfg))
        if not h:
            raise N.SapcuError("sapcu_model_create: " + L.sapcu_last_error().decode())
        try:
            for name, t in self.state_dict().items():
                if not t.dtype.is_floating_point:
                    continue   # num_batches_tracked
                host = t.detach().to("cpu", torch.float32).contiguous()
                N.check(L.sapcu_model_set_tensor(h, name.encode(), N.ptr(host), host.numel()), "set_tensor(%s)" % name)
            N.check(L.sapcu_model_finalize(h), "sapcu_model_finalize")
        except Exception:
            L.sapcu_model_destroy(h)
            raise
        self._handle, self._handle_version = h, ver
        return h

    def _drop_handle(self):
        if self._handle is not None:
            N.lib().sapcu_model_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._drop_handle()
        except Exception:
            pass

    # ------------------------------------------------------------------ reference API surface
    def reset_states(self):
        """No-op: the reference's reset_states() has no effect on eval outputs."""
        return None

    def train(self, mode=True):
        if mode:
            raise N.SapcuError("sapcu_b200 implements the inference hot path only (eval mode)")
        return super().train(False)

    def set_mode(self, mode):
        """'fp32' (FFMA, reference-faithful arithmetic), 'tc' (tcgen05 split products, parity-grade), 'tf32' (tcgen05
        single-pass TF32) or 'fast' (single fp16 product per MAC on fp16 spike tensors + cheap LIF chains; the fast mode
        whose deviation is reported separately)."""
        self.mode = {"fp32": N.MODE_FP32, "tc": N.MODE_TC, "tf32": N.MODE_TF32, "fast": N.MODE_FAST}[mode] if isinstance(mode, str) else int(mode)
        return self

    # ------------------------------------------------------------------ helpers for subclasses
    def _workspace(self, S, M, device):
        L = N.lib()
        h = self._ensure_handle(device)
        need = L.sapcu_model_workspace_bytes(h, S, M)
        one = L.sapcu_model_workspace_bytes(h, 1, M)
        want = max(one, min(need, self.WORKSPACE_CAP))
        if self._ws is None or self._ws.numel() < want or self._ws.device != device:
            self._ws = None
            self._ws = torch.empty(want, dtype=torch.uint8, device=device)
        return self._ws

    def _prep_input(self, x):
        if not x.is_cuda:
            raise N.SapcuError("sapcu_b200 models run on CUDA tensors only (no CPU fallback); got a %s tensor" % x.device)
        return x.detach().to(torch.float32).contiguous()

    def tap(self, name, S, M, dtype=torch.float32):
        """Debug view of a named intermediate of the last forward (tests only)."""
        off, rows, cols, ld = (ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64())
        N.check(N.lib().sapcu_model_tap(self._ensure_handle(), name.encode(), S, M, self.mode, ctypes.byref(off), ctypes.byref(rows),
                                        ctypes.byref(cols), ctypes.byref(ld)), "model_tap(%s)" % name)
        fmt = N.lib().sapcu_model_tap_format(self._ensure_handle(), name.encode())
        if fmt == 1:      # fp16 (hi, lo) planes of x * 2^13 (see include/sapcu_b200.h)
            n = rows.value * cols.value
            h = self._ws.view(torch.float16)[2 * off.value: 2 * off.value + 2 * n].view(2, rows.value, cols.value)
            return (h[0].float() + h[1].float()) / 8192.0
        if fmt == 2:      # fast mode: ONE fp16 plane of x * 2^13
            n = rows.value * cols.value
            h = self._ws.view(torch.float16)[2 * off.value: 2 * off.value + n].view(rows.value, cols.value)
            return h.float() / 8192.0
        flat = self._ws.view(torch.float32) if dtype == torch.float32 else self._ws.view(torch.int32)
        return flat[off.value: off.value + rows.value * ld.value].view(rows.value, ld.value)[:, :cols.value]
