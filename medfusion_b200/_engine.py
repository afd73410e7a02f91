"""Shared host-side plumbing between the torch-facing modules and the C-ABI engines."""
from __future__ import annotations

import ctypes

import torch
import torch.nn as nn

from . import _lib
from ._params import build_param_tree, param_signature


def cuda_stream_ptr(device=None) -> int:
    """Raw handle of torch's current stream ON `device` (default: the current device)."""
    return torch.cuda.current_stream(device).cuda_stream


def on_device(t: torch.Tensor):
    """Context manager making `t`'s device current: the C side allocates (weights, stream-K scratch) and launches on the
    current device, so every engine call runs under the device of the tensors it was handed."""
    return torch.cuda.device(t.device)


def require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(
            f"{what}: medfusion_b200 runs on CUDA (sm_100a) only — got a {t.device} tensor. "
            "There is no CPU fallback; move the module and its inputs to a B200.")


class _Handle:
    """Owns one C engine object; destroyed when the owning module is garbage collected."""

    def __init__(self, ptr, destroy):
        self.ptr, self._destroy = ptr, destroy

    def __del__(self):
        try:
            if self.ptr and self._destroy is not None:
                self._destroy(self.ptr)
        except Exception:
            pass
        self.ptr = None


class EngineModule(nn.Module):
    """nn.Module whose parameters mirror a C engine's registry and whose compute is the engine's."""

    _prefix = ""  # "mf_unet" / "mf_vae"

    def _engine_init(self, handle, zero_init=()):
        lib = _lib.load()
        self._handle = _Handle(handle, getattr(lib, f"{self._prefix}_destroy"))
        self._h = handle
        self._ws = {}          # (B,H,W) -> uint8 workspace tensor
        self._synced_sig = None
        n = getattr(lib, f"{self._prefix}_param_count")(handle)
        entries = []
        shape = (ctypes.c_int64 * 4)()
        ndim = ctypes.c_int()
        for i in range(n):
            name = getattr(lib, f"{self._prefix}_param_name")(handle, i).decode()
            _lib.check(getattr(lib, f"{self._prefix}_param_shape")(handle, i, shape, ctypes.byref(ndim)), "param_shape")
            entries.append((name, tuple(int(shape[k]) for k in range(ndim.value))))
        self._entries = entries
        build_param_tree(self, entries, zero_init=zero_init)

    @property
    def device(self):
        return next(self.parameters()).device

    def sync_params(self, force=False):
        """Push parameters whose storage/version changed since the last push into the engine."""
        sig = param_signature(self)
        if not force and sig == self._synced_sig:
            return
        lib = _lib.load()
        sd = dict(self.named_parameters())
        setter = getattr(lib, f"{self._prefix}_set_param")
        dev = self.device
        with torch.cuda.device(dev):
            stream = cuda_stream_ptr(dev)
            for name, shape in self._entries:
                p = sd[name]
                require_cuda(p, f"parameter {name}")
                if p.device != dev:
                    raise RuntimeError(f"parameter {name} is on {p.device}, the module on {dev}: one engine per device")
                # the engine reads numel*4 bytes: anything but fp32 (e.g. after .half()) is converted, never reinterpreted
                data = p.detach().to(torch.float32).contiguous()
                arr = (ctypes.c_int64 * len(shape))(*shape)
                _lib.check(setter(self._h, name.encode(), data.data_ptr(), arr, len(shape), stream),
                           f"set_param({name})")
            self._after_param_sync(stream)
        self._synced_sig = sig

    def _after_param_sync(self, stream):
        pass

    _WS_KEEP = 3   # workspaces kept per module (alternating batch sizes, e.g. the ragged tail of a chunked job)

    def _workspace_tensor(self, B, H, W):
        """Caller-owned scratch for one (B,H,W) plan.  A few recent shapes are kept so that alternating batch sizes
        return to the SAME address (the engine then rebuilds an identical plan, and CUDA graphs captured against it
        stay valid); graph holders additionally keep their tensor alive themselves."""
        key = (B, H, W, self.device.index)
        ws = self._ws.pop(key, None)
        if ws is None:
            nbytes = getattr(_lib.load(), f"{self._prefix}_workspace_bytes")(self._h, B, H, W)
            if nbytes == 0:
                _lib.check(2, "workspace_bytes")
            while len(self._ws) >= self._WS_KEEP:
                self._ws.pop(next(iter(self._ws)))
            ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=self.device)
        self._ws[key] = ws   # most recently used last
        return ws

    def _workspace(self, B, H, W):
        ws = self._workspace_tensor(B, H, W)
        off = (-ws.data_ptr()) % 1024
        return ws.data_ptr() + off, ws.numel() - off

    def _profile_call(self, fn_name, head_args, tail_args, max_ops=4096):
        """Run one engine call with a CUDA-event pair around every launch -> list of (ms, kind, flops)."""
        ms = (ctypes.c_float * max_ops)()
        kinds = (ctypes.c_int * max_ops)()
        flops = (ctypes.c_double * max_ops)()
        n = ctypes.c_int()
        with torch.cuda.device(self.device):
            _lib.check(getattr(_lib.load(), fn_name)(self._h, *head_args, *tail_args, cuda_stream_ptr(self.device), ms,
                                                     kinds, flops, max_ops, ctypes.byref(n)), fn_name)
        return [(ms[i], kinds[i], flops[i]) for i in range(n.value)]

    def plan_info(self):
        a, b, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        getattr(_lib.load(), f"{self._prefix}_plan_info")(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
        return dict(tc_convs=a.value, simt_convs=b.value, launches=c.value)
