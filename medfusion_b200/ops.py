"""Kernel-level wrappers over the C ABI for tests/profiling (torch tensors in, torch tensors out).

Layouts: activations are NHWC, either one fp32 plane ("raw") or two fp16 planes [2, N, H, W, C]
("split": hi = fp16(x), lo = fp16(x - hi)).  torch only provides memory and streams here.
"""
from __future__ import annotations

import torch

from . import _lib

NCHW, NHWC_RAW, NHWC_SPLIT = 0, 1, 2


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def pack_split(x_nchw: torch.Tensor) -> torch.Tensor:
    x = x_nchw.contiguous().float()
    N, C, H, W = x.shape
    out = torch.empty((2, N, H, W, C), device=x.device, dtype=torch.float16)
    _lib.check(_lib.load().mf_op_pack_split(_p(x), _p(out), out[0].numel(), N, C, H, W, _stream()), "pack_split")
    return out


def unpack_nchw(t: torch.Tensor) -> torch.Tensor:
    """[2,N,H,W,C] split or [N,H,W,C] raw -> [N,C,H,W] fp32"""
    split = t.dim() == 5
    N, H, W, C = t.shape[-4:]
    out = torch.empty((N, C, H, W), device=t.device, dtype=torch.float32)
    _lib.check(_lib.load().mf_op_unpack_nchw(_p(t), t[0].numel() if split else 0, NHWC_SPLIT if split else NHWC_RAW,
                                             _p(out), N, C, H, W, _stream()), "unpack_nchw")
    return out


def prep_weight_tc(w: torch.Tensor):
    """-> (fp16 planes [2, Cout, K] pre-scaled by 2^S, scales float32[4] = {2^S, 2^-S, scratch})"""
    w = w.contiguous().float()
    Cout, Cin, kh, kw = w.shape
    out = torch.empty((2, Cout, kh * kw * Cin), device=w.device, dtype=torch.float16)
    scales = torch.zeros(4, device=w.device, dtype=torch.float32)
    _lib.check(_lib.load().mf_op_prep_weight_tc(_p(w), _p(out), _p(scales), Cout, Cin, kh, kw, _stream()),
               "prep_weight_tc")
    return out, scales


def prep_weight_simt(w: torch.Tensor) -> torch.Tensor:
    w = w.contiguous().float()
    Cout, Cin, kh, kw = w.shape
    out = torch.empty((kh * kw * Cin, Cout), device=w.device, dtype=torch.float32)
    _lib.check(_lib.load().mf_op_prep_weight_simt(_p(w), _p(out), Cout, Cin, kh, kw, _stream()), "prep_weight_simt")
    return out


def conv_tc_supported(N, H, W, C0, C1, Cout, ksize, stride=1) -> bool:
    return bool(_lib.load().mf_op_conv_tc_supported(N, H, W, C0, C1, Cout, ksize, stride))


def conv_tc(src0, w_planes, bias, ksize, src1=None, split_out=False, want_stats=False, drain_interval=0, stride=1):
    """src*: split tensors [2,N,H,W,C]; w_planes: the (planes, scales) pair of prep_weight_tc; returns (out, stats)."""
    w_planes, scales = w_planes
    _, N, H, W, C0 = src0.shape
    C1 = 0 if src1 is None else src1.shape[-1]
    Cout = w_planes.shape[1]
    dev = src0.device
    Ho, Wo = H // stride, W // stride
    out = torch.empty(((2, N, Ho, Wo, Cout) if split_out else (N, Ho, Wo, Cout)), device=dev,
                      dtype=torch.float16 if split_out else torch.float32)
    stats = None
    if want_stats:
        chunks = _lib.load().mf_op_conv_tc_stats_chunks(Ho, Wo)
        stats = torch.zeros((N, chunks, Cout // 8, 2), device=dev, dtype=torch.float32)
    _lib.check(_lib.load().mf_op_conv_tc(
        _p(src0), src0[0].numel(), C0, _p(src1), 0 if src1 is None else src1[0].numel(), C1, N, H, W,
        _p(w_planes), _p(scales), Cout, ksize, _p(bias), _p(out), out[0].numel() if split_out else 0,
        NHWC_SPLIT if split_out else NHWC_RAW, _p(stats), drain_interval, stride, _stream()), "conv_tc")
    return out, stats


def conv_simt(x, in_layout, w_kc, bias, Cin, ksize, stride, out_layout):
    """x: NCHW [N,C,H,W] / raw [N,H,W,C] / split [2,N,H,W,C]."""
    if in_layout == NCHW:
        N, _, H, W = x.shape
        plane = 0
    elif in_layout == NHWC_RAW:
        N, H, W, _ = x.shape
        plane = 0
    else:
        _, N, H, W, _ = x.shape
        plane = x[0].numel()
    Cout = w_kc.shape[1]
    pad = ksize // 2
    Ho = (H + 2 * pad - ksize) // stride + 1
    Wo = (W + 2 * pad - ksize) // stride + 1
    shape = {NCHW: (N, Cout, Ho, Wo), NHWC_RAW: (N, Ho, Wo, Cout), NHWC_SPLIT: (2, N, Ho, Wo, Cout)}[out_layout]
    out = torch.empty(shape, device=x.device, dtype=torch.float16 if out_layout == NHWC_SPLIT else torch.float32)
    _lib.check(_lib.load().mf_op_conv_simt(_p(x), plane, in_layout, N, Cin, H, W, _p(w_kc), _p(bias), Cout, ksize,
                                           stride, _p(out), out[0].numel() if out_layout == NHWC_SPLIT else 0,
                                           out_layout, _stream()), "conv_simt")
    return out


def gn_partial(raw):
    N, H, W, C = raw.shape
    part = torch.empty((N, 1, C // 8, 2), device=raw.device, dtype=torch.float32)
    _lib.check(_lib.load().mf_op_gn_partial(_p(raw), _p(part), N, H * W, C, _stream()), "gn_partial")
    return part


def gn_finalize(partial, C, G, HW, eps=1e-5):
    N, chunks = partial.shape[:2]
    mr = torch.empty((N, G, 2), device=partial.device, dtype=torch.float32)
    _lib.check(_lib.load().mf_op_gn_finalize(_p(partial), _p(mr), N, chunks, C, G, HW, eps, _stream()), "gn_finalize")
    return mr


def gn_apply(raw, mean_rstd, gamma, beta, G, res=None, emb=None):
    N, H, W, C = raw.shape
    out = torch.empty((2, N, H, W, C), device=raw.device, dtype=torch.float16)
    if res is None:
        kind, rplane = 0, 0
    elif res.dim() == 5:
        kind, rplane = 1, res[0].numel()
    else:
        kind, rplane = 2, 0
    _lib.check(_lib.load().mf_op_gn_apply(_p(raw), _p(mean_rstd), _p(gamma), _p(beta), _p(res), rplane, kind, _p(emb),
                                          0 if emb is None else emb.shape[1], _p(out), out[0].numel(), N, H * W, C, G,
                                          _stream()), "gn_apply")
    return out


def upsample2x(x_split):
    _, N, H, W, C = x_split.shape
    out = torch.empty((2, N, 2 * H, 2 * W, C), device=x_split.device, dtype=torch.float16)
    _lib.check(_lib.load().mf_op_upsample2x(_p(x_split), x_split[0].numel(), _p(out), out[0].numel(), N, H, W, C,
                                            _stream()), "upsample2x")
    return out


def upconv_tc(src_split, w_oihw, bias):
    """conv3x3(nearest_x2(src)) + bias via the folded four-phase tensor-core kernel -> split [2,N,2H,2W,Cout]"""
    _, N, H, W, C = src_split.shape
    Cout = w_oihw.shape[0]
    w = w_oihw.contiguous().float()
    wup = torch.empty((2, 4 * Cout, 4 * C), device=w.device, dtype=torch.float16)
    scales = torch.zeros(4, device=w.device, dtype=torch.float32)
    _lib.check(_lib.load().mf_op_prep_weight_up_tc(_p(w), _p(wup), _p(scales), Cout, C, _stream()), "prep_weight_up_tc")
    out = torch.empty((2, N, 2 * H, 2 * W, Cout), device=w.device, dtype=torch.float16)
    _lib.check(_lib.load().mf_op_upconv_tc(_p(src_split), src_split[0].numel(), C, N, H, W, _p(wup), _p(scales), Cout,
                                           _p(bias), _p(out), out[0].numel(), _stream()), "upconv_tc")
    return out
