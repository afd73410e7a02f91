#!/usr/bin/env python
"""CPU numerics study for round 2: would a Winograd F(2x2,3x3) formulation of the 3x3 convolutions (2.25x fewer MMAs) stay
inside the parity budget when its GEMMs use the same fp16x3 operand split as conv_tc?

Emulation (torch CPU): operands are split x = hi + lo with hi = fp16(x), lo = fp16(x - hi); a product keeps the three
terms hi*hi + lo*hi + hi*lo, each accumulated in fp32 (torch's fp32 conv/matmul as stand-in for the RN register sums);
the reference is the fp64 direct convolution.  Prints max / mean abs error and the worst |err| / (1e-5 + 1e-3 |y|).
    python tools/winograd_numerics.py
"""
import json

import torch
import torch.nn.functional as F

torch.manual_seed(0)
torch.set_num_threads(8)


def split(x, scale=1.0):
    """hi/lo fp16 split of x*scale (returned in the scaled domain)."""
    x = x * scale
    hi = x.half().float()
    lo = (x - hi).half().float()
    return hi, lo


def pow2_scale(w):
    """per-tensor power of two with max|w|*2^S in [2^12, 2^13): keeps the lo parts in fp16's normal range (prep_weight_tc)"""
    import math
    return 2.0 ** (12 - math.floor(math.log2(float(w.abs().max()))))


def conv3(x, w):          # fp16x3 direct: three fp32 convs
    xh, xl = split(x)
    sc = pow2_scale(w)
    wh, wl = split(w, sc)
    return (F.conv2d(xh, wh, padding=1) + F.conv2d(xl, wh, padding=1) + F.conv2d(xh, wl, padding=1)) / sc


G = torch.tensor([[1, 0, 0], [.5, .5, .5], [.5, -.5, .5], [0, 0, 1]], dtype=torch.float64)
Bt = torch.tensor([[1, 0, -1, 0], [0, 1, 1, 0], [0, -1, 1, 0], [0, 1, 0, -1]], dtype=torch.float64)
At = torch.tensor([[1, 1, 1, 0], [0, 1, -1, -1]], dtype=torch.float64)


def winograd(x, w, split_ops=True, u_from_fp64=True):
    N, C, H, W = x.shape
    K = w.shape[0]
    # weight transform once at load (fp64 -> fp32), input transform in fp32 like a GPU prologue would
    U = torch.einsum("ij,kcjl,ml->kcim", G, w.double() if u_from_fp64 else w.double(), G).float()      # [K,C,4,4]
    xp = F.pad(x, (1, 1, 1, 1))
    tiles = xp.unfold(2, 4, 2).unfold(3, 4, 2)                                                          # [N,C,H/2,W/2,4,4]
    V = torch.einsum("ij,nchwjl,ml->nchwim", Bt.float(), tiles, Bt.float())                             # fp32 transform
    if split_ops:
        Vh, Vl = split(V)
        sc = pow2_scale(U)
        Uh, Ul = split(U, sc)
        M = (torch.einsum("nchwim,kcim->nkhwim", Vh, Uh) + torch.einsum("nchwim,kcim->nkhwim", Vl, Uh) +
             torch.einsum("nchwim,kcim->nkhwim", Vh, Ul)) / sc
    else:
        M = torch.einsum("nchwim,kcim->nkhwim", V, U)
    Y = torch.einsum("ij,nkhwjl,ml->nkhwim", At.float(), M, At.float())                                  # [N,K,H/2,W/2,2,2]
    return Y.permute(0, 1, 2, 4, 3, 5).reshape(N, K, H, W)


def report(name, y, ref):
    d = (y.double() - ref).abs()
    tol = 1e-5 + 1e-3 * ref.abs()
    return {name: dict(max_abs=float(d.max()), mean_abs=float(d.mean()), worst_err_over_tol=float((d / tol).max()))}


for (N, C, K, H) in [(2, 256, 256, 32), (2, 512, 512, 16), (2, 1024, 1024, 8)]:
    x = torch.randn(N, C, H, H)
    x = x * torch.sigmoid(x) + 0.3 * torch.randn(N, C, H, H)          # Swish-like activations + residual
    w = torch.randn(K, C, 3, 3) / (C * 9) ** 0.5
    ref = F.conv2d(x.double(), w.double(), padding=1)
    out = {"shape": [N, C, K, H], "ref_rms": float(ref.pow(2).mean().sqrt())}
    out.update(report("direct_fp32", F.conv2d(x, w, padding=1), ref))
    out.update(report("direct_fp16x3", conv3(x, w), ref))
    out.update(report("winograd_fp32", winograd(x, w, split_ops=False), ref))
    out.update(report("winograd_fp16x3", winograd(x, w, split_ops=True), ref))
    print(json.dumps(out), flush=True)
