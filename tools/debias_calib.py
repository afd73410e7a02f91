#!/usr/bin/env python
"""Calibrate the accumulator de-bias factor: for each (drain, eps) measure signed bias and abs error vs fp64.
    python tools/debias_calib.py     (GPU box)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from medfusion_b200 import _lib, ops  # noqa: E402

dev = "cuda:0"
lib = _lib.load()
lib.mf_set_cta_group(2)
cases = {
    "gauss_k2304": dict(N=2, H=32, W=32, C=256, Cout=256, k=3, dist="gauss"),
    "gauss_k9216": dict(N=2, H=16, W=16, C=1024, Cout=256, k=3, dist="gauss"),
    "pos_k2304": dict(N=2, H=32, W=32, C=256, Cout=256, k=3, dist="pos"),     # all-positive products: monotone growth
    "silu_k4608": dict(N=2, H=16, W=16, C=512, Cout=256, k=3, dist="silu"),   # activation-like inputs
    "gauss_1x1_k1024": dict(N=2, H=16, W=16, C=1024, Cout=256, k=1, dist="gauss"),
}
for name, c in cases.items():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(c["N"], c["C"], c["H"], c["W"], generator=g)
    w = torch.randn(c["Cout"], c["C"], c["k"], c["k"], generator=g) / (c["C"] * c["k"] ** 2) ** 0.5
    if c["dist"] == "pos":
        x, w = x.abs(), w.abs()
    if c["dist"] == "silu":
        x = x * torch.sigmoid(x) + 0.3
    b = torch.zeros(c["Cout"])
    x, w, b = x.to(dev), w.to(dev), b.to(dev)
    ref = F.conv2d(x.double(), w.double(), None, padding=c["k"] // 2)
    s0, wp = ops.pack_split(x), ops.prep_weight_tc(w)
    scale = float(ref.abs().mean())
    for drain in (1, 2, 4):
        for eps in (0.0, 1e-7, 1.5e-7, 2e-7, 2.5e-7, 3e-7, 4e-7):
            lib.mf_set_debias_eps(eps)
            out, _ = ops.conv_tc(s0, wp, b, c["k"], drain_interval=drain)
            d = ops.unpack_nchw(out).double() - ref
            big = ref.abs() > 0.5 * scale
            rel_bias = float((d[big] / ref[big]).mean())
            print(json.dumps(dict(case=name, drain=drain, eps=eps, rel_bias=rel_bias, mean_abs=float(d.abs().mean()),
                                  max_abs=float(d.abs().max()), ref_mean_abs=scale)), flush=True)
