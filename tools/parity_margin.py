#!/usr/bin/env python
"""How much parity margin do the tuning knobs leave?  Runs UNet (canonical fixture) and VAE against the reference
fixtures for several (cta_group, drain_interval) settings and prints max abs error / violation counts.
    python tools/parity_margin.py            (GPU box)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch  # noqa: E402
from util import load_golden, make_unet, make_vae, violations  # noqa: E402
from medfusion_b200 import _lib  # noqa: E402

dev = "cuda:0"
g = load_golden("unet_canonical.pt")
gv = load_golden("vae_canonical.pt")
keep = ("in_channels", "out_channels", "emb_channels", "spatial_dims", "hid_chs", "kernel_sizes", "strides",
        "deep_supervision", "use_attention")
lib = _lib.load()
for cg, di, eps in [(2, 1, 0.0), (2, 1, -1.0), (2, 2, 0.0), (2, 2, -1.0), (2, 3, -1.0), (2, 4, -1.0), (1, 2, -1.0)]:
    lib.mf_set_cta_group(cg)
    lib.mf_set_drain_interval(di)
    lib.mf_set_debias_eps(eps)
    m = make_unet(g["cfg"], dev)          # fresh module -> fresh plan with the new knobs
    y, _ = m(g["x"].to(dev), g["t"].to(dev), g["cond"].to(dev))
    n, mx, rmax = violations(y.cpu(), g["y_cond"])
    d = (y.cpu().double() - g["y_cond"].double()).abs()
    tol = 1e-5 + 1e-3 * g["y_cond"].double().abs()
    v = make_vae({k: v for k, v in gv["cfg"].items() if k in keep}, dev)
    x = v.decode(gv["z"].to(dev))
    nv, mxv, rmaxv = violations(x.cpu(), gv["x"])
    dv = (x.cpu().double() - gv["x"].double()).abs()
    tolv = 1e-5 + 1e-3 * gv["x"].double().abs()
    print(json.dumps(dict(cta_group=cg, drain=di, debias=eps, unet_viol=n, unet_max_abs=mx, unet_ref_max=rmax,
                          unet_worst_err_over_tol=float((d / tol).max()), unet_mean_abs=float(d.mean()),
                          vae_viol=nv, vae_max_abs=mxv, vae_worst_err_over_tol=float((dv / tolv).max()))), flush=True)
    del m, v
