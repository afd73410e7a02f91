#!/usr/bin/env python
"""One canonical UNet forward at B (default 64) — a short target for ncu kernel captures.  python tools/unet_once.py [n]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
from util import make_unet
import bench
B = int(os.environ.get("B", "64"))
LAT = int(os.environ.get("LATENT", "32"))
cfg = dict(bench.UNET_CFG_ATTN if os.environ.get("ATTN") else bench.UNET_CFG)
u = make_unet(cfg, "cuda:0")
x = torch.randn(B, 8, LAT, LAT, device="cuda:0")
t = torch.full((B,), 500, device="cuda:0", dtype=torch.int64)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    y, _ = u(x, t, None)
torch.cuda.synchronize()
print("ok", float(y.abs().max()))
