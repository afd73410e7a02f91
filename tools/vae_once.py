#!/usr/bin/env python
"""VAE.decode at B (default 64) — a short target for ncu kernel captures.  python tools/vae_once.py [n]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
from util import make_vae
import bench
B = int(os.environ.get("B", "64"))
v = make_vae(bench.VAE_CFG, "cuda:0")
z = torch.randn(B, 8, 32, 32, device="cuda:0")
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    y = v.decode(z)
torch.cuda.synchronize()
print("ok", float(y.abs().max()))
