#!/usr/bin/env python
"""Per-launch times of one VAE.decode at B (default 64): python tools/vae_ops.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
from util import make_vae
import bench
B = int(os.environ.get("B", "64"))
v = make_vae(bench.VAE_CFG, "cuda:0")
z = torch.randn(B, 8, 32, 32, device="cuda:0")
for _ in range(3):
    pv = v.profile(z)
names = {0: "conv_tc", 1: "conv_simt", 2: "gn", 3: "other"}
for i, (ms, k, fl) in enumerate(pv):
    print(i, names[k], round(ms, 4), "ms", round(fl / 1e9, 1), "GF", round(fl / ms / 1e9, 1) if ms > 0 and fl else "", "TF/s")
print("total", sum(p[0] for p in pv))
