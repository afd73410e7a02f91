#!/usr/bin/env python
"""Per-launch device times of one canonical UNet step and one VAE.decode at B=64 (engine CUDA events), grouped by kernel
family.  python tools/step_profile.py  (GPU box)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch  # noqa: E402
from util import make_unet, make_vae  # noqa: E402
sys.path.insert(0, ROOT)
import bench  # noqa: E402

dev = "cuda:0"
if os.environ.get("SPLIT_FILL") is not None:      # small-batch fill knob (mf_set_split_fill): A/B runs
    from medfusion_b200 import _lib
    _lib.load().mf_set_split_fill(int(os.environ["SPLIT_FILL"]))
if os.environ.get("BLOCK_N") is not None:
    from medfusion_b200 import _lib
    _lib.load().mf_set_block_n(int(os.environ["BLOCK_N"]))
B = int(os.environ.get("B", "64"))
names = {0: "conv_tc", 1: "conv_simt", 2: "groupnorm family", 3: "other (embedding MLP, pack, attention)"}
LAT = int(os.environ.get("LATENT", "32"))          # 64 + ATTN=1 = BASELINE.json configs[3] (config 4 of SURVEY.md §8d)
ucfg = dict(bench.UNET_CFG)
if os.environ.get("ATTN"):
    ucfg["use_attention"] = ["none", "none", "none", "spatial"]
u = make_unet(ucfg, dev)
x = torch.randn(B, 8, LAT, LAT, device=dev)
t = torch.full((B,), 500, device=dev, dtype=torch.int64)
for _ in range(3):
    prof = u.profile(x, t, None)
tot = sum(p[0] for p in prof)
out = {"unet_step_ms": tot, "launches": len(prof)}
for k, n in names.items():
    sel = [p for p in prof if p[1] == k]
    out[n] = {"ms": sum(p[0] for p in sel), "launches": len(sel), "share": sum(p[0] for p in sel) / tot,
              "tflops": (sum(p[2] for p in sel) / max(1e-9, sum(p[0] for p in sel)) / 1e9) if k < 2 else None}
slow = sorted(((p[0], i, p[1], p[2]) for i, p in enumerate(prof)), reverse=True)[:int(os.environ.get("TOP", "8"))]
out["slowest"] = [dict(ms=round(a, 4), index=i, kind=names[k], gflop=round(f / 1e9, 1)) for a, i, k, f in slow]
if os.environ.get("DUMP"):
    out["launch_us"] = [[k, round(ms * 1e3, 1)] for ms, k, f in prof]
print(json.dumps(out))
v = make_vae(bench.VAE_CFG, dev)
z = torch.randn(B, 8, LAT, LAT, device=dev)
for _ in range(2):
    pv = v.profile(z)
totv = sum(p[0] for p in pv)
outv = {"vae_decode_ms": totv, "launches": len(pv)}
for k, n in names.items():
    sel = [p for p in pv if p[1] == k]
    outv[n] = {"ms": sum(p[0] for p in sel), "launches": len(sel), "share": sum(p[0] for p in sel) / totv,
               "tflops": (sum(p[2] for p in sel) / max(1e-9, sum(p[0] for p in sel)) / 1e9) if k < 2 else None}
print(json.dumps(outv))
