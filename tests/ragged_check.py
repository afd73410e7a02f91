#!/usr/bin/env python
"""Bring-up helper (test infrastructure, not collected by pytest): run odd-batch cases one per subprocess; re-run failing
ones — or those named in $SANITIZE — under compute-sanitizer memcheck.
usage: python tests/ragged_check.py            (driver)   |   python tests/ragged_check.py --case NAME"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]   # test_gpu_models imports the oracle

CASES = ["vae_canon_b1", "vae_small_encode", "unet_attn_b3", "unet_canon_b5", "unet_small_b3", "unet_small_b1", "vae_small_b3", "vae_small_b3_u8", "pipe_b3_eager",
         "pipe_b3_graph", "pipe_b1_graph", "dataset"]


def run(name):
    import torch
    from util import load_golden, make_unet, make_vae
    dev = torch.device("cuda:0")
    g = load_golden("sample_small.pt")
    gen = torch.Generator().manual_seed(0)
    if name.startswith("unet"):
        cfg = load_golden("unet_canonical.pt")["cfg"] if "canon" in name else g["unet_cfg"]
        B = int(name.rsplit("b", 1)[1])
        m = make_unet(cfg, dev)
        x = torch.randn(B, 8, 32, 32, generator=gen).to(dev)
        t = torch.randint(0, 1000, (B,), generator=gen).to(dev)
        y = m(x, t, (torch.arange(B) % 2).to(dev))[0]
        torch.cuda.synchronize()
        print(name, float(y.abs().mean()))
    elif name == "unet_attn_b3":
        ga = load_golden("unet_attn_small.pt")
        m = make_unet(ga["cfg"], dev)
        x = torch.randn(3, 8, 32, 32, generator=gen).to(dev)
        y = m(x, torch.randint(0, 1000, (3,), generator=gen).to(dev), (torch.arange(3) % 2).to(dev))[0]
        torch.cuda.synchronize()
        print(name, float(y.abs().mean()))
    elif name == "vae_canon_b1":
        # canonical widths at 256x256: the row-patch mode of conv_tc (128x128 / 256x256 levels, folded up-conv phases)
        import bench
        m = make_vae(bench.VAE_CFG, dev)
        y = m.decode(torch.randn(1, 8, 32, 32, generator=gen).to(dev))
        torch.cuda.synchronize()
        print(name, float(y.abs().mean()))
    elif name == "vae_small_encode":
        from test_gpu_models import _vae_cfg
        m = make_vae(_vae_cfg(g["vae_cfg"]), dev)
        z = m.encode(torch.rand(3, 3, 64, 64, generator=gen).to(dev) * 2 - 1)
        out, _, kl = m(torch.rand(1, 3, 64, 64, generator=gen).to(dev))
        torch.cuda.synchronize()
        print(name, float(z.abs().mean()), float(kl))
    elif name.startswith("vae"):
        from test_gpu_models import _vae_cfg
        m = make_vae(_vae_cfg(g["vae_cfg"]), dev)
        z = torch.randn(3, 8, 32, 32, generator=gen).to(dev)
        y = m.decode_uint8(z) if name.endswith("u8") else m.decode(z)
        torch.cuda.synchronize()
        print(name, float(y.float().abs().mean()))
    else:
        from test_gpu_models import _make_pipe
        pipe = _make_pipe(g)
        if name == "dataset":
            from medfusion_b200.sample_dataset import generate_dataset
            got = []
            generate_dataset(pipe, 7, label=1, steps=4, sample_batch=3, workers=2, sink=lambda c, a: got.append(c))
            print(name, sorted(got))
        else:
            B = 3 if "b3" in name else 1
            pipe.use_cuda_graph = name.endswith("graph")
            c = torch.full((B,), 1, device=dev)
            y = pipe.sample(B, (8, 32, 32), condition=c, steps=4)
            torch.cuda.synchronize()
            print(name, float(y.abs().mean()))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--case":
        run(sys.argv[2])
        sys.exit(0)
    bad = []
    for c in CASES:
        p = subprocess.run([sys.executable, __file__, "--case", c], capture_output=True, text=True, timeout=300)
        ok = p.returncode == 0
        print(("OK   " if ok else "FAIL ") + c, (p.stdout.strip().splitlines() or [""])[-1][:200], flush=True)
        if not ok:
            bad.append(c)
            print("   ", "\n    ".join(p.stderr.strip().splitlines()[-4:])[:800], flush=True)
    if os.environ.get("SANITIZE"):
        bad = [c for c in os.environ["SANITIZE"].split(",") if c]
    for c in bad[:6]:
        p = subprocess.run(["compute-sanitizer", "--tool", "memcheck", "--print-limit", "3", sys.executable, __file__,
                            "--case", c], capture_output=True, text=True, timeout=900)
        lines = [l for l in (p.stdout + p.stderr).splitlines() if "=========" in l]
        print("SANITIZER", c, "rc", p.returncode)
        print("\n".join(lines[:40])[:5000], flush=True)
