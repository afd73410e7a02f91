#!/usr/bin/env python
"""GPU bring-up probe: runs each kernel-level case in its own subprocess (with a timeout) and compares
with torch's fp32 CUDA ops (TF32 disabled).  Usage on the GPU box:  python tools/gpu_probe.py
Writes one JSON line per case to gpurun_out/probe.jsonl.
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {
    # name: dict(kind, ...)
    "simt_stem": dict(kind="simt", N=2, Cin=8, H=32, W=32, Cout=256, k=3, stride=1, inl=0, outl=2),
    "simt_down": dict(kind="simt", N=2, Cin=64, H=16, W=16, Cout=64, k=3, stride=2, inl=2, outl=2),
    "simt_head": dict(kind="simt", N=2, Cin=256, H=32, W=32, Cout=8, k=1, stride=1, inl=2, outl=0),
    "simt_rgb": dict(kind="simt", N=1, Cin=64, H=64, W=64, Cout=3, k=1, stride=1, inl=2, outl=0),
    "tc_1x1_k32_n64": dict(kind="tc", N=1, H=16, W=16, C0=64, C1=0, Cout=64, k=1),
    "tc_1x1_k256_n64": dict(kind="tc", N=1, H=16, W=16, C0=256, C1=0, Cout=64, k=1),
    "tc_1x1_k64_n128": dict(kind="tc", N=2, H=16, W=16, C0=64, C1=0, Cout=128, k=1),
    "tc_1x1_k64_n256": dict(kind="tc", N=2, H=16, W=16, C0=64, C1=0, Cout=256, k=1),
    "tc_3x3_16": dict(kind="tc", N=2, H=16, W=16, C0=64, C1=0, Cout=64, k=3),
    "tc_3x3_32_256": dict(kind="tc", N=2, H=32, W=32, C0=256, C1=0, Cout=256, k=3, stats=True),
    "tc_3x3_8_cat": dict(kind="tc", N=3, H=8, W=8, C0=128, C1=64, Cout=128, k=3, stats=True),
    "tc_3x3_64": dict(kind="tc", N=1, H=64, W=64, C0=64, C1=0, Cout=128, k=3, stats=True, split=True),
    "tc_3x3_256w": dict(kind="tc", N=1, H=8, W=256, C0=64, C1=0, Cout=64, k=3, stats=True),
    "tc_1x1_cat_512": dict(kind="tc", N=2, H=16, W=16, C0=512, C1=512, Cout=512, k=1),
    "tc_drain_sweep": dict(kind="tc_sweep", N=2, H=32, W=32, C0=256, C1=0, Cout=256, k=3),
    "tc_drain_sweep_8": dict(kind="tc_sweep", N=4, H=8, W=8, C0=1024, C1=1024, Cout=1024, k=3),
    "perf_32_256": dict(kind="perf", N=64, H=32, W=32, C0=256, C1=0, Cout=256, k=3),
    "perf_16_512": dict(kind="perf", N=64, H=16, W=16, C0=512, C1=0, Cout=512, k=3),
    "perf_8_1024": dict(kind="perf", N=64, H=8, W=8, C0=1024, C1=0, Cout=1024, k=3),
    "perf_8_cat": dict(kind="perf", N=64, H=8, W=8, C0=1024, C1=1024, Cout=1024, k=3),
    "perf_vae_128": dict(kind="perf", N=16, H=128, W=128, C0=128, C1=0, Cout=128, k=3),
    "perf_vae_256": dict(kind="perf", N=8, H=256, W=256, C0=64, C1=0, Cout=64, k=3),
    "norm": dict(kind="norm"),
    "upsample": dict(kind="upsample"),
}


def run_case(name):
    import torch
    import torch.nn.functional as F
    from medfusion_b200 import ops

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    c = CASES[name]
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(1234)
    res = dict(case=name)

    def rnd(*shape, scale=1.0):
        return (torch.randn(*shape, generator=g) * scale).to(dev)

    def errs(a, b):
        d = (a - b).abs()
        tol = 1e-5 + 1e-3 * b.abs()
        return dict(max_abs=float(d.max()), ref_absmax=float(b.abs().max()), viol=int((d > tol).sum()),
                    numel=d.numel())

    if c["kind"] == "simt":
        x = rnd(c["N"], c["Cin"], c["H"], c["W"])
        w = rnd(c["Cout"], c["Cin"], c["k"], c["k"], scale=0.1)
        b = rnd(c["Cout"])
        ref = F.conv2d(x, w, b, stride=c["stride"], padding=c["k"] // 2)
        xin = x if c["inl"] == 0 else ops.pack_split(x)
        out = ops.conv_simt(xin, c["inl"], ops.prep_weight_simt(w), b, c["Cin"], c["k"], c["stride"], c["outl"])
        got = out if c["outl"] == 0 else ops.unpack_nchw(out)
        res.update(errs(got, ref))
    elif c["kind"] == "tc":
        C = c["C0"] + c["C1"]
        x = rnd(c["N"], C, c["H"], c["W"])
        w = rnd(c["Cout"], C, c["k"], c["k"], scale=0.05)
        b = rnd(c["Cout"])
        ref = F.conv2d(x, w, b, padding=c["k"] // 2)
        s0 = ops.pack_split(x[:, :c["C0"]].contiguous())
        s1 = ops.pack_split(x[:, c["C0"]:].contiguous()) if c["C1"] else None
        assert ops.conv_tc_supported(c["N"], c["H"], c["W"], c["C0"], c["C1"], c["Cout"], c["k"]), "unsupported"
        out, stats = ops.conv_tc(s0, ops.prep_weight_tc(w), b, c["k"], src1=s1, split_out=c.get("split", False),
                                 want_stats=c.get("stats", False))
        torch.cuda.synchronize()
        got = ops.unpack_nchw(out)
        res.update(errs(got, ref))
        # single-pass TF32 error for scale (what an un-compensated kernel would give)
        xh, wh = x.half().float(), w.half().float()
        res["tf32_1pass_max_abs"] = float((F.conv2d(xh, wh, b, padding=c["k"] // 2) - ref).abs().max())
        if stats is not None:
            cs = ref.view(c["N"], c["Cout"] // 8, 8, -1)
            s_ref = cs.sum(dim=(2, 3))
            ss_ref = (cs * cs).sum(dim=(2, 3))
            st = stats.sum(dim=1)
            res["stats_sum_err"] = float((st[..., 0] - s_ref).abs().max())
            res["stats_sumsq_relerr"] = float(((st[..., 1] - ss_ref).abs() / ss_ref.abs().clamp_min(1e-6)).max())
    elif c["kind"] in ("tc_sweep", "perf"):
        C = c["C0"] + c["C1"]
        x = rnd(c["N"], C, c["H"], c["W"])
        w = rnd(c["Cout"], C, c["k"], c["k"], scale=0.05)
        b = rnd(c["Cout"])
        s0 = ops.pack_split(x[:, :c["C0"]].contiguous())
        s1 = ops.pack_split(x[:, c["C0"]:].contiguous()) if c["C1"] else None
        wp = ops.prep_weight_tc(w)
        flops = 2.0 * c["N"] * c["H"] * c["W"] * c["Cout"] * C * c["k"] ** 2
        res["sweep"] = {}
        ref = None
        if c["kind"] == "tc_sweep":
            ref = F.conv2d(x.double(), w.double(), b.double(), padding=c["k"] // 2)
            r32 = F.conv2d(x, w, b, padding=c["k"] // 2)
            res["torch_fp32_vs_fp64"] = float((r32.double() - ref).abs().max())
        from medfusion_b200 import _lib
        combos = ([(1, 256, d, 1) for d in (1, 2, 4)] + [(2, 256, 1, 1), (2, 256, 2, 1), (2, 256, 2, 0)]
                  if c["kind"] == "tc_sweep" else
                  [(2, 256, 2, 0), (2, 256, 2, 1), (2, 256, 1, 1), (2, 256, 1000, 1), (1, 256, 2, 1), (2, 128, 2, 1)])
        for cg, bn, di, sk in combos:
            _lib.load().mf_set_cta_group(cg)
            _lib.load().mf_set_block_n(bn)
            _lib.load().mf_set_stream_k(sk)
            out, _ = ops.conv_tc(s0, wp, b, c["k"], src1=s1, drain_interval=di)
            torch.cuda.synchronize()
            ent = {}
            if ref is not None:
                d = (ops.unpack_nchw(out).double() - ref).abs()
                ent["max_abs"] = float(d.max())
                ent["mean_abs"] = float(d.mean())
                ent["viol"] = int((d > 1e-5 + 1e-3 * ref.abs()).sum())
                ent["ref_rms"] = float(ref.pow(2).mean().sqrt())
            else:
                st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                for _ in range(3):
                    ops.conv_tc(s0, wp, b, c["k"], src1=s1, drain_interval=di)
                st.record()
                iters = 10
                for _ in range(iters):
                    ops.conv_tc(s0, wp, b, c["k"], src1=s1, drain_interval=di)
                en.record()
                torch.cuda.synchronize()
                ms = st.elapsed_time(en) / iters
                ent["ms"] = round(ms, 4)
                ent["tflops_alg"] = round(flops / ms / 1e9, 1)
            res["sweep"][f"cg{cg}_n{bn}_drain{di}_sk{sk}"] = ent
        res["viol"] = 0
    elif c["kind"] == "norm":
        N, Cc, H, W, G = 2, 64, 16, 16, 8
        x = rnd(N, Cc, H, W) * 2 + 0.5
        gamma, beta = rnd(Cc), rnd(Cc)
        r = rnd(N, Cc, H, W)
        emb = rnd(N, Cc)
        ref = F.group_norm(x, G, gamma, beta, 1e-5)
        ref = ref * torch.sigmoid(ref) + r + emb[:, :, None, None]
        raw = x.permute(0, 2, 3, 1).contiguous()
        part = ops.gn_partial(raw)
        mr = ops.gn_finalize(part, Cc, G, H * W)
        out = ops.gn_apply(raw, mr, gamma, beta, G, res=ops.pack_split(r), emb=emb)
        res.update(errs(ops.unpack_nchw(out), ref))
        out2 = ops.gn_apply(raw, mr, gamma, beta, G, res=r.permute(0, 2, 3, 1).contiguous(), emb=None)
        res["raw_res"] = errs(ops.unpack_nchw(out2), ref - emb[:, :, None, None])
    elif c["kind"] == "upsample":
        x = rnd(2, 32, 8, 8)
        ref = F.interpolate(x, size=(16, 16), mode="nearest-exact")
        res.update(errs(ops.unpack_nchw(ops.upsample2x(ops.pack_split(x))), ref))
    torch.cuda.synchronize()
    res["ok"] = res.get("viol", 1) == 0
    print("RESULT " + json.dumps(res), flush=True)


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--case":
        run_case(sys.argv[2])
        return
    names = sys.argv[1:] or list(CASES)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    out_path = os.path.join(ROOT, "gpurun_out", "probe.jsonl")
    with open(out_path, "a") as fh:
        for name in names:
            t0 = time.time()
            try:
                p = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", name], capture_output=True,
                                   text=True, timeout=240)
                line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
                if line:
                    rec = json.loads(line[-1][7:])
                else:
                    rec = dict(case=name, ok=False, rc=p.returncode, stdout=p.stdout[-1500:], stderr=p.stderr[-1500:])
            except subprocess.TimeoutExpired:
                rec = dict(case=name, ok=False, error="timeout")
            rec["secs"] = round(time.time() - t0, 1)
            fh.write(json.dumps(rec) + "\n")
            fh.flush()
            print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
