#!/usr/bin/env python
"""Attention core, tcgen05 vs CUDA-core kernel, at the config-4 site shapes: python tools/attn_bench.py"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from medfusion_b200 import _lib
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
for (B, N, heads, d) in ((32, 256, 8, 128), (32, 256, 8, 64), (64, 64, 8, 128)):
    C = heads * d
    qkv = torch.randn(B, N, 3 * C, device="cuda")
    out = torch.empty((2, B, N, 1, C), device="cuda", dtype=torch.float16)
    res = {}
    for mode in (1, 0):
        lib.mf_set_attn_tc(mode)
        def run():
            _lib.check(lib.mf_op_attention(qkv.data_ptr(), qkv.data_ptr() + 4 * C, qkv.data_ptr() + 8 * C, 3 * C, out.data_ptr(),
                                           out[0].numel(), B, N, heads, d, st), "attention")
        for _ in range(3): run()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20): run()
        b.record(); torch.cuda.synchronize()
        res["tcgen05" if mode else "cuda_core"] = round(a.elapsed_time(b) / 20, 4)
    res.update(B=B, N=N, heads=heads, d=d, gflop=round(4.0 * B * heads * N * N * d / 1e9, 2))
    print(json.dumps(res), flush=True)
