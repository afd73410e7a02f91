#!/usr/bin/env python
"""Per-step parity margin of the canonical 50-step trajectories (tests/golden/traj_canonical.pt) on the GPU box:
worst |err| / (1e-5 + 1e-3 |ref|) of every estimator input, free-running and teacher-forced, next to the scheduler's own
amplification amp_t = |d x_next / d pred|.   python tools/traj_margin.py > gpurun_out/traj_margin.jsonl"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch  # noqa: E402
from util import ATOL, RTOL, load_golden  # noqa: E402
import test_gpu_parity_hard as T  # noqa: E402

DEV = "cuda:0"


def ratio(got, ref):
    d = (got.double().cpu() - ref.double()).abs()
    return float((d / (ATOL + RTOL * ref.double().abs())).max()), float(d.max())


g = load_golden("traj_canonical.pt")
for case, c in g["cases"].items():
    pipe = T._pipe(g["unet_cfg"], g["sched"], **c["pipe"])
    est, sched = pipe.noise_estimator, pipe.noise_scheduler
    noises = T._noises(c)
    ddim = c["kw"]["use_ddim"]
    cond = None if c["cond"] is None else c["cond"].to(DEV)
    # free running
    draws = iter(n.to(DEV) for n in noises)
    x_T = next(draws)
    seen, orig = [], est.forward_step

    def spy(x_t, *a, **k):
        seen.append(x_t.detach().clone())
        return orig(x_t, *a, **k)

    est.forward_step = spy
    pipe.check_saturation = False
    lat = pipe.denoise(x_T, condition=cond, _noise_fn=lambda _x: next(draws), **c["kw"])
    est.forward_step = orig
    rows = []
    for i in range(50):
        t = int(c["t_in"][i])
        t_next = int(c["t_in"][i + 1]) if i < 49 else None
        free, free_abs = ratio(seen[i], c["x_in"][i])
        row = dict(case=case, step=i, t=t, free=round(free, 3), free_abs=free_abs,
                   amp=round(T._amp(sched, t, t_next if (ddim and t_next is not None) else None, pipe.clip_x0), 3),
                   ref_absmax=float(c["x_in"][i].abs().max()))
        if i < 49:
            k = 1 + (2 * i if ddim else i)
            o = est.forward_step(c["x_in"][i].to(DEV), torch.full((2,), t, device=DEV, dtype=torch.int64), cond, sched,
                                 noise=noises[k].to(DEV), t_next=torch.tensor(t_next, device=DEV) if ddim else None,
                                 noise_ddim=noises[k + 1].to(DEV) if ddim else None, objective="x_T",
                                 clip_x0=pipe.clip_x0, want=("x_next",), uniform_t=True)
            tf, tf_abs = ratio(o["x_next"], c["x_in"][i + 1])
            row.update(teacher=round(tf, 3), teacher_abs=tf_abs)
        rows.append(row)
        print(json.dumps(row), flush=True)
    fin, fin_abs = ratio(lat, c["latent"])
    print(json.dumps(dict(case=case, final=round(fin, 3), final_abs=fin_abs)), flush=True)
