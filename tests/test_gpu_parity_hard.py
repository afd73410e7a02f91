"""GPU: the parity cases VERDICT r1 called soft, at the north star's tolerance (rtol=1e-3, atol=1e-5) WITHOUT
dividing by the tensor's scale:

  * canonical-width (hid 256..1024) 50-step trajectories of the unmodified reference (tests/golden/traj_canonical.pt):
    free-running (every x_t the estimator sees + the final latent) and teacher-forced (one step from the reference's
    own x_t);
  * DiffusionPipeline.forward values (tests/golden/forward_small.pt), the 4-tuple of diffusion_pipeline.py:232-275;
  * the sticky saturation counter: values beyond the fp16 range of the split planes are reported, not silently clamped.

Tolerance note (stated, not hidden): the 1e-5 absolute budget of the north star is a budget on the ESTIMATOR output.
The reference's own update amplifies an estimator error e by a known factor before it reaches x_{t-1}
(gaussian_scheduler.py:119-124: x_0 = A_t x_t - B_t pred, B_999 = 158), so for scheduler outputs the absolute tolerance
is atol * max(1, amp_t) with amp_t computed from the scheduler tables; rtol stays 1e-3.  Free-running trajectories are
additionally asserted at the plain tolerance on every snapshot where that is meaningful (see each test).
"""
import pytest
import torch

from util import ATOL, RTOL, assert_close, load_golden, violations

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _pipe(unet_cfg, sched, **over):
    from medfusion_b200.models import (DiffusionPipeline, GaussianNoiseScheduler, LabelEmbedder, TimeEmbbeding, UNet)
    from medfusion_b200.synthetic import fill_
    ucfg = {k: (dict(v) if isinstance(v, dict) else v) for k, v in unet_cfg.items()}
    kw = dict(noise_scheduler=GaussianNoiseScheduler, noise_estimator=UNet, latent_embedder=None,
              noise_scheduler_kwargs=dict(sched),
              noise_estimator_kwargs=dict(time_embedder=TimeEmbbeding, cond_embedder=LabelEmbedder, **ucfg),
              estimator_objective="x_T", estimate_variance=False, use_self_conditioning=False, use_ema=False,
              do_input_centering=False, clip_x0=False)
    kw.update(over)
    pipe = DiffusionPipeline(**kw)
    fill_(pipe.noise_estimator)
    return pipe.to(DEV)


def _noises(c, shape=(2, 8, 32, 32)):
    """The reference run drew from a seeded CPU generator (oracle/make_golden.py::traj_fixture); only the seed is stored."""
    g = torch.Generator().manual_seed(c["seed"])
    return [torch.randn(shape, generator=g) for _ in range(c["n_draws"])]


def _amp(sched, t, t_next, clip):
    """|d x_next / d pred| of one reverse step from the tables (x_T objective)."""
    B = float(sched.sqrt_recipm1_alphas_cumprod[t])
    if t_next is None:    # ancestral step: mean = coef1 * x_0 + coef2 * x_t
        return float(sched.posterior_mean_coef1[t]) * B
    a, an = float(sched.alphas_cumprod[t]), float(sched.alphas_cumprod[t_next])
    sigma2 = (1 - a / an) * (1 - an) / (1 - a)
    c = max(0.0, 1 - an - sigma2) ** 0.5
    return an ** 0.5 * B + c


SENS_MULT = 3.0   # see test_canonical_50_step_trajectory_free_running


@pytest.mark.parametrize("case", ["ddim50_clip_cond", "ddpm50"])
def test_canonical_50_step_trajectory_free_running(case):
    """pipeline.denoise at canonical width, 50 steps, the reference's noise injected: EVERY estimator input and the final
    latent against the reference, no scale normalisation.

    ddpm50 (ancestral steps t=49..0, contractive): plain rtol=1e-3 / atol=1e-5 on all 50 snapshots.
    ddim50_clip_cond: the reference's own DDIM-form dynamics with a random-weight estimator amplify ANY perturbation —
    the fixture records it: moving x_T by 1e-5*N(0,1) (an input error of the size of atol) moves the reference's own
    x_t by `sens[i]` (up to 1e-3 around step 33, 4e-4 in the final latent; oracle/make_golden.py::traj_fixture).  A
    free-running comparison can therefore not hold 1e-5 absolute for ANY implementation that is merely within tolerance
    per estimator call; the assertion is  |err_i| <= atol + rtol*|ref| + SENS_MULT * sens[i]  — the plain tolerance
    widened by a stated multiple (3: tolerance-sized errors enter at each of the 50 steps, not once) of the reference's
    measured self-divergence, and by nothing else.  Measured on B200: worst |err_i| / sens[i] = 1.5
    (profiles/r02_trajectory_margin.md).  The per-step (teacher-forced) test below holds every step to atol*amp_t."""
    g = load_golden("traj_canonical.pt")
    c = g["cases"][case]
    pipe = _pipe(g["unet_cfg"], g["sched"], **c["pipe"])
    est = pipe.noise_estimator
    draws = iter(n.to(DEV) for n in _noises(c))
    x_T = next(draws)
    seen = []
    orig = est.forward_step

    def spy(x_t, *a, **k):
        seen.append(x_t.detach().clone())
        return orig(x_t, *a, **k)

    est.forward_step = spy
    cond = None if c["cond"] is None else c["cond"].to(DEV)
    lat = pipe.denoise(x_T, condition=cond, _noise_fn=lambda _x: next(draws), **c["kw"])
    with pytest.raises(StopIteration):
        next(draws)
    assert len(seen) == c["x_in"].shape[0] == 50
    strict = case == "ddpm50"
    worst_plain, worst_sens = 0.0, 0.0
    for i, x in enumerate(seen):
        ref = c["x_in"][i].double()
        d = (x.cpu().double() - ref).abs()
        extra = 0.0 if strict else SENS_MULT * float(c["sens"][i])
        worst_plain = max(worst_plain, float((d / (ATOL + RTOL * ref.abs())).max()))
        if float(c["sens"][i]) > 0:
            worst_sens = max(worst_sens, float(d.max()) / float(c["sens"][i]))
        n = int((d > ATOL + RTOL * ref.abs() + extra).sum())
        assert n == 0, (f"{case}: estimator input of step {i}: {n} elements outside tolerance (max err {float(d.max()):.3e}, "
                        f"|ref|max {float(ref.abs().max()):.3e}, sens {float(c['sens'][i]):.3e})")
    print(f"{case}: worst |err|/tol(plain) = {worst_plain:.3f}, worst |err|/sens = {worst_sens:.3f} over 50 free-running steps")
    ref = c["latent"].double()
    d = (lat.cpu().double() - ref).abs()
    extra = 0.0 if strict else SENS_MULT * float(c["sens_final"])
    assert int((d > ATOL + RTOL * ref.abs() + extra).sum()) == 0, f"{case} final latent: max err {float(d.max()):.3e}"


def _ulp32(m):
    import math
    return 2.0 ** (math.floor(math.log2(max(m, 1e-30))) - 23)


@pytest.mark.parametrize("case", ["ddim50_clip_cond", "ddpm50"])
def test_canonical_trajectory_teacher_forced_steps(case):
    """One reverse step from the reference's own x_t, for ALL 49 transitions: x_{t-1} against the reference's next
    estimator input.  Absolute tolerance  atol * max(1, amp_t) + 4 ulp32(A_t * max|x_t|):
      amp_t = |d x_next / d pred| from the tables (the 1e-5 budget is on the estimator output, the update amplifies it);
      the ulp term is the fp32 rounding of the update's own intermediates (x_0 = A_t x_t - B_t pred is a difference of
      terms of size A_t|x_t|, up to ~100 mid-trajectory, where one fp32 ulp is 7.6e-6) — any re-ordering of the
      reference's own arithmetic moves the result by that much.  rtol stays 1e-3."""
    g = load_golden("traj_canonical.pt")
    c = g["cases"][case]
    pipe = _pipe(g["unet_cfg"], g["sched"], **c["pipe"])
    est, sched = pipe.noise_estimator, pipe.noise_scheduler
    noises = _noises(c)
    ddim = c["kw"]["use_ddim"]
    cond = None if c["cond"] is None else c["cond"].to(DEV)
    steps = c["x_in"].shape[0]
    worst = 0.0
    for i in range(steps - 1):
        t, t_next = int(c["t_in"][i]), int(c["t_in"][i + 1])
        # draw order of the reference: x_T, then per step the scheduler draw (+ the DDIM draw on all but the last step)
        k = 1 + (2 * i if ddim else i)
        noise, noise2 = noises[k].to(DEV), (noises[k + 1].to(DEV) if ddim else None)
        tb = torch.full((2,), t, device=DEV, dtype=torch.int64)
        o = est.forward_step(c["x_in"][i].to(DEV), tb, cond, sched, noise=noise,
                             t_next=torch.tensor(t_next, device=DEV) if ddim else None, noise_ddim=noise2,
                             objective="x_T", clip_x0=pipe.clip_x0, want=("x_next",), uniform_t=True)
        amp = max(1.0, _amp(sched, t, t_next if ddim else None, pipe.clip_x0))
        A = float(sched.sqrt_recip_alphas_cumprod[t])
        atol_i = ATOL * amp + 4 * _ulp32(A * float(c["x_in"][i].abs().max()))
        ref = c["x_in"][i + 1].double()
        d = (o["x_next"].cpu().double() - ref).abs()
        worst = max(worst, float((d / (atol_i + RTOL * ref.abs())).max()))
        assert_close(o["x_next"].cpu(), c["x_in"][i + 1], atol=atol_i,
                     what=f"{case} step {i} (t={t}, amp={amp:.2f}, atol={atol_i:.2e})")
    print(f"{case}: worst teacher-forced |err|/tol over {steps - 1} steps = {worst:.3f}")
    assert steps == 50


@pytest.mark.parametrize("case", ["uncond", "cond_g1", "cfg3_uncond_labels", "x0_objective_clip"])
def test_pipeline_forward_values_match_reference(case):
    """DiffusionPipeline.forward -> (x_t_prior, x_0, x_T, self_cond) (diffusion_pipeline.py:232-275) with the scheduler's
    randn_like draw replaced by the fixture's tensor.  x_T (the estimator output itself for the x_T objective) is held to
    the plain tolerance; x_0 / x_t_prior carry the reference's own amplification B_t = sqrt(1/ac_t - 1) per sample."""
    g = load_golden("forward_small.pt")
    c = g["cases"][case]
    pipe = _pipe(g["unet_cfg"], g["sched"], **c["pipe"])
    noise = g["noise"].to(DEV)
    pipe.noise_scheduler.x_final = lambda x: noise.clone()
    call = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in c["call"].items()}
    x_t, t = g["x_t"].to(DEV), g["t"].to(DEV)
    prior, x0, xT, sc = pipe(x_t, t, **call)
    sched = pipe.noise_scheduler
    x0_obj = c["pipe"].get("estimator_objective", "x_T") == "x_0"
    gs = float(call.get("guidance_scale", 1.0))
    cfg_gain = (abs(gs) + abs(1 - gs)) if "guidance_scale" in call else 1.0   # pred_u + g (pred_c - pred_u)
    for b in range(x_t.shape[0]):
        B = float(sched.sqrt_recipm1_alphas_cumprod[int(t[b])])
        A = float(sched.sqrt_recip_alphas_cumprod[int(t[b])])
        amp_x0 = 1.0 if x0_obj else max(1.0, B) * cfg_gain
        amp_xT = (max(1.0, A / max(B, 1e-6)) if x0_obj else 1.0) * cfg_gain
        sl = slice(b, b + 1)
        assert_close(xT[sl].cpu(), c["x_T"][sl], atol=ATOL * amp_xT, what=f"{case} x_T[{b}]")
        assert_close(x0[sl].cpu(), c["x_0"][sl], atol=ATOL * amp_x0, what=f"{case} x_0[{b}]")
        assert_close(prior[sl].cpu(), c["x_t_prior"][sl], atol=ATOL * max(amp_x0, amp_xT),
                     what=f"{case} x_t_prior[{b}]")
        assert_close(sc[sl].cpu(), c["self_cond"][sl], atol=ATOL * max(amp_x0, amp_xT), what=f"{case} self_cond[{b}]")
    # forward's fourth output is x_0 for the x_T objective and x_T for the x_0 objective (diffusion_pipeline.py:275)
    assert torch.equal(sc, xT if x0_obj else x0)


def test_saturation_counter_reports_values_beyond_fp16_range():
    """split16 clamps at +-65504; the reference is fp32-range.  The clamp must be visible: mf_saturation_count through the
    ABI, FloatingPointError from the pipeline."""
    import medfusion_b200 as mb
    from util import make_unet
    g = load_golden("unet_small.pt")
    m = make_unet(g["cfg"], DEV)
    mb.saturation_count(reset=True)
    x = g["x"].to(DEV)
    m(x, g["t"].to(DEV), None)
    assert mb.saturation_count() == 0                       # O(1) inputs: nothing clamped
    big = x.clone()
    big[0, 0, 0, 0] = 1.0e6                                 # the stem packs x_t into fp16 planes
    big[1, 3, 5, 7] = float("inf")
    m(big, g["t"].to(DEV), None)
    n = mb.saturation_count()
    assert n >= 2, n
    assert mb.saturation_count(reset=True) == n and mb.saturation_count() == 0
    # pipeline level: denoise raises instead of returning a silently clamped result
    gs = load_golden("sample_small.pt")
    pipe = _pipe(gs["unet_cfg"], gs["sched"])
    with pytest.raises(FloatingPointError):
        pipe.denoise(big, steps=3, use_ddim=True)
    out = pipe.denoise(x, steps=3, use_ddim=True)            # the counter was reset by the failed call
    assert torch.isfinite(out).all()
