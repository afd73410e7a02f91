#!/usr/bin/env python
"""Generate tests/golden/*.pt by running the UNMODIFIED reference (/root/reference) on CPU fp32.

Run in the build container only (the reference is not present on the GPU box):
    python oracle/make_golden.py
Third-party packages the reference imports but that are not installable offline (monai, pytorch_lightning,
streamlit, lpips, pytorch_msssim) are provided by oracle/shims/ — minimal stand-ins for the symbols used.
Weights are NOT stored: both sides regenerate them from medfusion_b200.synthetic (per-key seeded), which also
overwrites the reference's zero-initialised tensors (SURVEY.md finding 4).  Noise is injected by replacing
torch.randn_like with a recording seeded CPU generator, so trajectories are reproducible on any device.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle", "shims"), "/root/reference", ROOT]

import torch  # noqa: E402

from medfusion_b200.synthetic import fill_  # noqa: E402  (RNG recipe only, no compute)
from medical_diffusion.models.estimators import UNet  # noqa: E402
from medical_diffusion.models.embedders import TimeEmbbeding, LabelEmbedder  # noqa: E402
from medical_diffusion.models.embedders.latent_embedders import VAE, VQVAE  # noqa: E402
from medical_diffusion.models.noise_schedulers import GaussianNoiseScheduler  # noqa: E402
from medical_diffusion.models.pipelines import DiffusionPipeline  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
torch.set_num_threads(8)

UNET_SMALL = dict(in_ch=8, out_ch=8, spatial_dims=2, hid_chs=[64, 64, 128, 256], kernel_sizes=[3, 3, 3, 3],
                  strides=[1, 2, 2, 2], norm_name=("GROUP", {"num_groups": 8, "affine": True}),
                  time_embedder_kwargs={"emb_dim": 256}, cond_embedder_kwargs={"emb_dim": 256, "num_classes": 2},
                  deep_supervision=False, use_res_block=True, use_attention="none")
UNET_ATTN = dict(in_ch=8, out_ch=8, spatial_dims=2, hid_chs=[64, 64, 256, 256], kernel_sizes=[3, 3, 3, 3],
                 strides=[1, 2, 2, 2], norm_name=("GROUP", {"num_groups": 8, "affine": True}),
                 time_embedder_kwargs={"emb_dim": 256}, cond_embedder_kwargs={"emb_dim": 256, "num_classes": 2},
                 deep_supervision=False, use_res_block=True, use_attention=["none", "linear", "none", "spatial"])
UNET_CANON = dict(in_ch=8, out_ch=8, spatial_dims=2, hid_chs=[256, 256, 512, 1024], kernel_sizes=[3, 3, 3, 3],
                  strides=[1, 2, 2, 2], time_embedder_kwargs={"emb_dim": 1024},
                  cond_embedder_kwargs={"emb_dim": 1024, "num_classes": 2}, deep_supervision=False,
                  use_res_block=True, use_attention="none")
VAE_CANON = dict(in_channels=3, out_channels=3, emb_channels=8, spatial_dims=2, hid_chs=[64, 128, 256, 512],
                 kernel_sizes=[3, 3, 3, 3], strides=[1, 2, 2, 2], deep_supervision=1, use_attention="none")
VAE_SMALL = dict(in_channels=3, out_channels=3, emb_channels=8, spatial_dims=2, hid_chs=[64, 128],
                 kernel_sizes=[3, 3], strides=[1, 2], deep_supervision=False, use_attention="none")
SCHED = dict(timesteps=1000, beta_start=0.002, beta_end=0.02, schedule_strategy="scaled_linear")


def fresh(cfg):
    """Deep-ish copy; also hands TimeEmbbeding a FRESH pos_embedder_kwargs dict: the reference mutates its
    mutable default argument (time_embedder.py:57,63), which would leak the first model's sinusoidal width
    into every later model built in the same process."""
    kw = {k: (dict(v) if isinstance(v, dict) else v) for k, v in cfg.items()}
    kw["time_embedder_kwargs"] = dict(kw["time_embedder_kwargs"], pos_embedder_kwargs={})
    return kw


def make_unet(cfg):
    kw = fresh(cfg)
    m = UNet(time_embedder=TimeEmbbeding, cond_embedder=LabelEmbedder, **kw).eval()
    return fill_(m)


def gen(seed):
    return torch.Generator(device="cpu").manual_seed(seed)


@torch.no_grad()
def unet_fixture(name, cfg, seed):
    m = make_unet(cfg)
    g = gen(seed)
    x = torch.randn(2, 8, 32, 32, generator=g)
    t = torch.tensor([999, 17])
    c = torch.tensor([1, 0])
    y_c, _ = m(x, t, c)
    y_u, _ = m(x, t, None)
    keys = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    torch.save(dict(cfg=cfg, x=x, t=t, cond=c, y_cond=y_c, y_uncond=y_u, keys=keys), os.path.join(OUT, name))
    print(name, float(y_c.abs().max()), float(y_u.abs().max()))


@torch.no_grad()
def vae_fixture():
    m = fill_(VAE(loss=torch.nn.MSELoss, **VAE_CANON).eval())
    z = torch.randn(1, 8, 32, 32, generator=gen(11))
    x = m.decode(z)
    z2 = torch.randn(3, 8, 8, 8, generator=gen(12))
    x2 = m.decode(z2)
    keys = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    torch.save(dict(cfg=VAE_CANON, z=z, x=x, z2=z2, x2=x2, keys=keys), os.path.join(OUT, "vae_canonical.pt"))
    print("vae", float(x.abs().max()), float(x2.abs().max()))


@torch.no_grad()
def sched_fixture():
    s = GaussianNoiseScheduler(**SCHED)
    g = gen(21)
    x_t = torch.randn(5, 8, 16, 16, generator=g)
    pred = torch.randn(5, 8, 16, 16, generator=g)
    noise = torch.randn(5, 8, 16, 16, generator=g)
    t = torch.tensor([999, 500, 37, 1, 0])
    out = {}
    orig = torch.randn_like
    torch.randn_like = lambda x, **k: noise.clone()
    try:
        for clip in (False, True):
            p, x0 = s.estimate_x_t_prior_from_x_T(x_t, t, pred, clip_x0=clip, var_scale=0)
            out[f"xT_clip{int(clip)}"] = dict(prior=p, x_0=x0)
            p, x0 = s.estimate_x_t_prior_from_x_0(x_t, t, pred, clip_x0=clip, var_scale=0)
            xT = s.estimate_x_T(x_t, x_0=pred, t=t, clip_x0=clip)
            out[f"x0_clip{int(clip)}"] = dict(prior=p, x_0=x0, x_T=xT)
    finally:
        torch.randn_like = orig
    buffers = {k: v.clone() for k, v in s.state_dict().items()}
    torch.save(dict(sched=SCHED, x_t=x_t, pred=pred, noise=noise, t=t, out=out, buffers=buffers),
               os.path.join(OUT, "sched.pt"))
    print("sched ok")


@torch.no_grad()
def sample_fixture():
    pipe = DiffusionPipeline(noise_scheduler=GaussianNoiseScheduler, noise_estimator=UNet, latent_embedder=None,
                             noise_scheduler_kwargs=dict(SCHED),
                             noise_estimator_kwargs=dict(time_embedder=TimeEmbbeding, cond_embedder=LabelEmbedder,
                                                         **fresh(UNET_SMALL)),
                             estimator_objective="x_T", estimate_variance=False, use_self_conditioning=False,
                             use_ema=False, do_input_centering=False, clip_x0=False).eval()
    fill_(pipe.noise_estimator)
    pipe.latent_embedder = fill_(VAE(loss=torch.nn.MSELoss, **VAE_SMALL).eval())
    cases = {
        "ddim5": dict(n=2, kw=dict(steps=5, use_ddim=True)),
        "ddpm4": dict(n=2, kw=dict(steps=4, use_ddim=False)),
        "cfg_ddim3": dict(n=2, kw=dict(steps=3, use_ddim=True, guidance_scale=3.0), cond=torch.tensor([0, 1])),
        "cond_ddim3_g1": dict(n=2, kw=dict(steps=3, use_ddim=True, guidance_scale=1.0), cond=torch.tensor([1, 0])),
    }
    out = {}
    orig = torch.randn_like
    for name, c in cases.items():
        g = gen(100 + len(out))
        rec = []

        def fake(x, **k):
            n = torch.randn(x.shape, generator=g, dtype=x.dtype)
            rec.append(n)
            return n.clone()

        torch.randn_like = fake
        try:
            img = pipe.sample(c["n"], (8, 32, 32), condition=c.get("cond"), **c["kw"])
        finally:
            torch.randn_like = orig
        out[name] = dict(kw=c["kw"], cond=c.get("cond"), noises=torch.stack(rec), image=img)
        print("sample", name, len(rec), "draws", float(img.abs().max()))
    torch.save(dict(unet_cfg=UNET_SMALL, vae_cfg=VAE_SMALL, sched=SCHED, cases=out),
               os.path.join(OUT, "sample_small.pt"))


def ckpt_fixture():
    """Skeleton of the two Lightning checkpoints the reference's training scripts write (model_base.py:11-14 →
    `save_hyperparameters()`; scripts/train_diffusion.py:117-132, scripts/train_latent_embedder_2d.py:68-90):
    the *pickled* `hyper_parameters` (class objects by qualified reference name, incl. training-side classes that do
    not exist at sampling time) and the state_dict key/shape lists.  Weights are not stored (tests rebuild them from
    medfusion_b200.synthetic per key); scheduler buffers are (they are tiny and computed, not learned)."""
    import lpips
    vae_hp = dict(in_channels=3, out_channels=3, spatial_dims=2, emb_channels=8, hid_chs=[64, 128], kernel_sizes=[3, 3],
                  strides=[1, 2], norm_name=("GROUP", {"num_groups": 8, "affine": True}), act_name=("Swish", {}),
                  dropout=None, use_res_block=True, deep_supervision=False, learnable_interpolation=True,
                  use_attention="none", embedding_loss_weight=1e-6, perceiver=lpips.LPIPS, perceiver_kwargs={},
                  perceptual_loss_weight=1.0, optimizer=torch.optim.Adam, optimizer_kwargs={"lr": 1e-4},
                  lr_scheduler=None, lr_scheduler_kwargs={}, loss=torch.nn.L1Loss, loss_kwargs={"reduction": "none"},
                  sample_every_n_steps=1000)
    vae = VAE(**{k: (dict(v) if isinstance(v, dict) else v) for k, v in vae_hp.items()})
    pipe_hp = dict(noise_scheduler=GaussianNoiseScheduler, noise_estimator=UNet, latent_embedder=VAE,
                   noise_scheduler_kwargs=dict(SCHED),
                   noise_estimator_kwargs=dict(time_embedder=TimeEmbbeding, cond_embedder=LabelEmbedder,
                                               **fresh(UNET_SMALL)),
                   latent_embedder_checkpoint="runs/2022_12_12_133315_chest_vaegan/last_vae.ckpt",
                   estimator_objective="x_T", estimate_variance=False, use_self_conditioning=False,
                   classifier_free_guidance_dropout=0.5, num_samples=4, do_input_centering=False, clip_x0=False,
                   use_ema=False, ema_kwargs={}, optimizer=torch.optim.AdamW, optimizer_kwargs={"lr": 1e-4},
                   lr_scheduler=None, lr_scheduler_kwargs={}, loss=torch.nn.L1Loss, loss_kwargs={},
                   sample_every_n_steps=1000)
    kw = {k: v for k, v in pipe_hp.items() if k not in ("latent_embedder", "latent_embedder_checkpoint")}
    kw["noise_estimator_kwargs"] = dict(kw["noise_estimator_kwargs"])
    pipe = DiffusionPipeline(latent_embedder=None, **kw)
    pipe.latent_embedder = vae

    def keys(m):
        return [(k, tuple(v.shape), str(v.dtype)) for k, v in m.state_dict().items()]

    lightning_extras = {"epoch": 3, "global_step": 1234, "pytorch-lightning_version": "1.8.6", "callbacks": {},
                        "optimizer_states": [], "lr_schedulers": []}
    torch.save(dict(vae=dict(hyper_parameters=vae_hp, keys=keys(vae), **lightning_extras),
                    pipeline=dict(hyper_parameters=pipe_hp, keys=keys(pipe), **lightning_extras),
                    sched_buffers={k: v.clone() for k, v in pipe.noise_scheduler.state_dict().items()}),
               os.path.join(OUT, "ckpt_skeleton.pt"))
    print("ckpt skeleton:", len(keys(vae)), "vae keys,", len(keys(pipe)), "pipeline keys")


@torch.no_grad()
def opts_fixture():
    """The optional paths of DiffusionPipeline.forward (diffusion_pipeline.py:240-273): learned variance, the
    self-conditioning quirk, cold diffusion, x_0 objective with clipping — scheduler level and 3-step trajectories
    (no latent decoder: outputs are latents)."""
    s = GaussianNoiseScheduler(**SCHED)
    g = gen(31)
    x_t = torch.randn(5, 8, 16, 16, generator=g)
    pred = torch.randn(5, 8, 16, 16, generator=g)
    noise = torch.randn(5, 8, 16, 16, generator=g)
    pvar = torch.randn(5, 8, 16, 16, generator=g).clamp(-1.3, 1.3)
    t = torch.tensor([999, 500, 37, 1, 0])
    sched_out = {}
    orig = torch.randn_like
    torch.randn_like = lambda x, **k: noise.clone()
    try:
        for obj, fn in (("x_T", s.estimate_x_t_prior_from_x_T), ("x_0", s.estimate_x_t_prior_from_x_0)):
            for clip in (False, True):
                p, x0 = fn(x_t, t, pred, clip_x0=clip, var_scale=pvar / 2 + 0.5)
                sched_out[f"var_{obj}_clip{int(clip)}"] = dict(prior=p, x_0=x0)
                p, x0 = fn(x_t, t, pred, clip_x0=clip, var_scale=0, cold_diffusion=True)
                sched_out[f"cold_{obj}_clip{int(clip)}"] = dict(prior=p, x_0=x0)
    finally:
        torch.randn_like = orig

    def make_pipe(**over):
        est = dict(time_embedder=TimeEmbbeding, cond_embedder=LabelEmbedder, **fresh(UNET_SMALL))
        kw = dict(noise_scheduler=GaussianNoiseScheduler, noise_estimator=UNet, latent_embedder=None,
                  noise_scheduler_kwargs=dict(SCHED), noise_estimator_kwargs=est, estimator_objective="x_T",
                  estimate_variance=False, use_self_conditioning=False, use_ema=False, do_input_centering=False,
                  clip_x0=False)
        kw.update(over)
        pipe = DiffusionPipeline(**kw).eval()
        fill_(pipe.noise_estimator)
        return pipe

    cases = {
        "learned_var_cond": dict(pipe=dict(estimate_variance=True), n=2, cond=torch.tensor([1, 0]),
                                 kw=dict(steps=3, use_ddim=True, guidance_scale=1.0)),
        "learned_var_ddpm": dict(pipe=dict(estimate_variance=True, clip_x0=True), n=2, kw=dict(steps=3, use_ddim=False)),
        "self_cond_ddim": dict(pipe=dict(use_self_conditioning=True), n=2, kw=dict(steps=3, use_ddim=True)),
        "self_cond_x0_cfg": dict(pipe=dict(use_self_conditioning=True, estimator_objective="x_0", clip_x0=True), n=2,
                                 cond=torch.tensor([0, 1]), kw=dict(steps=3, use_ddim=True, guidance_scale=2.0)),
        "cold_ddim": dict(pipe={}, n=2, kw=dict(steps=3, use_ddim=True, cold_diffusion=True)),
        "cold_ddpm_x0": dict(pipe=dict(estimator_objective="x_0", clip_x0=True), n=2,
                             kw=dict(steps=3, use_ddim=False, cold_diffusion=True)),
    }
    out = {}
    for name, c in cases.items():
        pipe = make_pipe(**c["pipe"])
        gg = gen(200 + len(out))
        rec = []

        def fake(x, **k):
            n = torch.randn(x.shape, generator=gg, dtype=x.dtype)
            rec.append(n)
            return n.clone()

        torch.randn_like = fake
        try:
            lat = pipe.sample(c["n"], (8, 32, 32), condition=c.get("cond"), **c["kw"])
        finally:
            torch.randn_like = orig
        out[name] = dict(pipe=c["pipe"], kw=c["kw"], cond=c.get("cond"), noises=torch.stack(rec), latent=lat,
                         keys=[(k, tuple(v.shape)) for k, v in pipe.noise_estimator.state_dict().items()])
        print("opts", name, len(rec), "draws", float(lat.abs().max()))
    torch.save(dict(unet_cfg=UNET_SMALL, sched=SCHED, x_t=x_t, pred=pred, noise=noise, pvar=pvar, t=t,
                    sched_out=sched_out, cases=out), os.path.join(OUT, "sample_opts.pt"))


@torch.no_grad()
def vae_encode_fixture():
    """VAE.encode / VAE.forward of the unmodified reference (latent_embedders.py:756-790) with the quantizer's
    torch.randn draw recorded; weights are the per-key synthetic ones."""
    m = fill_(VAE(loss=torch.nn.MSELoss, **dict(VAE_CANON, deep_supervision=False)).eval())
    x = torch.randn(2, 3, 64, 64, generator=gen(41)).clamp(-1, 1)
    rec = []
    g = gen(42)
    orig = torch.randn

    def fake(*shape, **k):
        shp = shape[0] if len(shape) == 1 and not isinstance(shape[0], int) else shape
        n = orig(tuple(shp), generator=g)
        rec.append(n)
        return n.clone()

    torch.randn = fake
    try:
        z = m.encode(x)
        out, out_hor, emb_loss = m(x)
    finally:
        torch.randn = orig
    h = m.inc(x)
    for e in m.encoders:
        h = e(h)
    mom = m.out_enc(h)
    keys = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    torch.save(dict(cfg=dict(VAE_CANON, deep_supervision=False), x=x, noise_encode=rec[0], z=z, moments=mom,
                    noise_forward=rec[1], out=out, emb_loss=emb_loss, keys=keys),
               os.path.join(OUT, "vae_encode.pt"))
    print("vae encode", float(z.abs().max()), float(mom.abs().max()), float(emb_loss), len(out_hor))


VQVAE_DEMO = dict(in_channels=3, out_channels=3, emb_channels=4, num_embeddings=8192, spatial_dims=2,
                  hid_chs=[64, 128, 256, 512], embedding_loss_weight=1, beta=1)   # scripts/train_latent_embedder_2d.py:101-110
VQVAE_DEFAULT = dict(in_channels=3, out_channels=3, emb_channels=4, num_embeddings=8192, spatial_dims=2)  # ctor defaults:
#                     hid_chs [32, 64, 128, 256], GroupNorm with 32 groups -> 1, 2, 4, 8 channels per group


@torch.no_grad()
def vqvae_fixture():
    """VQVAE.decode of the unmodified reference (latent_embedders.py:314-320), SURVEY.md section 8 f4:
      demo:    the training script's VQVAE (hid 64..512) at a (4,32,32) latent -> 3x256x256 (streamlit/pages/eye.py:34 shape)
      default: the constructor defaults (hid 32..256, 32 groups) at a (4,64,64) latent -> 3x512x512 (colon.py:36 shape)
    plus the quantiser's own outputs (z_q, code indices)."""
    out = {}
    for name, cfg, shape, seed in (("demo", VQVAE_DEMO, (2, 4, 32, 32), 61), ("default", VQVAE_DEFAULT, (1, 4, 64, 64), 62)):
        m = fill_(VQVAE(loss=torch.nn.L1Loss, perceiver=None, **cfg).eval())
        z = torch.randn(shape, generator=gen(seed))
        x = m.decode(z)
        z_q, _ = m.quantizer(z)
        zf = torch.moveaxis(z, 1, -1).reshape(-1, 4)
        w = m.quantizer.embedder.weight
        dist = (torch.sum(zf ** 2, dim=1, keepdim=True) + torch.sum(w ** 2, dim=1) - 2 * torch.einsum("bd,dn->bn", zf, w.t()))
        idx = torch.argmin(dist, dim=1).view(shape[0], shape[2], shape[3]).to(torch.int32)
        keys = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
        out[name] = dict(cfg=cfg, z=z, x=x, z_q=z_q, idx=idx, keys=keys)
        print("vqvae", name, tuple(x.shape), float(x.abs().max()), int(idx.unique().numel()), "codes used")
    torch.save(out, os.path.join(OUT, "vqvae.pt"))


# /root/reference/tests/models/test_unet.py:13-28, verbatim: 3-channel images, widths 32..256 (1/2/4/8 channels per GroupNorm
# group), a 1x1 stem, 'linear' attention at every level, the default 64-wide time embedding (16-wide sinusoid), the default
# deep supervision (True -> depth-2 heads) and FLOATING-POINT timesteps (`time = torch.randn([1,])`, :34)
UNET_REFTEST = dict(in_ch=3, out_ch=3, spatial_dims=2, hid_chs=[32, 64, 128, 256], kernel_sizes=[1, 3, 3, 3],
                    strides=[1, 2, 2, 2], cond_embedder_kwargs={"emb_dim": 64, "num_classes": 2}, use_attention="linear")


@torch.no_grad()
def reftest_fixture():
    """The model of the reference's own tests/models/test_unet.py (which asserts nothing): both outputs of forward — y and the
    deep-supervision list — for a seeded 64x64 input, float timesteps and labels."""
    m = UNet(cond_embedder=LabelEmbedder, time_embedder_kwargs={"pos_embedder_kwargs": {}},
             **{k: (dict(v) if isinstance(v, dict) else v) for k, v in UNET_REFTEST.items()}).eval()
    fill_(m)
    g = gen(71)
    x = torch.randn(2, 3, 64, 64, generator=g)
    t = torch.randn(2, generator=g)
    c = torch.tensor([0, 1])
    y, y_ver = m(x, t, c)
    keys = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    torch.save(dict(cfg=UNET_REFTEST, x=x, t=t, cond=c, y=y, y_ver=list(y_ver), keys=keys),
               os.path.join(OUT, "unet_reftest.pt"))
    print("reftest", float(y.abs().max()), [tuple(v.shape) for v in y_ver], len(keys))


UNET_CONFIG4 = dict(in_ch=8, out_ch=8, spatial_dims=2, hid_chs=[256, 256, 512, 1024], kernel_sizes=[3, 3, 3, 3],
                    strides=[1, 2, 2, 2], time_embedder_kwargs={"emb_dim": 1024},
                    cond_embedder_kwargs={"emb_dim": 1024, "num_classes": 2}, deep_supervision=False,
                    use_res_block=True, use_attention=["none", "none", "none", "spatial"])


@torch.no_grad()
def config4_fixture():
    """BASELINE.json configs[3] estimator (8x64x64 latent, spatial attention at the deepest level), one forward at B=1;
    same inputs as tests/test_gpu_models.py::test_config4_shapes_attention_at_64x64_latent_vs_oracle."""
    m = make_unet(UNET_CONFIG4)
    g = gen(4)
    x = torch.randn(1, 8, 64, 64, generator=g)
    t = torch.tensor([321])
    c = torch.tensor([1])
    y, _ = m(x, t, c)
    keys = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    torch.save(dict(cfg=UNET_CONFIG4, x=x, t=t, cond=c, y=y, keys=keys), os.path.join(OUT, "unet_config4.pt"))
    print("config4", float(y.abs().max()), len(keys))


@torch.no_grad()
def traj_fixture():
    """Canonical-width 50-step trajectories of the unmodified reference: B=2, injected noise
    (regenerated from a seed by the test: only the seed is stored), latents only (no decoder).  Every x_t the
    estimator saw is recorded through a forward pre-hook, so the GPU test can compare free-running snapshots AND do a
    teacher-forced per-step comparison.
      ddim50_clip_cond: use_ddim=True, steps=50, clip_x0=True (the ctor default, diffusion_pipeline.py:37), 2-class
                        condition with guidance_scale=1.  Without clipping a random-weight estimator makes the DDIM
                        trajectory grow like 1/sqrt(alphas_cumprod) (|x| ~ 6e5 after 50 steps); with the clamp
                        latents stay O(1) and the absolute tolerance is meaningful.
      ddpm50:           use_ddim=False, steps=50 (ancestral steps t=49..0), clip_x0=False, unconditional."""
    cases = {
        "ddim50_clip_cond": dict(pipe=dict(clip_x0=True), kw=dict(steps=50, use_ddim=True, guidance_scale=1.0),
                                 cond=torch.tensor([1, 0]), seed=301),
        "ddpm50": dict(pipe=dict(clip_x0=False), kw=dict(steps=50, use_ddim=False), cond=None, seed=302),
    }
    out = {}
    orig = torch.randn_like
    for name, c in cases.items():
        pipe = DiffusionPipeline(noise_scheduler=GaussianNoiseScheduler, noise_estimator=UNet, latent_embedder=None,
                                 noise_scheduler_kwargs=dict(SCHED),
                                 noise_estimator_kwargs=dict(time_embedder=TimeEmbbeding, cond_embedder=LabelEmbedder,
                                                             **fresh(UNET_CANON)),
                                 estimator_objective="x_T", estimate_variance=False, use_self_conditioning=False,
                                 use_ema=False, do_input_centering=False, **c["pipe"]).eval()
        fill_(pipe.noise_estimator)
        xs, ts = [], []

        def record(_module, args):           # a pre-hook must return None, or its value replaces the inputs
            xs.append(args[0].clone())
            ts.append(args[1].clone())

        pipe.noise_estimator.register_forward_pre_hook(record)

        def run(perturb):
            g = gen(c["seed"])
            n_draws = [0]

            def fake(x, **k):
                n_draws[0] += 1
                n = torch.randn(x.shape, generator=g, dtype=x.dtype)
                if perturb is not None and n_draws[0] == 1:
                    n = n + perturb            # x_T only
                return n

            torch.randn_like = fake
            try:
                lat = pipe.sample(2, (8, 32, 32), condition=c["cond"], **c["kw"])
            finally:
                torch.randn_like = orig
            return lat, n_draws[0]

        lat, n_draws = run(None)
        x_in, t_in = torch.stack(xs), torch.stack([t.reshape(-1)[0] for t in ts])
        # The reference's OWN sensitivity: the same run with x_T moved by 1e-5 * N(0,1) (an input error of the size of the
        # absolute tolerance).  max |x_t' - x_t| per step tells how the trajectory itself amplifies a tolerance-sized
        # error; the GPU test widens the free-running tolerance by a stated multiple of it (and by nothing else).
        xs.clear(); ts.clear()
        lat_p, _ = run(1e-5 * torch.randn(2, 8, 32, 32, generator=gen(999)))
        sens = torch.tensor([float((xs[i] - x_in[i]).abs().max()) for i in range(len(xs))])
        out[name] = dict(pipe=c["pipe"], kw=c["kw"], cond=c["cond"], seed=c["seed"], n_draws=n_draws,
                         x_in=x_in, t_in=t_in, latent=lat, sens=sens, sens_final=float((lat_p - lat).abs().max()),
                         keys=[(k, tuple(v.shape)) for k, v in pipe.noise_estimator.state_dict().items()])
        print("traj", name, n_draws, "draws", [round(float(x.abs().max()), 2) for x in x_in[::10]],
              float(lat.abs().max()), "sens", [f"{float(v):.1e}" for v in sens[::10]], out[name]["sens_final"])
    torch.save(dict(unet_cfg=UNET_CANON, sched=SCHED, cases=out), os.path.join(OUT, "traj_canonical.pt"))


@torch.no_grad()
def forward_fixture():
    """DiffusionPipeline.forward values (diffusion_pipeline.py:232-275): the 4-tuple
    (x_t_prior, x_0, x_T, self_cond) for seeded inputs, the scheduler's randn_like draw replaced by a stored tensor."""
    g = gen(51)
    x_t = torch.randn(3, 8, 32, 32, generator=g)
    noise = torch.randn(3, 8, 32, 32, generator=g)
    t = torch.tensor([700, 250, 0])
    cases = {
        "uncond": dict(pipe={}, call={}),
        "cond_g1": dict(pipe={}, call=dict(condition=torch.tensor([1, 0, 1]))),
        "cfg3_uncond_labels": dict(pipe={}, call=dict(condition=torch.tensor([1, 0, 1]), guidance_scale=3.0,
                                                      un_cond=torch.tensor([0, 1, 0]))),
        "x0_objective_clip": dict(pipe=dict(estimator_objective="x_0", clip_x0=True), call={}),
    }
    out = {}
    orig = torch.randn_like
    torch.randn_like = lambda x, **k: noise.clone()
    try:
        for name, c in cases.items():
            kw = dict(noise_scheduler=GaussianNoiseScheduler, noise_estimator=UNet, latent_embedder=None,
                      noise_scheduler_kwargs=dict(SCHED),
                      noise_estimator_kwargs=dict(time_embedder=TimeEmbbeding, cond_embedder=LabelEmbedder,
                                                  **fresh(UNET_SMALL)),
                      estimator_objective="x_T", estimate_variance=False, use_self_conditioning=False, use_ema=False,
                      do_input_centering=False, clip_x0=False)
            kw.update(c["pipe"])
            pipe = DiffusionPipeline(**kw).eval()
            fill_(pipe.noise_estimator)
            prior, x0, xT, sc = pipe(x_t, t, **c["call"])
            out[name] = dict(pipe=c["pipe"], call=c["call"], x_t_prior=prior, x_0=x0, x_T=xT, self_cond=sc)
            print("forward", name, float(prior.abs().max()), float(x0.abs().max()), float(xT.abs().max()))
    finally:
        torch.randn_like = orig
    torch.save(dict(unet_cfg=UNET_SMALL, sched=SCHED, x_t=x_t, t=t, noise=noise, cases=out),
               os.path.join(OUT, "forward_small.pt"))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    ALL = dict(unet_small=lambda: unet_fixture("unet_small.pt", UNET_SMALL, 1),
               unet_canonical=lambda: unet_fixture("unet_canonical.pt", UNET_CANON, 2),
               unet_attn_small=lambda: unet_fixture("unet_attn_small.pt", UNET_ATTN, 3),
               vae=vae_fixture, sched=sched_fixture, sample=sample_fixture, ckpt=ckpt_fixture, opts=opts_fixture,
               vae_encode=vae_encode_fixture, config4=config4_fixture, vqvae=vqvae_fixture, reftest=reftest_fixture, traj=traj_fixture, forward=forward_fixture)
    todo = sys.argv[1:] or list(ALL)      # python oracle/make_golden.py [fixture names]
    for name in todo:
        ALL[name]()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
