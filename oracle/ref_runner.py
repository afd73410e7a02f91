"""BENCH / TEST INFRASTRUCTURE — NOT product code.  Times the reference's own implementation of the hot path.

`kind == "reference"`: the UNMODIFIED reference package, staged by `__graft_entry__.build()` from /root/reference into
the git-ignored `oracle/_ref/medical_diffusion/` (it travels to the GPU box with the gpurun snapshot; /root/reference
does not exist there), imported through the third-party stand-ins of `oracle/shims/` (monai, pytorch_lightning,
streamlit, lpips, pytorch_msssim are not installable offline).
`kind == "port"`: fallback when the staged copy is absent — the CPU oracle restatement (oracle/medfusion_oracle.py) with
state_dict key lists read from the committed fixtures (never from the product engine).

Only bench.py's reference arm / cpu_baseline / gpu_comparator legs and tests may import this module.
"""
from __future__ import annotations

import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_DIR = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "medical_diffusion", "models", "pipelines", "diffusion_pipeline.py"))


def stage(src="/root/reference") -> bool:
    """Copy the reference's python package (sources only, unmodified) into oracle/_ref/.  Build container only."""
    import shutil
    pkg = os.path.join(src, "medical_diffusion")
    if not os.path.isdir(pkg):
        return available()
    dst = os.path.join(REF_DIR, "medical_diffusion")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    shutil.copytree(pkg, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    return True


def _import_reference():
    for p in (REF_DIR, os.path.join(HERE, "shims"), ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch  # noqa: F401
    from medical_diffusion.models.embedders import LabelEmbedder, TimeEmbbeding
    from medical_diffusion.models.embedders.latent_embedders import VAE
    from medical_diffusion.models.estimators import UNet
    from medical_diffusion.models.noise_schedulers import GaussianNoiseScheduler
    from medical_diffusion.models.pipelines import DiffusionPipeline
    import medical_diffusion.models.pipelines.diffusion_pipeline as dp     # (the package itself has no __init__.py)
    assert os.path.realpath(dp.__file__).startswith(os.path.realpath(REF_DIR)), \
        "medical_diffusion did not resolve to the staged reference"
    return dict(UNet=UNet, VAE=VAE, TimeEmbbeding=TimeEmbbeding, LabelEmbedder=LabelEmbedder,
                GaussianNoiseScheduler=GaussianNoiseScheduler, DiffusionPipeline=DiffusionPipeline)


def _fresh(cfg):
    kw = {k: (dict(v) if isinstance(v, dict) else v) for k, v in cfg.items()}
    # the reference mutates a mutable default (time_embedder.py:57,63): hand it a fresh dict
    kw["time_embedder_kwargs"] = dict(kw["time_embedder_kwargs"], pos_embedder_kwargs={})
    return kw


class ReferenceRunner:
    """Builds the reference pipeline for one bench configuration and times bounded pieces of its sampling loop."""

    def __init__(self, unet_cfg, vae_cfg, sched, clip_x0, device="cpu", threads=None):
        import torch
        from medfusion_b200.synthetic import fill_      # per-key seeded weights (RNG recipe only, no compute)
        self.torch = torch
        self.device = torch.device(device)
        if threads:
            torch.set_num_threads(int(threads))
        if self.device.type == "cuda":
            # the fp32 comparator the survey asked for: cuDNN / cuBLAS with TF32 disabled
            torch.backends.cudnn.allow_tf32 = False
            torch.backends.cuda.matmul.allow_tf32 = False
        self.kind = "reference" if available() else "port"
        self.unet_cfg, self.vae_cfg, self.sched_cfg, self.clip_x0 = unet_cfg, vae_cfg, sched, clip_x0
        if self.kind == "reference":
            R = _import_reference()
            pipe = R["DiffusionPipeline"](
                noise_scheduler=R["GaussianNoiseScheduler"], noise_estimator=R["UNet"], latent_embedder=None,
                noise_scheduler_kwargs=dict(sched),
                noise_estimator_kwargs=dict(time_embedder=R["TimeEmbbeding"], cond_embedder=R["LabelEmbedder"],
                                            **_fresh(unet_cfg)),
                estimator_objective="x_T", estimate_variance=False, use_self_conditioning=False, use_ema=False,
                do_input_centering=False, clip_x0=clip_x0).eval()
            fill_(pipe.noise_estimator)
            self.vae = fill_(R["VAE"](loss=torch.nn.MSELoss, **vae_cfg).eval()).to(self.device)
            self.pipe = pipe.to(self.device)
        else:
            sys.path.insert(0, HERE)
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import medfusion_oracle as O
            from medfusion_b200.synthetic import synth_tensor
            from util import load_golden, unet_oracle_cfg, vae_oracle_cfg
            self.O = O
            attn = unet_cfg.get("use_attention", "none")
            fixture = "unet_config4.pt" if isinstance(attn, (list, tuple)) and "spatial" in attn else "unet_canonical.pt"
            ukeys = load_golden(fixture)["keys"]
            vkeys = [(k, s) for k, s in load_golden("vae_canonical.pt")["keys"]]
            self.usd = {k: synth_tensor(k, s).to(self.device) for k, s in ukeys}
            self.vsd = {k: synth_tensor(k, s).to(self.device) for k, s in vkeys}
            self.ucfg, self.vcfg = unet_oracle_cfg(unet_cfg), vae_oracle_cfg(vae_cfg)
            self.tabs = O.scheduler_tables(sched["timesteps"], sched["schedule_strategy"], sched["beta_start"],
                                           sched["beta_end"])

    def _sync(self):
        if self.device.type == "cuda":
            self.torch.cuda.synchronize(self.device)

    def time_timesteps(self, B, latent, n, *, use_ddim, conditional, guidance_scale):
        """Wall time of n reverse timesteps at batch B through the reference's own loop (no decode) -> seconds."""
        torch = self.torch
        g = torch.Generator().manual_seed(0)
        x_T = torch.randn(B, *latent, generator=g).to(self.device)
        cond = (torch.arange(B) % 2).to(self.device) if conditional else None
        with torch.no_grad():
            self._sync()
            t0 = time.perf_counter()
            if self.kind == "reference":
                kw = dict(steps=n, use_ddim=use_ddim)
                if conditional:
                    kw["guidance_scale"] = guidance_scale
                self.pipe.denoise(x_T, condition=cond, **kw)          # latent_embedder is None -> returns the latent
            else:
                O = self.O
                noises = [torch.randn(B, *latent, generator=g).to(self.device) for _ in range(2 * n + 1)]
                O.denoise(lambda xx, tt, cc, sc=None: O.unet_forward(self.usd, self.ucfg, xx, tt, cc), self.tabs, x_T,
                          noises, n, use_ddim=use_ddim, guidance_scale=guidance_scale, cond=cond, clip_x0=self.clip_x0)
            self._sync()
            return time.perf_counter() - t0

    def time_decode(self, B, latent):
        torch = self.torch
        z = torch.randn(B, *latent, generator=torch.Generator().manual_seed(1)).to(self.device)
        with torch.no_grad():
            self._sync()
            t0 = time.perf_counter()
            if self.kind == "reference":
                self.vae.decode(z)
            else:
                self.O.vae_decode(self.vsd, self.vcfg, z)
            self._sync()
            return time.perf_counter() - t0
