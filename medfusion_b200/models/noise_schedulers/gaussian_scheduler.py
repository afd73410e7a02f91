"""Gaussian (DDPM) noise scheduler — reverse-process half
(reference: medical_diffusion/models/noise_schedulers/gaussian_scheduler.py:7-151, scheduler_base.py:5-46).

Buffers are the reference's (same names, fp64 -> fp32), so state_dicts interchange.  The per-step
arithmetic (x_0 estimate, posterior mean/std, + noise, CFG combine, DDIM-form re-noise) is one fused
elementwise kernel, `mf_sched_step`; the methods below are thin views of it.
"""
from __future__ import annotations

import ctypes
import math

import torch
import torch.nn as nn

from ... import _lib
from ..._engine import cuda_stream_ptr, require_cuda

_TABLES = ("sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_mean_coef1",
           "posterior_mean_coef2", "posterior_variance", "betas", "alphas_cumprod")


def make_betas(schedule_strategy, timesteps, beta_start, beta_end):
    """Closed-form beta schedules in fp64 (gaussian_scheduler.py:22-36)."""
    f64 = torch.float64
    if schedule_strategy == "linear":
        return torch.linspace(beta_start, beta_end, timesteps, dtype=f64)
    if schedule_strategy == "scaled_linear":
        return torch.linspace(beta_start ** 0.5, beta_end ** 0.5, timesteps, dtype=f64) ** 2
    if schedule_strategy == "cosine":
        s = 0.008
        x = torch.linspace(0, timesteps, timesteps + 1, dtype=f64)
        ac = torch.cos(((x / timesteps) + s) / (1 + s) * torch.pi * 0.5) ** 2
        ac = ac / ac[0]
        return torch.clip(1 - (ac[1:] / ac[:-1]), 0, 0.999)
    raise NotImplementedError(f"{schedule_strategy} does is not implemented for GaussianNoiseScheduler")


class BasicNoiseScheduler(nn.Module):
    def __init__(self, timesteps=1000, T=None):
        super().__init__()
        self.timesteps = timesteps
        self.T = timesteps if T is None else T
        self.register_buffer("timesteps_array", torch.linspace(0, self.T - 1, self.timesteps, dtype=torch.long))

    @staticmethod
    def extract(x, t, ndim):
        return x.gather(0, t).reshape(-1, *((1,) * (ndim - 1)))


class GaussianNoiseScheduler(BasicNoiseScheduler):
    def __init__(self, timesteps=1000, T=None, schedule_strategy="cosine", beta_start=0.0001, beta_end=0.02,
                 betas=None):
        super().__init__(timesteps, T)
        self.schedule_strategy = schedule_strategy
        b = torch.as_tensor(betas, dtype=torch.float64) if betas is not None else make_betas(
            schedule_strategy, timesteps, beta_start, beta_end)
        a = 1 - b
        ac = torch.cumprod(a, dim=0)
        ac_prev = torch.cat([torch.ones(1, dtype=torch.float64), ac[:-1]])
        tables = {
            "betas": b, "alphas": a, "alphas_cumprod": ac, "alphas_cumprod_prev": ac_prev,
            "sqrt_alphas_cumprod": ac.sqrt(), "sqrt_one_minus_alphas_cumprod": (1. - ac).sqrt(),
            "sqrt_recip_alphas_cumprod": (1. / ac).sqrt(), "sqrt_recipm1_alphas_cumprod": (1. / ac - 1).sqrt(),
            "posterior_mean_coef1": b * ac_prev.sqrt() / (1. - ac),
            "posterior_mean_coef2": (1. - ac_prev) * a.sqrt() / (1. - ac),
            "posterior_variance": b * (1. - ac_prev) / (1. - ac),
        }
        for name, val in tables.items():  # registration order == gaussian_scheduler.py:46-58
            self.register_buffer(name, val.to(torch.float32))

    # ------------------------------------------------------------------------------------------
    def _tables(self):
        tab = _lib.SchedTables()
        for n in _TABLES:
            buf = getattr(self, n)
            require_cuda(buf, f"scheduler buffer {n}")
            setattr(tab, n, buf.data_ptr())
        return tab

    def step(self, x_t, t, pred, *, pred_uncond=None, guidance_scale=1.0, noise=None, t_next=None, noise_ddim=None,
             objective="x_T", clip_x0=True, want=("x_prior", "x_0", "x_T"), learned_variance=False,
             cold_diffusion=False):
        """One fused reverse step. Returns dict with the requested tensors among x_prior, x_0, x_T, x_next.

        learned_variance=True: `pred` (and `pred_uncond`) are the estimator's [B, 2*C, ...] outputs; the second channel
        half is the variance interpolation coefficient (diffusion_pipeline.py:246-256).  cold_diffusion=True: the
        deterministic update of gaussian_scheduler.py:88-93 (no noise is used)."""
        require_cuda(x_t, "scheduler step")
        x_t = x_t.contiguous()
        pred = pred.contiguous()
        B = x_t.shape[0]
        chw = x_t[0].numel()
        if pred[0].numel() != (2 * chw if learned_variance else chw):
            raise ValueError(f"pred has {pred[0].numel()} elements per sample, expected {(2 if learned_variance else 1) * chw}")
        t = t.to(device=x_t.device, dtype=torch.int64).expand(B).contiguous()
        outs = {k: torch.empty_like(x_t) for k in want}
        if t_next is not None:
            t_next = t_next.to(device=x_t.device, dtype=torch.int64).reshape(1).contiguous()
        tab = self._tables()
        if pred_uncond is not None:
            pred_uncond = pred_uncond.contiguous()

        def ptr(v):
            return None if v is None else v.contiguous().data_ptr()

        opts = None
        if learned_variance or cold_diffusion:
            opts = _lib.SchedOpts()
            if learned_variance:
                opts.d_pred_var = pred.data_ptr() + 4 * chw
                opts.d_pred_var_uncond = None if pred_uncond is None else pred_uncond.data_ptr() + 4 * chw
                opts.pred_batch_stride = 2 * chw
            if cold_diffusion:
                opts.cold_diffusion = 1
                opts.sqrt_alphas_cumprod = self.sqrt_alphas_cumprod.data_ptr()
                opts.sqrt_one_minus_alphas_cumprod = self.sqrt_one_minus_alphas_cumprod.data_ptr()
                opts.T = int(self.T)
        with torch.cuda.device(x_t.device):
            _lib.check(_lib.load().mf_sched_step_opts(
                ctypes.byref(tab), x_t.data_ptr(), pred.data_ptr(), ptr(pred_uncond), float(guidance_scale),
                t.data_ptr(), ptr(noise), ptr(t_next), ptr(noise_ddim), 1 if objective == "x_0" else 0,
                1 if clip_x0 else 0, ptr(outs.get("x_prior")), ptr(outs.get("x_0")), ptr(outs.get("x_T")),
                ptr(outs.get("x_next")), B, chw, None if opts is None else ctypes.byref(opts),
                cuda_stream_ptr(x_t.device)), "mf_sched_step_opts")
        return outs

    # --- reference-named views (gaussian_scheduler.py:80-151) ---------------------------------
    def _prior(self, x_t, t, pred, objective, use_log, clip_x0, var_scale, cold_diffusion):
        self._check_step_opts(use_log)
        learned = torch.is_tensor(var_scale)
        if not learned and var_scale != 0:
            raise NotImplementedError("a scalar var_scale != 0 is not implemented (pass the per-element tensor)")
        if learned:   # var_scale = v/2 + 0.5 in the pipeline; the kernel takes v stacked behind the prediction
            v = (var_scale.to(pred.dtype).expand_as(pred) - 0.5) * 2
            pred = torch.cat([pred, v], dim=1)
        noise = None if cold_diffusion else self.x_final(x_t)            # gaussian_scheduler.py:99 (not drawn when cold)
        o = self.step(x_t, t, pred, noise=noise, objective=objective, clip_x0=clip_x0, want=("x_prior", "x_0"),
                      learned_variance=learned, cold_diffusion=cold_diffusion)
        return o["x_prior"], o["x_0"]

    def estimate_x_t_prior_from_x_T(self, x_t, t, x_T, use_log=True, clip_x0=True, var_scale=0, cold_diffusion=False):
        return self._prior(x_t, t, x_T, "x_T", use_log, clip_x0, var_scale, cold_diffusion)

    def estimate_x_t_prior_from_x_0(self, x_t, t, x_0, use_log=True, clip_x0=True, var_scale=0, cold_diffusion=False):
        return self._prior(x_t, t, x_0, "x_0", use_log, clip_x0, var_scale, cold_diffusion)

    def estimate_x_0(self, x_t, x_T, t, clip_x0=True):
        return self.step(x_t, t, x_T, objective="x_T", clip_x0=clip_x0, want=("x_0",))["x_0"]

    def estimate_x_T(self, x_t, x_0, t, clip_x0=True):
        return self.step(x_t, t, x_0, objective="x_0", clip_x0=clip_x0, want=("x_T",))["x_T"]

    def estimate_mean_t(self, x_t, x_0, t):
        # posterior mean == x_prior with zero noise and no clipping of the supplied x_0
        return self.step(x_t, t, x_0, objective="x_0", clip_x0=False, want=("x_prior",))["x_prior"]

    def estimate_variance_t(self, t, ndim, log=True, var_scale=0, eps=1e-20):
        lo = self.extract(self.posterior_variance, t, ndim)
        hi = self.extract(self.betas, t, ndim)
        if log:
            lo, hi = torch.log(lo.clamp(min=eps)), torch.log(hi.clamp(min=eps))
        return var_scale * hi + (1 - var_scale) * lo

    @staticmethod
    def _check_step_opts(use_log):
        if not use_log:
            raise NotImplementedError("use_log=False is not implemented")

    @classmethod
    def x_final(cls, x):
        return torch.randn_like(x)

    @classmethod
    def _clip_x_0(cls, x_0):
        return x_0.clamp(-1, 1)
