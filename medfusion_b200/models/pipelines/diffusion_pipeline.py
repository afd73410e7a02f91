"""DiffusionPipeline — sampling half, drop-in for
/root/reference/medical_diffusion/models/pipelines/diffusion_pipeline.py (forward :232-275, denoise :278-310,
sample :312-317).

Semantics kept exactly: `use_ddim=True` default with eta == 1 (two RNG draws per step, the
scheduler draw being discarded except on the last step), partial DDPM schedule for `use_ddim=False, steps=k`,
CFG with two estimator passes, noise drawn with `torch.randn_like` in the reference's order so that a given
`torch.manual_seed` yields the reference's trajectory.  What changes is where the arithmetic runs: the UNet,
the scheduler update and the VAE decoder are sm_100a launch plans behind the C ABI, and the loop issues no
host synchronisation and no progress-bar calls.

Multi-GPU (not in the reference): `sample(..., shard=True)` under torch.distributed partitions
the batch across ranks with no per-step communication and all-gathers the decoded images once at the end.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ..._engine import require_cuda
from ...checkpoint import CheckpointMixin


class _StepGraph:
    """One reverse-diffusion timestep captured as a CUDA graph (estimator pass(es) + noise draws + fused update).

    Static buffers: x (updated in place by every replay), t [B], t_next [1], condition / un_cond.  The launch plan
    of the estimator is built by the warm-up call, so the capture contains kernel launches only; the torch RNG draws
    inside use the graph-safe generator offsets, so the noise stream continues exactly as in eager mode.
    """

    def __init__(self, pipe, est, x, condition, un_cond, guidance_scale, ddim, noise_fn):
        self.x = x.clone()
        B = x.shape[0]
        self.t = torch.zeros(B, dtype=torch.int64, device=x.device)
        self.t_next = torch.zeros(1, dtype=torch.int64, device=x.device)
        self.cond = None if condition is None else condition.clone()
        self.un_cond = None if un_cond is None else un_cond.clone()
        cfg = (condition is not None) and (guidance_scale != 1.0)
        sched = pipe.noise_scheduler
        one_batch = cfg and pipe.cfg_single_batch and est.supports_cfg_batch()
        self.cond2 = est.cfg_labels(condition, un_cond) if one_batch else None

        def body():
            if one_batch:     # both estimator passes of diffusion_pipeline.py:240-244 as ONE 2B batch
                noise = noise_fn(self.x)
                noise2 = noise_fn(self.x) if ddim else None
                o = est.forward_step_cfg(self.x, self.t, self.cond2, sched, guidance_scale=guidance_scale, noise=noise,
                                         t_next=self.t_next if ddim else None, noise_ddim=noise2,
                                         objective=pipe.estimator_objective, clip_x0=pipe.clip_x0, want=("x_next",))
                self.x.copy_(o["x_next"])
                return
            pred_u = est(self.x, self.t, condition=self.un_cond, self_cond=None)[0] if cfg else None
            noise = noise_fn(self.x)
            noise2 = noise_fn(self.x) if ddim else None
            o = est.forward_step(self.x, self.t, self.cond, sched, pred_uncond=pred_u, guidance_scale=guidance_scale,
                                 noise=noise, t_next=self.t_next if ddim else None, noise_ddim=noise2,
                                 objective=pipe.estimator_objective, clip_x0=pipe.clip_x0, want=("x_next",),
                                 uniform_t=True)
            self.x.copy_(o["x_next"])

        # warm-up on a side stream (builds the launch plan, primes allocations) without disturbing the noise stream
        rng = torch.cuda.get_rng_state(x.device)
        side = torch.cuda.Stream(device=x.device)
        side.wait_stream(torch.cuda.current_stream(x.device))
        with torch.cuda.stream(side):
            body()
        torch.cuda.current_stream(x.device).wait_stream(side)
        torch.cuda.set_rng_state(rng, x.device)
        self.graph = torch.cuda.CUDAGraph()
        # thread-local capture mode: host threads of the caller (e.g. the PNG writers of sample_dataset, which wait on
        # CUDA events while the next chunk is being sampled) must not invalidate this capture
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            body()
        # the captured launches point into the estimator's workspace for this shape and at its prepared weights:
        # keep the former alive with the graph, remember the version of the latter
        self._workspace = est._workspace_tensor(2 * B if one_batch else B, x.shape[2], x.shape[3])
        self._workspace2 = est._workspace_tensor(B, x.shape[2], x.shape[3]) if (cfg and not one_batch) else None
        self.param_sig = est._synced_sig
        torch.cuda.set_rng_state(rng, x.device)   # capture does not draw, but keep the contract explicit

    def replay(self):
        self.graph.replay()


class _EMAWeights(nn.Module):
    """Holder with the reference's attribute name (`EMAModel.averaged_model`, utils/train_utils.py:33)."""

    def __init__(self, model):
        super().__init__()
        self.averaged_model = model
        for p in self.averaged_model.parameters():
            p.requires_grad = False


class DiffusionPipeline(CheckpointMixin, nn.Module):
    def __init__(
        self,
        noise_scheduler,
        noise_estimator,
        latent_embedder=None,
        noise_scheduler_kwargs=None,
        noise_estimator_kwargs=None,
        latent_embedder_checkpoint="",
        estimator_objective="x_T",
        estimate_variance=False,
        use_self_conditioning=False,
        classifier_free_guidance_dropout=0.5,
        num_samples=4,
        do_input_centering=True,
        clip_x0=True,
        use_ema=False,
        ema_kwargs=None,
        optimizer=None,
        optimizer_kwargs=None,
        lr_scheduler=None,
        lr_scheduler_kwargs=None,
        loss=None,
        loss_kwargs=None,
        sample_every_n_steps=1000,
    ):
        super().__init__()
        if estimator_objective not in ("x_T", "x_0"):
            raise ValueError("Unknown Objective")
        est_kwargs = dict(noise_estimator_kwargs or {})
        est_kwargs["estimate_variance"] = estimate_variance          # diffusion_pipeline.py:50-51
        est_kwargs["use_self_conditioning"] = use_self_conditioning
        self.noise_scheduler = noise_scheduler(**dict(noise_scheduler_kwargs or {}))
        self.noise_estimator = noise_estimator(**est_kwargs)
        if latent_embedder is not None:
            self.latent_embedder = latent_embedder.load_from_checkpoint(latent_embedder_checkpoint)
            for p in self.latent_embedder.parameters():
                p.requires_grad = False
        else:
            self.latent_embedder = None
        self.estimator_objective = estimator_objective
        self.use_self_conditioning = use_self_conditioning
        self.num_samples = num_samples
        self.classifier_free_guidance_dropout = classifier_free_guidance_dropout
        self.do_input_centering = do_input_centering
        self.estimate_variance = estimate_variance
        self.clip_x0 = clip_x0
        self.use_ema = use_ema
        self.use_cuda_graph = True    # capture one timestep as a CUDA graph inside denoise() (medfusion_b200 extension)
        # after every denoise(): read the sticky saturation counter of the kernels (one stream sync per call) and raise
        # FloatingPointError if a value left the fp16 range of the split planes — the reference is fp32-range, so the
        # result would silently differ from it
        self.check_saturation = True
        # classifier-free guidance: run the unconditional and the conditional estimator pass as ONE 2B batch (per-sample
        # label index; un_cond=None -> an all-zero "no label" row) instead of the reference's two B passes — fills the GPU
        # at the small batches of scripts/sample.py (B=16, guidance 8).  False: two passes, like the reference.
        self.cfg_single_batch = True
        self._step_graphs = {}
        if use_ema:
            # weight selection only (diffusion_pipeline.py:234-237): a second estimator holding the averaged weights under
            # the reference's key prefix `ema_model.averaged_model.*`; the EMA *update* is training-side (train_utils.py).
            self.ema_model = _EMAWeights(noise_estimator(**est_kwargs))

    @property
    def device(self):
        return self.noise_scheduler.betas.device

    def load_state_dict(self, state_dict, strict=True, **kw):
        """Accepts the reference pipeline's full state_dict (model_base.py:77-83 / Lightning): training-side entries
        (`loss_fct.*`, the VAE's encoder / discriminator / LPIPS tensors, `ema_model.*` when use_ema is off) are not
        part of the sampling path and are skipped; everything on the path must match exactly when strict."""
        from ..embedders.latent_embedders import _ENCODER_SIDE_PREFIXES
        drop = ("loss_fct.",) + (() if self.use_ema else ("ema_model.",))
        drop += tuple("latent_embedder." + p for p in _ENCODER_SIDE_PREFIXES)
        if self.latent_embedder is None:
            drop += ("latent_embedder.",)
        own = {k: v for k, v in state_dict.items() if not k.startswith(drop)}
        return super().load_state_dict(own, strict=strict, **kw)

    # ---------------------------------------------------------------------------------------------
    def _predict(self, x_t, t, condition, guidance_scale, un_cond, self_cond=None):
        """Estimator pass(es); returns (pred, pred_uncond|None). CFG combine happens in the step kernel."""
        est = self.ema_model.averaged_model if self.use_ema else self.noise_estimator
        if (condition is not None) and (guidance_scale != 1.0):
            pred_uncond, _ = est(x_t, t, condition=un_cond, self_cond=self_cond)
            pred_cond, _ = est(x_t, t, condition=condition, self_cond=self_cond)
            return pred_cond, pred_uncond
        pred, _ = est(x_t, t, condition=condition, self_cond=self_cond)
        return pred, None

    def forward(self, x_t, t, condition=None, self_cond=None, guidance_scale=1.0, cold_diffusion=False, un_cond=None):
        """-> (x_t_prior, x_0, x_T, self_cond)   (diffusion_pipeline.py:232-275)"""
        require_cuda(x_t, "DiffusionPipeline.forward(x_t)")
        pred, pred_u = self._predict(x_t, t, condition, guidance_scale, un_cond, self_cond)
        noise = None if cold_diffusion else self.noise_scheduler.x_final(x_t)   # not drawn on the cold path
        o = self.noise_scheduler.step(x_t, t, pred, pred_uncond=pred_u, guidance_scale=guidance_scale, noise=noise,
                                      objective=self.estimator_objective, clip_x0=self.clip_x0,
                                      want=("x_prior", "x_0", "x_T"), learned_variance=self.estimate_variance,
                                      cold_diffusion=cold_diffusion)
        self_cond_out = o["x_T"] if self.estimator_objective == "x_0" else o["x_0"]
        return o["x_prior"], o["x_0"], o["x_T"], self_cond_out

    @torch.no_grad()
    def denoise(self, x_t, steps=None, condition=None, use_ddim=True, **kwargs):
        """Reverse loop + latent decode (diffusion_pipeline.py:278-310)."""
        custom_noise = kwargs.pop("_noise_fn", None)
        graph_ok = kwargs.pop("_cuda_graph", self.use_cuda_graph)
        as_uint8 = kwargs.pop("_uint8", False)
        unknown = set(kwargs) - {"guidance_scale", "un_cond", "cold_diffusion"}
        if unknown:  # the reference forwards **kwargs to forward(), which raises TypeError on anything else
            raise TypeError(f"forward() got an unexpected keyword argument '{sorted(unknown)[0]}'")
        cold = bool(kwargs.get("cold_diffusion", False))
        guidance_scale = kwargs.get("guidance_scale", 1.0)
        un_cond = kwargs.get("un_cond", None)
        require_cuda(x_t, "DiffusionPipeline.denoise(x_t)")
        est0 = self.ema_model.averaged_model if self.use_ema else self.noise_estimator
        if hasattr(est0, "_check_inputs"):      # label range, once per call (the loop itself never synchronises)
            est0._check_inputs(x_t, None, condition)
            if un_cond is not None:
                est0._check_inputs(x_t, None, un_cond)
        check_sat = self.check_saturation and not torch.cuda.is_current_stream_capturing()
        if check_sat:
            from ... import saturation_count
            saturation_count(reset=True, device=x_t.device)
        torch.cuda.nvtx.range_push("medfusion_b200.denoise")          # NVTX ranges for nsys / ncu timelines
        try:
            out = self._denoise(x_t, steps, condition, use_ddim, custom_noise, graph_ok, as_uint8, cold, guidance_scale,
                                un_cond, kwargs)
        finally:
            torch.cuda.nvtx.range_pop()
        if check_sat:
            n_sat = saturation_count(reset=True, device=x_t.device)
            if n_sat:
                raise FloatingPointError(
                    f"medfusion_b200: {n_sat} value(s) exceeded the fp16 range (+-65504) of the split activation planes "
                    "during denoise(); the reference computes in fp32 range, so this result would not match it "
                    "(set pipeline.check_saturation = False to get the clamped result)")
        return out

    def _denoise(self, x_t, steps, condition, use_ddim, custom_noise, graph_ok, as_uint8, cold, guidance_scale, un_cond,
                 kwargs):
        sched = self.noise_scheduler
        noise_fn = custom_noise or self.noise_scheduler.x_final
        if use_ddim:
            steps = sched.timesteps if steps is None else steps
            timesteps_array = torch.linspace(0, sched.T - 1, steps, dtype=torch.long, device=x_t.device)
        else:
            timesteps_array = sched.timesteps_array[slice(0, steps)]
            steps = len(timesteps_array)
        B = x_t.shape[0]
        x_t = x_t.contiguous().float()
        ts = timesteps_array.flip(0)
        est = self.ema_model.averaged_model if self.use_ema else self.noise_estimator
        # the fused head (+ CUDA graph) covers the canonical configuration; learned variance, self-conditioning and cold
        # diffusion take the estimator + mf_sched_step_opts path below
        fused = (hasattr(est, "forward_step") and est.supports_fused_step() and not self.use_self_conditioning
                 and not cold and x_t.shape[1] == est.in_ch)
        cfg = (condition is not None) and (guidance_scale != 1.0)
        # CUDA-graph path: one captured timestep replayed `steps` times (a second capture without the DDIM re-noise
        # for the last step).  Host-side noise injection (tests) cannot be captured -> eager loop below.
        if fused and graph_ok and (custom_noise is None or getattr(custom_noise, "graph_safe", False)) and steps > 2:
            def get_graph(ddim):
                # a noise function may carry a stable `cache_key` (the sharded one does: (lo, hi, full shape)), so that
                # repeated sample(shard=True) calls reuse the captured graphs instead of re-capturing
                nkey = getattr(custom_noise, "cache_key", id(custom_noise))
                key = (id(est), tuple(x_t.shape), condition is not None, un_cond is not None, float(guidance_scale),
                       ddim, self.estimator_objective, self.clip_x0, nkey, self.cfg_single_batch)
                g = self._step_graphs.get(key)
                est.sync_params()
                if g is not None and g.param_sig != est._synced_sig:   # weights changed since the capture
                    g = None
                if g is None:
                    if len(self._step_graphs) >= 4:
                        self._step_graphs.clear()
                    g = _StepGraph(self, est, x_t, condition, un_cond, guidance_scale, ddim, noise_fn)
                    self._step_graphs[key] = g
                if condition is not None:
                    g.cond.copy_(condition)
                if un_cond is not None:
                    g.un_cond.copy_(un_cond)
                if g.cond2 is not None:
                    g.cond2.copy_(est.cfg_labels(condition, un_cond))
                return g

            n_main = steps - 1 if use_ddim else steps        # the last DDIM step has no re-noise
            try:
                g = get_graph(use_ddim)
                g2 = get_graph(False) if use_ddim else None
            except RuntimeError as exc:                      # capture refused (e.g. an outer capture is active):
                import warnings                              # same kernels, just launched one by one
                warnings.warn(f"medfusion_b200: CUDA-graph capture of the timestep failed ({exc}); running eagerly")
                self.use_cuda_graph = False
                return self._denoise(x_t, steps, condition, use_ddim, custom_noise, False, as_uint8, cold,
                                     guidance_scale, un_cond, kwargs)
            g.x.copy_(x_t)
            torch.cuda.nvtx.range_push(f"medfusion_b200.timesteps[{steps}] (CUDA graph replays)")
            for i in range(n_main):
                g.t.copy_(ts[i].expand(B))
                if use_ddim:
                    g.t_next.copy_(timesteps_array[steps - i - 2].reshape(1))
                g.replay()
            x_t = g.x
            if use_ddim:
                g2.x.copy_(x_t)
                g2.t.copy_(ts[steps - 1].expand(B))
                g2.replay()
                x_t = g2.x
            torch.cuda.nvtx.range_pop()
            return self._decode(x_t.clone(), as_uint8)
        for i in range(steps):
            t = ts[i]
            tb = t.expand(B)
            ddim = use_ddim and (steps - i - 1 > 0)
            t_next = timesteps_array[steps - i - 2] if ddim else None
            if fused and cfg and self.cfg_single_batch and est.supports_cfg_batch():
                if i == 0:
                    cond2 = est.cfg_labels(condition, un_cond)
                noise = noise_fn(x_t)
                noise2 = noise_fn(x_t) if ddim else None
                o = est.forward_step_cfg(x_t, tb, cond2, sched, guidance_scale=guidance_scale, noise=noise, t_next=t_next,
                                         noise_ddim=noise2, objective=self.estimator_objective, clip_x0=self.clip_x0,
                                         want=("x_next",))
            elif fused:
                # estimator pass(es) first (as in the reference), then the draws; the last estimator pass carries the
                # CFG combine + scheduler update (+ DDIM re-noise) in the epilogue of its output head
                pred_u = est(x_t, tb, condition=un_cond, self_cond=None)[0] if cfg else None
                noise = noise_fn(x_t)                         # scheduler draw (gaussian_scheduler.py:99)
                noise2 = noise_fn(x_t) if ddim else None      # DDIM draw (diffusion_pipeline.py:303)
                # NOTE: the draws happen before the (asynchronous) estimator launch but consume the generator in the
                # reference's order; their values do not depend on the estimator.
                o = est.forward_step(x_t, tb, condition, sched, pred_uncond=pred_u, guidance_scale=guidance_scale,
                                     noise=noise, t_next=t_next, noise_ddim=noise2,
                                     objective=self.estimator_objective, clip_x0=self.clip_x0, want=("x_next",),
                                     uniform_t=True)
            else:
                # self-conditioning (diffusion_pipeline.py:291-292 + unet2.py:243-246): None on the first step, then
                # "some tensor" — the estimator only looks at whether it is None
                sc = x_t if (self.use_self_conditioning and i > 0) else None
                pred, pred_u = self._predict(x_t, tb, condition, guidance_scale, un_cond, sc)
                noise = None if cold else noise_fn(x_t)
                noise2 = noise_fn(x_t) if ddim else None
                o = sched.step(x_t, tb, pred, pred_uncond=pred_u, guidance_scale=guidance_scale, noise=noise,
                               t_next=t_next, noise_ddim=noise2, objective=self.estimator_objective,
                               clip_x0=self.clip_x0, want=("x_next",), learned_variance=self.estimate_variance,
                               cold_diffusion=cold)
            x_t = o["x_next"]
        return self._decode(x_t, as_uint8)

    def _decode(self, x_t, as_uint8=False):
        torch.cuda.nvtx.range_push("medfusion_b200.decode")
        try:
            return self._decode_impl(x_t, as_uint8)
        finally:
            torch.cuda.nvtx.range_pop()

    def _decode_impl(self, x_t, as_uint8=False):
        if as_uint8:
            if self.latent_embedder is None or not hasattr(self.latent_embedder, "decode_uint8"):
                raise RuntimeError("sample_uint8 needs a latent embedder with decode_uint8 (medfusion_b200 VAE)")
            return self.latent_embedder.decode_uint8(x_t)
        if self.latent_embedder is not None:
            x_t = self.latent_embedder.decode(x_t)
        return x_t

    @torch.no_grad()
    def sample(self, num_samples, img_size, condition=None, shard=False, **kwargs):
        """x_T ~ N(0, I) -> denoise -> images (diffusion_pipeline.py:312-317).

        shard=True (torch.distributed initialised): every rank draws the FULL-batch noise stream (so seeds match
        the single-GPU result), keeps its contiguous slice of the batch, and the decoded images are all-gathered
        once at the end — no per-step collective.
        """
        template = torch.zeros((num_samples, *img_size), device=self.device)
        if shard:
            from ...parallel import sharded_sample
            return sharded_sample(self, template, condition, **kwargs)
        x_T = self.noise_scheduler.x_final(template)
        return self.denoise(x_T, condition=condition, **kwargs)

    @torch.no_grad()
    def sample_uint8(self, num_samples, img_size, condition=None, **kwargs):
        """sample() with the `clip(-1,1) -> (x+1)/2*255 -> HWC -> uint8` conversion of the bulk generator
        (scripts/helpers/sample_dataset.py:47-50) fused into the decoder's output head: uint8 [B, H, W, C] on device.
        Same noise stream as sample() (medfusion_b200 extension; used by medfusion_b200.sample_dataset)."""
        template = torch.zeros((num_samples, *img_size), device=self.device)
        x_T = self.noise_scheduler.x_final(template)
        return self.denoise(x_T, condition=condition, _uint8=True, **kwargs)

    def interpolate(self, *a, **k):
        raise NotImplementedError("interpolate() needs the forward diffusion (training side); out of scope")
