"""UNet noise estimator — drop-in for the reference's live `unet2.UNet`
(/root/reference/medical_diffusion/models/estimators/unet2.py:15-269; exported by estimators/__init__.py:1).

Same constructor keywords, same state_dict keys, same `forward(x_t, t, condition, self_cond) -> (y, y_ver)`
contract; the arithmetic is a launch plan of sm_100a kernels inside libmedfusion_b200.so
(mf_unet_forward).  Attention blocks ('linear' / 'spatial', attention_blocks.py) are supported; options of the reference that
are outside the sampling hot path raise NotImplementedError instead of silently diverging.
"""
from __future__ import annotations

import ctypes

import torch

from ... import _lib
from ..._engine import EngineModule, cuda_stream_ptr, on_device, require_cuda
from ..embedders import TimeEmbbeding


def _name_of(spec):
    return str(spec[0] if isinstance(spec, (tuple, list)) else spec).lower()


class UNet(EngineModule):
    _prefix = "mf_unet"

    def __init__(
        self,
        in_ch=1,
        out_ch=1,
        spatial_dims=3,
        hid_chs=(256, 256, 512, 1024),
        kernel_sizes=(3, 3, 3, 3),
        strides=(1, 2, 2, 2),
        act_name=("SWISH", {}),
        norm_name=("GROUP", {"num_groups": 32, "affine": True}),
        time_embedder=TimeEmbbeding,
        time_embedder_kwargs=None,
        cond_embedder=None,
        cond_embedder_kwargs=None,
        deep_supervision=True,
        use_res_block=True,
        estimate_variance=False,
        use_self_conditioning=False,
        dropout=0.0,
        learnable_interpolation=True,
        use_attention="none",
        num_res_blocks=2,
    ):
        super().__init__()
        if spatial_dims != 2:
            raise NotImplementedError("medfusion_b200.UNet implements the 2-D model (spatial_dims=2)")
        if not use_res_block:
            raise NotImplementedError("use_res_block=False (UnetBasicBlock) is not on the accelerated hot path")
        if not learnable_interpolation:
            raise NotImplementedError("learnable_interpolation=False (pooling) is not implemented")
        if dropout not in (0, 0.0, None):
            raise NotImplementedError("dropout is identity at inference; construct with dropout=0.0")
        if _name_of(act_name) != "swish" or _name_of(norm_name) != "group":
            raise NotImplementedError("only Swish + GroupNorm are implemented")
        attn = list(use_attention) if isinstance(use_attention, (list, tuple)) else [use_attention] * len(strides)
        attn_codes = {"none": 0, "linear": 1, "spatial": 2}
        if any(a not in attn_codes for a in attn) or len(attn) != len(strides):
            raise ValueError("use_attention entries must be 'none', 'linear' or 'spatial' (one per level)")
        if any(a != "none" for a in attn) and time_embedder is None:
            raise NotImplementedError("attention blocks without a time embedder are not implemented")
        depth = len(strides)
        if not (len(hid_chs) == len(kernel_sizes) == depth) or depth > _lib.MF_MAX_LEVELS:
            raise ValueError("hid_chs, kernel_sizes and strides must have equal length <= 8")

        self.use_self_conditioning = use_self_conditioning
        self.use_res_block = use_res_block
        self.depth = depth
        self.num_res_blocks = num_res_blocks
        self.in_ch, self.out_ch = in_ch, out_ch
        self.estimate_variance = estimate_variance
        self._hid0, self._k0, self._s0 = int(hid_chs[0]), int(kernel_sizes[0]), int(strides[0])
        self._has_attention = any(a != "none" for a in attn)

        self.time_spec = time_embedder(**dict(time_embedder_kwargs or {})) if time_embedder is not None else None
        self.cond_spec = cond_embedder(**dict(cond_embedder_kwargs or {})) if cond_embedder is not None else None
        if self.cond_spec is not None and self.time_spec is None:
            raise NotImplementedError("a condition embedder without a time embedder is not supported")
        if self.cond_spec is not None and self.cond_spec.emb_dim != self.time_spec.emb_dim:
            raise ValueError("cond_embedder emb_dim must equal the time embedding dim")

        cfg = _lib.UNetConfig()
        cfg.in_ch = in_ch * 2 if use_self_conditioning else in_ch                       # unet2.py:65
        cfg.out_ch, cfg.depth = (out_ch * 2 if estimate_variance else out_ch), depth      # unet2.py:212
        for i in range(depth):
            cfg.hid_chs[i], cfg.kernel_sizes[i], cfg.strides[i] = hid_chs[i], kernel_sizes[i], strides[i]
            cfg.attention[i] = attn_codes[attn[i]]
        cfg.num_res_blocks = num_res_blocks
        cfg.emb_dim = self.time_spec.emb_dim if self.time_spec is not None else 0
        cfg.pos_emb_dim = self.time_spec.pos_emb_dim if self.time_spec is not None else 0
        cfg.num_classes = self.cond_spec.num_classes if self.cond_spec is not None else 0
        cfg.norm_groups = dict(norm_name[1]).get("num_groups", 32) if isinstance(norm_name, (tuple, list)) else 32
        # deep supervision (unet2.py:214-217): True = depth-2 heads on the concat inputs of levels 2.., plain out_ch each
        self.deep_supervision = (depth - 2 if deep_supervision else 0) if isinstance(deep_supervision, bool) \
            else int(deep_supervision)
        cfg.deep_supervision = self.deep_supervision
        cfg.ds_out_ch = out_ch
        self._strides = [int(v) for v in strides]
        handle = ctypes.c_void_p()
        _lib.check(_lib.load().mf_unet_create(ctypes.byref(cfg), ctypes.byref(handle)), "mf_unet_create")
        # zero-initialised modules of the reference: 2nd conv of every res block (conv_blocks.py:336 -> :174)
        # and the output head (unet2.py:213)
        # ... plus every attention output projection (attention_blocks.py:149-152)
        self._engine_init(handle, zero_init=("*.block_seq.1.basic_block.conv.", "outc.", "outc_ver.", "*.to_out.0."))

    # ------------------------------------------------------------------------------------------
    def _after_param_sync(self, stream):
        if self.time_spec is not None:
            freqs = self.time_spec.pos_embedder.frequencies().to(self.device, torch.float32).contiguous()
            self._freqs = freqs  # keep alive until the async copy has been enqueued/ordered
            _lib.check(_lib.load().mf_unet_set_time_freqs(self._h, freqs.data_ptr(), freqs.numel(), stream),
                       "set_time_freqs")

    def forward(self, x_t, t=None, condition=None, self_cond=None):
        """x_t [B,C,H,W] fp32, t [B] int64, condition [B] int64 | None  ->  (y [B,out,H,W], [])  (unet2.py:222-269)"""
        require_cuda(x_t, "UNet.forward(x_t)")
        if x_t.dim() != 4 or x_t.shape[1] != self.in_ch:
            raise ValueError(f"x_t must be [B,{self.in_ch},H,W], got {tuple(x_t.shape)}")
        if self.time_spec is not None and t is None:
            raise NotImplementedError("t=None with a time embedder is not supported")
        self.sync_params()
        self._check_inputs(x_t, t, condition)
        B, _, H, W = x_t.shape
        x = x_t.contiguous().float()
        if self.use_self_conditioning:
            # unet2.py:243-246, quirk kept: the second half is zeros on the first call and x_t ITSELF (not the
            # self_cond tensor) whenever a self_cond is passed
            x = torch.cat([x, torch.zeros_like(x) if self_cond is None else x], dim=1)
        t_float = t is not None and torch.is_floating_point(t)      # the reference's sinusoid takes any dtype
        if t is None:
            tt = None
        elif t_float:
            tt = t.to(device=x.device, dtype=torch.float32).expand(B).contiguous()
        else:
            tt = t.to(device=x.device, dtype=torch.int64).expand(B).contiguous()
        cc = None
        if condition is not None and self.cond_spec is not None:
            cc = condition.to(device=x.device, dtype=torch.int64).contiguous()
        y = torch.empty((B, self.out_ch * (2 if self.estimate_variance else 1), H, W), device=x.device,
                        dtype=torch.float32)
        if not t_float and self.deep_supervision == 0:
            self._forward_into(x, tt, cc, y)
            return y, []
        # deep-supervision outputs (unet2.py:258-269): head k sits at the resolution of level k+1
        y_ver, h, w = [], H, W
        k0, s0 = self._k0, self._s0
        h, w = (h + 2 * (k0 // 2) - k0) // s0 + 1, (w + 2 * (k0 // 2) - k0) // s0 + 1
        for lvl in range(1, self.deep_supervision + 1):
            st = self._strides[lvl]
            h, w = (h + 2 - 3) // st + 1, (w + 2 - 3) // st + 1
            y_ver.append(torch.empty((B, self.out_ch, h, w), device=x.device, dtype=torch.float32))
        ptrs = (ctypes.c_void_p * max(1, len(y_ver)))(*[v.data_ptr() for v in y_ver])
        with on_device(x):
            ws, ws_bytes = self._workspace(B, H, W)
            _lib.check(_lib.load().mf_unet_forward_ex(
                self._h, x.data_ptr(), None if (tt is None or t_float) else tt.data_ptr(),
                tt.data_ptr() if t_float else None, None if cc is None else cc.data_ptr(), y.data_ptr(), ptrs, len(y_ver),
                B, H, W, ws, ws_bytes, cuda_stream_ptr(x.device)), "mf_unet_forward_ex")
        return y, y_ver

    def _check_inputs(self, x_t, t, condition):
        """The kernels index the embedding table with `condition` and the sinusoid / scheduler tables with `t`: out-of-range
        values would read out of bounds where nn.Embedding / gather raise in the reference.  One host check per public
        call (skipped while a CUDA graph is being captured: the captured values were checked by the warm-up call)."""
        if x_t.device != self.device:
            raise RuntimeError(f"input on {x_t.device}, module on {self.device}")
        if torch.cuda.is_current_stream_capturing():
            return
        if condition is not None and self.cond_spec is not None:
            lo, hi = int(condition.min()), int(condition.max())
            if lo < 0 or hi >= self.cond_spec.num_classes:
                raise IndexError(f"condition values must be in [0, {self.cond_spec.num_classes}), got [{lo}, {hi}]")
        if t is not None and torch.is_tensor(t) and not torch.is_floating_point(t) and t.numel() > 0 and int(t.min()) < 0:
            raise IndexError("timesteps must be >= 0")     # (fp32 timesteps only feed the sinusoid: any value is fine)

    def supports_fused_step(self):
        """Mirror of the C-side requirements of mf_unet_forward_step (narrow 1x1 head with the scheduler update in its
        epilogue): out_ch <= 8, first level a multiple of 64 channels, odd stem kernel with stride 1, no learned
        variance, x_t and the estimate of the same channel count."""
        return (self.out_ch <= 8 and not self.estimate_variance and not self.use_self_conditioning
                and self._hid0 % 64 == 0 and self._k0 % 2 == 1 and self._s0 == 1 and self.in_ch == self.out_ch)

    def _forward_into(self, x, t, cond, y):
        """Hot-loop entry: contiguous CUDA tensors of the right dtype, no checks, no parameter sync."""
        B, _, H, W = x.shape
        with on_device(x):
            ws, ws_bytes = self._workspace(B, H, W)
            _lib.check(_lib.load().mf_unet_forward(self._h, x.data_ptr(), None if t is None else t.data_ptr(),
                                                   None if cond is None else cond.data_ptr(), y.data_ptr(), B, H, W, ws,
                                                   ws_bytes, cuda_stream_ptr(x.device)), "mf_unet_forward")

    def forward_step(self, x_t, t, condition, scheduler, *, pred_uncond=None, guidance_scale=1.0, noise=None,
                     t_next=None, noise_ddim=None, objective="x_T", clip_x0=True, want=("x_next",), want_pred=False,
                     uniform_t=False):
        """UNet.forward with `scheduler`'s reverse step fused into the output head's epilogue (mf_unet_forward_step).

        Returns a dict with the requested tensors among x_prior, x_0, x_T, x_next (+ 'pred' if want_pred).
        uniform_t=True promises that all entries of t are equal (true in the sampling loop): the time/label embedding
        MLP is then evaluated once per class instead of once per sample."""
        require_cuda(x_t, "UNet.forward_step(x_t)")
        self.sync_params()
        B, _, H, W = x_t.shape
        x = x_t.contiguous().float()
        tt = t.to(device=x.device, dtype=torch.int64).expand(B).contiguous()
        cc = None
        if condition is not None and self.cond_spec is not None:
            cc = condition.to(device=x.device, dtype=torch.int64).contiguous()
        outs = {k: torch.empty_like(x) for k in want}
        pred = torch.empty_like(x) if want_pred else None
        if t_next is not None:
            t_next = t_next.to(device=x.device, dtype=torch.int64).reshape(1).contiguous()
        tab = scheduler._tables()
        keep = [v.contiguous() if v is not None else None for v in (pred_uncond, noise, noise_ddim)]

        def ptr(v):
            return None if v is None else v.data_ptr()

        args = _lib.StepArgs(ctypes.pointer(tab), ptr(keep[0]), float(guidance_scale), ptr(keep[1]), ptr(t_next),
                             ptr(keep[2]), 1 if objective == "x_0" else 0, 1 if clip_x0 else 0,
                             ptr(outs.get("x_prior")), ptr(outs.get("x_0")), ptr(outs.get("x_T")), ptr(outs.get("x_next")),
                             1 if uniform_t else 0)
        with on_device(x):
            ws, ws_bytes = self._workspace(B, H, W)
            _lib.check(_lib.load().mf_unet_forward_step(self._h, x.data_ptr(), tt.data_ptr(), ptr(cc), ptr(pred), B, H, W,
                                                        ws, ws_bytes, ctypes.byref(args), cuda_stream_ptr(x.device)),
                       "mf_unet_forward_step")
        if want_pred:
            outs["pred"] = pred
        return outs

    def supports_cfg_batch(self):
        """mf_unet_forward_step_cfg: classifier-free guidance as one 2B batch (label embedder, no attention blocks)."""
        return (self.supports_fused_step() and self.cond_spec is not None and self.in_ch < 64
                and not getattr(self, "_has_attention", False))

    def cfg_labels(self, condition, un_cond):
        """int64 [2B] label vector of the one-batch CFG step: unconditional half first; `un_cond=None` (no label at all,
        diffusion_pipeline.py:241 passes condition=None) is encoded as the extra all-zero row `num_classes`."""
        cond = condition.to(device=self.device, dtype=torch.int64).reshape(-1)
        if un_cond is None:
            unc = torch.full_like(cond, self.cond_spec.num_classes)
        else:
            unc = un_cond.to(device=self.device, dtype=torch.int64).reshape(-1)
        return torch.cat([unc, cond]).contiguous()

    def forward_step_cfg(self, x_t, t, cond2, scheduler, *, guidance_scale, noise=None, t_next=None, noise_ddim=None,
                         objective="x_T", clip_x0=True, want=("x_next",)):
        """One reverse step with classifier-free guidance, both estimator passes of diffusion_pipeline.py:240-244 run as a
        single 2B batch (mf_unet_forward_step_cfg).  `cond2` from cfg_labels(); all entries of t are equal."""
        require_cuda(x_t, "UNet.forward_step_cfg(x_t)")
        self.sync_params()
        B, _, H, W = x_t.shape
        x = x_t.contiguous().float()
        tt = t.to(device=x.device, dtype=torch.int64).expand(B).contiguous()
        outs = {k: torch.empty_like(x) for k in want}
        if t_next is not None:
            t_next = t_next.to(device=x.device, dtype=torch.int64).reshape(1).contiguous()
        tab = scheduler._tables()
        keep = [v.contiguous() if v is not None else None for v in (noise, noise_ddim)]

        def ptr(v):
            return None if v is None else v.data_ptr()

        args = _lib.StepArgs(ctypes.pointer(tab), None, float(guidance_scale), ptr(keep[0]), ptr(t_next), ptr(keep[1]),
                             1 if objective == "x_0" else 0, 1 if clip_x0 else 0, ptr(outs.get("x_prior")),
                             ptr(outs.get("x_0")), ptr(outs.get("x_T")), ptr(outs.get("x_next")), 1)
        with on_device(x):
            ws, ws_bytes = self._workspace(2 * B, H, W)
            _lib.check(_lib.load().mf_unet_forward_step_cfg(self._h, x.data_ptr(), tt.data_ptr(), cond2.data_ptr(), B, H, W,
                                                            ws, ws_bytes, ctypes.byref(args), cuda_stream_ptr(x.device)),
                       "mf_unet_forward_step_cfg")
        return outs

    def profile(self, x_t, t, condition=None):
        """Per-launch device times of one forward: list of (ms, kind, algorithmic_flops); kind 0 = tcgen05 conv."""
        self.sync_params()
        B, _, H, W = x_t.shape
        x = x_t.contiguous().float()
        tt = t.to(device=x.device, dtype=torch.int64).expand(B).contiguous()
        cc = None if condition is None else condition.to(device=x.device, dtype=torch.int64).contiguous()
        y = torch.empty((B, self.out_ch, H, W), device=x.device, dtype=torch.float32)
        ws, ws_bytes = self._workspace(B, H, W)
        return self._profile_call("mf_unet_profile", (x.data_ptr(), tt.data_ptr(), None if cc is None else cc.data_ptr(),
                                                      y.data_ptr(), B, H, W), (ws, ws_bytes))
