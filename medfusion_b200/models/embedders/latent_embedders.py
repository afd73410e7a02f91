"""VAE latent embedder (reference: medical_diffusion/models/embedders/latent_embedders.py:620-855).

`VAE.decode` (:764-769) is on the sampling hot path and runs as an sm_100a launch plan (mf_vae_decode); `VAE.encode`
(:756-762, the step on the other side of the latent) is a second plan of the same handle
(mf_vae_encode).  Losses / perceptual nets / GAN variants are training-side and outside this package's scope;
`load_state_dict` accepts a full reference VAE state_dict and ignores those entries.
"""
from __future__ import annotations

import ctypes

import torch

from ... import _lib
from ..._engine import EngineModule, cuda_stream_ptr, on_device, require_cuda
from ...checkpoint import CheckpointMixin

# training-side entries of a reference VAE state_dict (deep-supervision heads, LPIPS, loss modules)
_ENCODER_SIDE_PREFIXES = ("outc_ver.", "perceiver.", "loss_fct.", "quantizer.")


class VAE(CheckpointMixin, EngineModule):
    _prefix = "mf_vae"

    def __init__(
        self,
        in_channels=3,
        out_channels=3,
        spatial_dims=2,
        emb_channels=4,
        hid_chs=(64, 128, 256, 512),
        kernel_sizes=(3, 3, 3, 3),
        strides=(1, 2, 2, 2),
        norm_name=("GROUP", {"num_groups": 8, "affine": True}),
        act_name=("Swish", {}),
        dropout=None,
        use_res_block=True,
        deep_supervision=False,
        learnable_interpolation=True,
        use_attention="none",
        embedding_loss_weight=1e-6,
        perceiver=None,
        perceiver_kwargs=None,
        perceptual_loss_weight=1.0,
        optimizer=None,
        optimizer_kwargs=None,
        lr_scheduler=None,
        lr_scheduler_kwargs=None,
        loss=None,
        loss_kwargs=None,
        sample_every_n_steps=1000,
    ):
        super().__init__()
        if spatial_dims != 2:
            raise NotImplementedError("medfusion_b200.VAE implements the 2-D decoder (spatial_dims=2)")
        if not use_res_block or not learnable_interpolation:
            raise NotImplementedError("only use_res_block=True, learnable_interpolation=True is implemented")
        attn = list(use_attention) if isinstance(use_attention, (list, tuple)) else [use_attention] * len(strides)
        if any(a != "none" for a in attn):
            raise NotImplementedError("VAE attention is not implemented")
        if any(k != 3 for k in kernel_sizes[1:]):
            raise NotImplementedError("decoder res-blocks use 3x3 convolutions")
        depth = len(strides)
        self.depth = depth
        self.emb_channels, self.out_channels = emb_channels, out_channels
        self.in_channels = in_channels
        self.up_factor = 1
        for s in strides[1:]:
            self.up_factor *= s
        cfg = _lib.VAEConfig()
        cfg.emb_channels, cfg.out_channels, cfg.depth = emb_channels, out_channels, depth
        cfg.in_channels = in_channels
        for i in range(depth):
            cfg.hid_chs[i], cfg.strides[i] = hid_chs[i], strides[i]
        cfg.norm_groups = dict(norm_name[1]).get("num_groups", 8) if isinstance(norm_name, (tuple, list)) else 8
        handle = ctypes.c_void_p()
        _lib.check(_lib.load().mf_vae_create(ctypes.byref(cfg), ctypes.byref(handle)), "mf_vae_create")
        # zero-init: 2nd conv of each res block (conv_blocks.py:336) and the image head (latent_embedders.py:743)
        self._engine_init(handle, zero_init=("*.block_seq.1.basic_block.conv.", "outc."))
        self._enc_ws = {}

    # --- checkpoint compatibility -------------------------------------------------------------
    def load_state_dict(self, state_dict, strict=True, **kw):
        own = {k: v for k, v in state_dict.items() if not k.startswith(_ENCODER_SIDE_PREFIXES)}
        return super().load_state_dict(own, strict=strict, **kw)

    # --- hot path -------------------------------------------------------------------------------
    def decode(self, z):
        """z [B,emb_channels,h,w] -> x [B,out_channels,h*f,w*f]  (latent_embedders.py:764-769)"""
        require_cuda(z, "VAE.decode(z)")
        if z.dim() != 4 or z.shape[1] != self.emb_channels:
            raise ValueError(f"z must be [B,{self.emb_channels},h,w], got {tuple(z.shape)}")
        self.sync_params()
        B, _, H, W = z.shape
        zc = z.contiguous().float()
        x = torch.empty((B, self.out_channels, H * self.up_factor, W * self.up_factor), device=z.device,
                        dtype=torch.float32)
        with on_device(zc):
            ws, ws_bytes = self._workspace(B, H, W)
            _lib.check(_lib.load().mf_vae_decode(self._h, zc.data_ptr(), x.data_ptr(), B, H, W, ws, ws_bytes,
                                                 cuda_stream_ptr(zc.device)), "mf_vae_decode")
        return x

    def decode_uint8(self, z, also_float=False):
        """decode + `clip(-1,1) -> (x+1)/2*255 -> HWC -> uint8` (scripts/helpers/sample_dataset.py:47-50) fused into the
        output head: returns uint8 [B, H, W, C] (and the fp32 NCHW images too if also_float)."""
        require_cuda(z, "VAE.decode_uint8(z)")
        self.sync_params()
        B, _, H, W = z.shape
        zc = z.contiguous().float()
        Ho, Wo = H * self.up_factor, W * self.up_factor
        img = torch.empty((B, Ho, Wo, self.out_channels), device=z.device, dtype=torch.uint8)
        x = torch.empty((B, self.out_channels, Ho, Wo), device=z.device, dtype=torch.float32) if also_float else None
        with on_device(zc):
            ws, ws_bytes = self._workspace(B, H, W)
            _lib.check(_lib.load().mf_vae_decode_u8(self._h, zc.data_ptr(), None if x is None else x.data_ptr(),
                                                    img.data_ptr(), B, H, W, ws, ws_bytes, cuda_stream_ptr(zc.device)),
                       "mf_vae_decode_u8")
        return (img, x) if also_float else img

    def profile(self, z):
        """Per-launch device times of one decode: list of (ms, kind, algorithmic_flops)."""
        self.sync_params()
        B, _, H, W = z.shape
        zc = z.contiguous().float()
        x = torch.empty((B, self.out_channels, H * self.up_factor, W * self.up_factor), device=z.device,
                        dtype=torch.float32)
        ws, ws_bytes = self._workspace(B, H, W)
        return self._profile_call("mf_vae_profile", (zc.data_ptr(), x.data_ptr(), B, H, W), (ws, ws_bytes))

    def _encode(self, x, sample=True, want_moments=False):
        require_cuda(x, "VAE.encode(x)")
        if x.dim() != 4 or x.shape[1] != self.in_channels:
            raise ValueError(f"x must be [B,{self.in_channels},H,W], got {tuple(x.shape)}")
        self.sync_params()
        B, _, H, W = x.shape
        f = self.up_factor
        if H % f or W % f:
            raise ValueError(f"image height/width must be multiples of {f}")
        xc = x.contiguous().float()
        z = torch.empty((B, self.emb_channels, H // f, W // f), device=x.device, dtype=torch.float32)
        # DiagonalGaussianDistribution.forward (latent_embedders.py:26): torch.randn(mean.shape, device=x.device)
        noise = torch.randn(z.shape, device=x.device) if sample else None
        mom = torch.empty((B, 2 * self.emb_channels, H // f, W // f), device=x.device) if want_moments else None
        key = ("enc", B, H, W, x.device.index)
        ws = self._enc_ws.get(key)
        if ws is None:
            with on_device(xc):
                nbytes = _lib.load().mf_vae_encode_workspace_bytes(self._h, B, H, W)
            if nbytes == 0:
                _lib.check(2, "mf_vae_encode_workspace_bytes")
            self._enc_ws.clear()
            ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=x.device)
            self._enc_ws[key] = ws
        off = (-ws.data_ptr()) % 1024
        with on_device(xc):
            _lib.check(_lib.load().mf_vae_encode(self._h, xc.data_ptr(), None if noise is None else noise.data_ptr(),
                                                 z.data_ptr(), None if mom is None else mom.data_ptr(), B, H, W,
                                                 ws.data_ptr() + off, ws.numel() - off, cuda_stream_ptr(xc.device)),
                       "mf_vae_encode")
        return z, mom

    def encode(self, x):
        """x [B,in_channels,H,W] -> z [B,emb_channels,H/f,W/f], one reparameterised sample (latent_embedders.py:756-762)"""
        return self._encode(x)[0]

    def forward(self, x_in):
        """-> (out, out_hor, emb_loss) as latent_embedders.py:771-790 (deep supervision heads are not built: out_hor = [])"""
        z, mom = self._encode(x_in, want_moments=True)
        mean, logvar = mom.chunk(2, dim=1)
        logvar = logvar.clamp(-30.0, 20.0)
        kl = 0.5 * torch.sum(mean.pow(2) + logvar.exp() - 1.0 - logvar) / x_in.shape[0]       # :29-31
        return self.decode(z), [], kl


class VQVAE(CheckpointMixin, EngineModule):
    """VQVAE latent embedder, decode half (reference: latent_embedders.py:180-320; SURVEY.md section 8 f4 — the embedder
    of the colon / eye demos, streamlit/pages/colon.py:36 with (4, 64, 64) latents, eye.py:34 with (4, 32, 32)).

    `decode(z)` = quantizer(z) -> inc_dec -> decoders -> outc exactly as :314-320: the latent is first snapped to the
    nearest codebook row (VectorQuantizer.forward :50-69, including its z + (z_q - z) evaluation), then decoded by the same
    launch plan as VAE.decode.  `encode` / `forward` (training side) are not built; `load_state_dict` accepts a full
    reference VQVAE / VQGAN state_dict and skips the encoder-side entries."""
    _prefix = "mf_vae"
    _SKIP = ("inc.", "encoders.", "out_enc.", "outc_ver.", "perceiver.", "loss_fct.")

    def __init__(
        self,
        in_channels=3,
        out_channels=3,
        spatial_dims=2,
        emb_channels=4,
        num_embeddings=8192,
        hid_chs=(32, 64, 128, 256),
        kernel_sizes=(3, 3, 3, 3),
        strides=(1, 2, 2, 2),
        norm_name=("GROUP", {"num_groups": 32, "affine": True}),
        act_name=("Swish", {}),
        dropout=0.0,
        use_res_block=True,
        deep_supervision=False,
        learnable_interpolation=True,
        use_attention="none",
        beta=0.25,
        embedding_loss_weight=1.0,
        perceiver=None,
        perceiver_kwargs=None,
        perceptual_loss_weight=1.0,
        optimizer=None,
        optimizer_kwargs=None,
        lr_scheduler=None,
        lr_scheduler_kwargs=None,
        loss=None,
        loss_kwargs=None,
        sample_every_n_steps=1000,
    ):
        super().__init__()
        if spatial_dims != 2:
            raise NotImplementedError("medfusion_b200.VQVAE implements the 2-D decoder (spatial_dims=2)")
        if not use_res_block or not learnable_interpolation:
            raise NotImplementedError("only use_res_block=True, learnable_interpolation=True is implemented")
        attn = list(use_attention) if isinstance(use_attention, (list, tuple)) else [use_attention] * len(strides)
        if any(a != "none" for a in attn):
            raise NotImplementedError("VQVAE attention is not implemented")
        if any(k != 3 for k in kernel_sizes[1:]):
            raise NotImplementedError("decoder res-blocks use 3x3 convolutions")
        depth = len(strides)
        self.depth = depth
        self.emb_channels, self.out_channels, self.in_channels = emb_channels, out_channels, in_channels
        self.num_embeddings = num_embeddings
        self.up_factor = 1
        for s in strides[1:]:
            self.up_factor *= s
        cfg = _lib.VAEConfig()
        cfg.emb_channels, cfg.out_channels, cfg.depth = emb_channels, out_channels, depth
        cfg.in_channels = 0                      # decoder-only handle
        cfg.num_embeddings = num_embeddings
        for i in range(depth):
            cfg.hid_chs[i], cfg.strides[i] = hid_chs[i], strides[i]
        cfg.norm_groups = dict(norm_name[1]).get("num_groups", 32) if isinstance(norm_name, (tuple, list)) else 32
        handle = ctypes.c_void_p()
        _lib.check(_lib.load().mf_vae_create(ctypes.byref(cfg), ctypes.byref(handle)), "mf_vae_create")
        self._engine_init(handle, zero_init=("*.block_seq.1.basic_block.conv.", "outc."))
        with torch.no_grad():                    # VectorQuantizer init (:47): U(-1/K, 1/K)
            self.quantizer.embedder.weight.uniform_(-1.0 / num_embeddings, 1.0 / num_embeddings)

    def load_state_dict(self, state_dict, strict=True, **kw):
        own = {k: v for k, v in state_dict.items() if not k.startswith(self._SKIP)}
        return super().load_state_dict(own, strict=strict, **kw)

    decode = VAE.decode
    decode_uint8 = VAE.decode_uint8
    profile = VAE.profile

    def quantize(self, z):
        """VectorQuantizer.forward's first output (z_q with the straight-through evaluation) and the code indices."""
        require_cuda(z, "VQVAE.quantize(z)")
        self.sync_params()
        B, C = z.shape[:2]
        zc = z.contiguous().float()
        zq = torch.empty_like(zc)
        idx = torch.empty((B,) + tuple(z.shape[2:]), device=z.device, dtype=torch.int32)
        with on_device(zc):
            _lib.check(_lib.load().mf_op_vq_quantize(zc.data_ptr(), self.quantizer.embedder.weight.data_ptr(),
                                                     zq.data_ptr(), idx.data_ptr(), B, C, zc[0, 0].numel(),
                                                     self.num_embeddings, cuda_stream_ptr(zc.device)), "mf_op_vq_quantize")
        return zq, idx

    def encode(self, x):
        raise NotImplementedError("VQVAE.encode is training-side (SURVEY.md section 8 f4 names decode only)")

    def forward(self, x_in):
        raise NotImplementedError("VQVAE.forward (autoencoding + losses) is training-side")
