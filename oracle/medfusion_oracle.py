"""CPU ORACLE — test infrastructure, NOT product code.

A plain-PyTorch (CPU, fp32) restatement of the reference's sampling hot path, written as pure
functions over a reference-format state_dict.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this file; the product package never does.

The arithmetic of the reference lives in torch ops (ATen/oneDNN), so the restatement uses the same
functional ops (F.conv2d, F.group_norm, F.linear, F.interpolate); every function cites the reference
lines it follows (paths relative to /root/reference/medical_diffusion/).  Pinned by tests/golden/*:
fixtures produced by running the UNMODIFIED reference in the build container (oracle/make_golden.py)
— the reference's own tests contain no assertions or golden vectors, so this is the only
pin available.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def swish(x):
    # monai Swish: x * sigmoid(alpha*x), alpha = 1 (models/utils/conv_blocks.py:191 via get_act_layer)
    return x * torch.sigmoid(x)


# ------------------------------------------------------------------------------------------------
# conv blocks
# ------------------------------------------------------------------------------------------------
def conv2d(sd, pre, x, stride=1):
    w = sd[pre + ".weight"]
    k = w.shape[-1]
    # padding = get_padding(k, stride) = (k - stride + 1) // 2   (conv_blocks.py:168, :47)
    return F.conv2d(x, w, sd[pre + ".bias"], stride=stride, padding=(k - stride + 1) // 2)


def basic_block(sd, pre, x, groups):
    """BasicBlock.forward: conv -> GroupNorm -> (Dropout p=0) -> Swish   (models/utils/conv_blocks.py:184-192)"""
    out = conv2d(sd, pre + ".conv", x)
    if pre + ".norm.weight" in sd:
        out = F.group_norm(out, groups, sd[pre + ".norm.weight"], sd[pre + ".norm.bias"], eps=1e-5)
        out = swish(out)
    return out


def basic_res_block(sd, pre, x, groups):
    """BasicResBlock.forward: basic_block(x) + (conv_res(x) | x)   (models/utils/conv_blocks.py:236-240)"""
    out = basic_block(sd, pre + ".basic_block", x, groups)
    res = conv2d(sd, pre + ".conv_res", x) if pre + ".conv_res.weight" in sd else x
    return out + res


def unet_res_block(sd, pre, x, emb, groups):
    """UnetResBlock.forward (models/utils/conv_blocks.py:347-364): emb added after the first half only."""
    e = None
    if emb is not None and pre + ".local_embedder.1.weight" in sd:
        e = F.linear(swish(emb), sd[pre + ".local_embedder.1.weight"], sd[pre + ".local_embedder.1.bias"])
        e = e[:, :, None, None]
    x = basic_res_block(sd, pre + ".block_seq.0", x, groups)
    if e is not None:
        x = x + e
    x = basic_res_block(sd, pre + ".block_seq.1", x, groups)
    return x


def basic_up(sd, pre, x, factor=2):
    """BasicUp.forward: F.interpolate(nearest-exact, x2) then conv3x3 s1 p1  (models/utils/conv_blocks.py:121-131)"""
    if factor != 1:
        x = F.interpolate(x, size=(x.shape[2] * factor, x.shape[3] * factor), mode="nearest-exact")
    return conv2d(sd, pre, x)


# ------------------------------------------------------------------------------------------------
# embedders
# ------------------------------------------------------------------------------------------------
def sinusoidal(t, dim, max_period=10000, shift=1):
    """SinusoidalPosEmb.forward (models/embedders/time_embedder.py:15-28)"""
    half = dim // 2
    e = math.log(max_period) / (half - shift)
    e = torch.exp(-e * torch.arange(half))
    e = t[:, None] * e[None, :]
    return torch.cat((e.sin(), e.cos()), dim=-1)


def time_embedding(sd, t, pos_dim):
    """TimeEmbbeding.forward: sinus -> Linear -> Swish -> Linear (models/embedders/time_embedder.py:67-75)"""
    h = sinusoidal(t, pos_dim)
    h = F.linear(h, sd["time_embedder.time_emb.1.weight"], sd["time_embedder.time_emb.1.bias"])
    h = swish(h)
    return F.linear(h, sd["time_embedder.time_emb.3.weight"], sd["time_embedder.time_emb.3.bias"])


# ------------------------------------------------------------------------------------------------
# attention blocks (models/utils/attention_blocks.py)
# ------------------------------------------------------------------------------------------------
def compute_attention(q, k, v, heads, scale):
    """attention_blocks.py:35-43 — q,k,v [B, heads*d, N]; softmax((q*s)^T (k*s)) v"""
    b, c, n = q.shape
    d = c // heads
    q, k, v = (t.reshape(b * heads, d, -1) for t in (q, k, v))
    attn = torch.einsum("bdi,bdj->bij", q * scale, k * scale).softmax(dim=-1)
    out = torch.einsum("bij,bdj->bdi", attn, v)
    return out.reshape(b, c, -1)


def linear_transformer(sd, pre, x, groups, heads, embedding=None):
    """LinearTransformer.forward (attention_blocks.py:160-195); embedding [B, E] -> one key/value token"""
    b, c = x.shape[:2]
    spatial = x.shape[2:]
    x_n = F.group_norm(x, groups, sd[pre + ".norm_x.weight"], sd[pre + ".norm_x.bias"], eps=1e-5)
    emb = x_n if embedding is None else embedding.reshape(*embedding.shape[:2], *([1] * (x.ndim - 2)))
    x_n = x_n.reshape(b, c, -1)
    emb = emb.reshape(*emb.shape[:2], -1)
    q = F.conv1d(x_n, sd[pre + ".to_q.weight"], sd[pre + ".to_q.bias"])
    k = F.conv1d(emb, sd[pre + ".to_k.weight"], sd[pre + ".to_k.bias"])
    v = F.conv1d(emb, sd[pre + ".to_v.weight"], sd[pre + ".to_v.bias"])
    scale = (c // heads) ** -0.25
    out = compute_attention(q, k, v, heads, scale)
    out = F.conv1d(out, sd[pre + ".to_out.0.weight"], sd[pre + ".to_out.0.bias"])
    return x + out.reshape(b, c, *spatial)


def geglu(sd, pre, x):
    """GEGLU.forward (attention_blocks.py:17-25)"""
    b, c = x.shape[:2]
    spatial = x.shape[2:]
    h = x.reshape(b, c, -1).transpose(1, 2)
    h = F.layer_norm(h, (c,), sd[pre + ".norm.weight"], sd[pre + ".norm.bias"], eps=1e-5)
    h, gate = F.linear(h, sd[pre + ".proj.weight"], sd[pre + ".proj.bias"]).chunk(2, dim=-1)
    h = h * F.gelu(gate)
    return h.transpose(1, 2).reshape(b, -1, *spatial)


def spatial_transformer(sd, pre, x, groups, heads, embedding):
    """SpatialTransformer.forward (attention_blocks.py:276-288) with one BasicTransformerBlock (:223-231)"""
    h = F.group_norm(x, groups, sd[pre + ".norm.weight"], sd[pre + ".norm.bias"], eps=1e-5)
    h = F.conv2d(h, sd[pre + ".proj_in.weight"], sd[pre + ".proj_in.bias"])
    tb = pre + ".transformer_blocks.0"
    h = linear_transformer(sd, tb + ".self_atn", h, groups, heads, None)
    if embedding is not None and tb + ".cros_atn.to_q.weight" in sd:
        h = linear_transformer(sd, tb + ".cros_atn", h, groups, heads, embedding)
    o = geglu(sd, tb + ".proj_out.0", h)
    o = F.conv2d(o, sd[tb + ".proj_out.2.weight"], sd[tb + ".proj_out.2.bias"])
    h = o + h
    h = F.conv2d(h, sd[pre + ".proj_out.weight"], sd[pre + ".proj_out.bias"])
    return h + x


def attention(sd, pre, x, emb, groups, heads=8):
    """Attention.forward (attention_blocks.py:331-335): dispatch on which parameters exist"""
    pre = pre + ".attention"
    if pre + ".proj_in.weight" in sd:
        return spatial_transformer(sd, pre, x, groups, heads, emb)
    if pre + ".to_q.weight" in sd:
        return linear_transformer(sd, pre, x, groups, heads, emb)
    return x


# ------------------------------------------------------------------------------------------------
# UNet (models/estimators/unet2.py:222-269)
# ------------------------------------------------------------------------------------------------
def unet_forward(sd, cfg, x_t, t, cond=None, self_cond=None, with_ver=False):
    """cfg: dict(hid_chs, strides, num_res_blocks, groups, pos_emb_dim[, use_self_conditioning]).
    Returns y, or (y, y_ver) with the deep-supervision outputs (unet2.py:262,269) when with_ver."""
    hid, strides, nrb, G = cfg["hid_chs"], cfg["strides"], cfg.get("num_res_blocks", 2), cfg.get("groups", 32)
    depth = len(hid)
    if cfg.get("use_self_conditioning", False):                                        # unet2.py:243-246 (x_t, not self_cond)
        x_t = torch.cat([x_t, torch.zeros_like(x_t) if self_cond is None else x_t], dim=1)
    emb = time_embedding(sd, t, cfg["pos_emb_dim"])                                   # unet2.py:233
    if cond is not None and "cond_embedder.embedding.weight" in sd:
        emb = emb + F.embedding(cond, sd["cond_embedder.embedding.weight"])            # unet2.py:239-241
    xs = [conv2d(sd, "in_conv.conv", x_t, stride=strides[0])]                          # unet2.py:249
    idx = 0
    for i in range(1, depth):                                                          # unet2.py:250-251
        for _ in range(nrb):
            h = unet_res_block(sd, f"in_blocks.{idx}.0", xs[-1], emb, G)
            xs.append(attention(sd, f"in_blocks.{idx}.1", h, emb, G))                  # SequentialEmb, conv_blocks.py:21-25
            idx += 1
        if i < depth - 1:
            xs.append(conv2d(sd, f"in_blocks.{idx}.down_op", xs[-1], stride=strides[i]))  # conv_blocks.py:66-70
            idx += 1
    h = unet_res_block(sd, "middle_block.0", xs[-1], emb, G)                           # unet2.py:254
    h = attention(sd, "middle_block.1", h, emb, G)
    h = unet_res_block(sd, "middle_block.2", h, emb, G)
    n_out = (depth - 1) * (nrb + 1)
    y_ver = []
    for i in range(n_out, 0, -1):                                                      # unet2.py:258-264
        h = torch.cat([h, xs.pop()], dim=1)
        dpt, j = i // (nrb + 1), i % (nrb + 1) - 1                                     # unet2.py:261-262
        if dpt > 0 and j == 0 and f"outc_ver.{dpt - 1}.conv.conv.weight" in sd:
            y_ver.append(conv2d(sd, f"outc_ver.{dpt - 1}.conv.conv", h))
        pre = f"out_blocks.{i - 1}"
        h = unet_res_block(sd, pre + ".0", h, emb, G)
        h = attention(sd, pre + ".1", h, emb, G)
        if pre + ".2.up_op.weight" in sd:
            level = (i - 1) // (nrb + 1) + 1
            h = basic_up(sd, pre + ".2.up_op", h, strides[level])
    y = conv2d(sd, "outc.conv.conv", h)                                                # unet2.py:267
    return (y, y_ver[::-1]) if with_ver else y


# ------------------------------------------------------------------------------------------------
# VAE.encode (models/embedders/latent_embedders.py:756-762; DownBlock.forward conv_blocks.py:430-441;
# DiagonalGaussianDistribution :20-33)
# ------------------------------------------------------------------------------------------------
def vae_encode_moments(sd, cfg, x):
    G, depth = cfg.get("groups", 8), len(cfg["hid_chs"])
    h = unet_res_block(sd, "inc", x, None, G)
    for i in range(1, depth):
        h = conv2d(sd, f"encoders.{i - 1}.down_op.down_op", h, stride=cfg["strides"][i])
        h = unet_res_block(sd, f"encoders.{i - 1}.conv_block", h, None, G)
    return conv2d(sd, "out_enc.1.conv", conv2d(sd, "out_enc.0.conv", h))


def vae_encode(sd, cfg, x, noise):
    """noise: the torch.randn(mean.shape) draw of the quantizer (injected)."""
    mean, logvar = torch.chunk(vae_encode_moments(sd, cfg, x), 2, dim=1)
    return mean + torch.exp(0.5 * torch.clamp(logvar, -30.0, 20.0)) * noise


# ------------------------------------------------------------------------------------------------
# VAE.decode (models/embedders/latent_embedders.py:764-769; UpBlock.forward conv_blocks.py:510-528)
# ------------------------------------------------------------------------------------------------
def vae_decode(sd, cfg, z):
    G, depth = cfg.get("groups", 8), len(cfg["hid_chs"])
    h = unet_res_block(sd, "inc_dec", z, None, G)
    for i in range(depth - 1, 0, -1):
        pre = f"decoders.{i - 1}"
        h = basic_up(sd, pre + ".up_op.up_op", h, cfg["strides"][i])
        h = unet_res_block(sd, pre + ".conv_block", h, None, G)
    return conv2d(sd, "outc.conv", h)


# ------------------------------------------------------------------------------------------------
# VQVAE.decode (models/embedders/latent_embedders.py:314-320) and its VectorQuantizer (:40-72)
# ------------------------------------------------------------------------------------------------
def vq_quantize(codebook, z):
    """VectorQuantizer.forward, first output (latent_embedders.py:50-69): nearest codebook row by the expanded distance
    ||z||^2 + ||e||^2 - 2 z.e (argmin = first minimum), returned in the straight-through form z + (z_q - z).
    Also returns the indices and the [vectors, codes] distance matrix (for near-tie analysis in tests)."""
    C = codebook.shape[1]
    z_ch = torch.moveaxis(z, 1, -1)
    zf = z_ch.reshape(-1, C)
    dist = (torch.sum(zf ** 2, dim=1, keepdim=True) + torch.sum(codebook ** 2, dim=1)
            - 2 * torch.einsum("bd,dn->bn", zf, codebook.t()))
    idx = torch.argmin(dist, dim=1)
    z_q = torch.moveaxis(codebook[idx].view(z_ch.shape), -1, 1)
    return z + (z_q - z), idx.view(z_ch.shape[:-1]), dist


def vqvae_decode(sd, cfg, z):
    """VQVAE.decode: quantizer -> inc_dec -> decoders (last to first) -> outc  (latent_embedders.py:314-320)"""
    z_q, _, _ = vq_quantize(sd["quantizer.embedder.weight"], z)
    return vae_decode(sd, cfg, z_q)


# ------------------------------------------------------------------------------------------------
# scheduler (models/noise_schedulers/gaussian_scheduler.py)
# ------------------------------------------------------------------------------------------------
def scheduler_tables(timesteps=1000, schedule="scaled_linear", beta_start=0.002, beta_end=0.02):
    """gaussian_scheduler.py:22-58 (fp64 tables cast to fp32)"""
    f64 = torch.float64
    if schedule == "linear":
        betas = torch.linspace(beta_start, beta_end, timesteps, dtype=f64)
    elif schedule == "scaled_linear":
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, timesteps, dtype=f64) ** 2
    elif schedule == "cosine":
        s = 0.008
        x = torch.linspace(0, timesteps, timesteps + 1, dtype=f64)
        ac = torch.cos(((x / timesteps) + s) / (1 + s) * torch.pi * 0.5) ** 2
        ac = ac / ac[0]
        betas = torch.clip(1 - (ac[1:] / ac[:-1]), 0, 0.999)
    else:
        raise NotImplementedError(schedule)
    alphas = 1 - betas
    ac = torch.cumprod(alphas, dim=0)
    ac_prev = F.pad(ac[:-1], (1, 0), value=1.)
    tabs = dict(
        betas=betas, alphas=alphas, alphas_cumprod=ac, alphas_cumprod_prev=ac_prev,
        sqrt_alphas_cumprod=torch.sqrt(ac), sqrt_one_minus_alphas_cumprod=torch.sqrt(1. - ac),
        sqrt_recip_alphas_cumprod=torch.sqrt(1. / ac), sqrt_recipm1_alphas_cumprod=torch.sqrt(1. / ac - 1),
        posterior_mean_coef1=betas * torch.sqrt(ac_prev) / (1. - ac),
        posterior_mean_coef2=(1. - ac_prev) * torch.sqrt(alphas) / (1. - ac),
        posterior_variance=betas * (1. - ac_prev) / (1. - ac))
    return {k: v.to(torch.float32) for k, v in tabs.items()}


def _ext(tab, t, ndim):
    return tab.gather(0, t).reshape(-1, *((1,) * (ndim - 1)))  # scheduler_base.py:43-46


def estimate_x_t(tabs, x_0, t, x_T, T):
    """gaussian_scheduler.py:61-77 (per-sample clipper: t < 0 -> x_0, t >= T -> x_T)"""
    out = []
    for b in range(t.shape[0]):
        tb = int(t[b])
        if tb < 0:
            out.append(x_0[b])
        elif tb >= T:
            out.append(x_T[b])
        else:
            out.append(tabs["sqrt_alphas_cumprod"][tb] * x_0[b] + tabs["sqrt_one_minus_alphas_cumprod"][tb] * x_T[b])
    return torch.stack(out)


def sched_step(tabs, x_t, t, pred, noise, objective="x_T", clip_x0=False, var_scale=0, cold_diffusion=False):
    """estimate_x_t_prior_from_x_T/_x_0 (gaussian_scheduler.py:80-124). Returns (prior, x_0, x_T).
    var_scale: 0 or the per-element tensor pred_var/2+0.5 (diffusion_pipeline.py:254); cold_diffusion: :88-93."""
    nd = x_t.ndim
    A, Bm = _ext(tabs["sqrt_recip_alphas_cumprod"], t, nd), _ext(tabs["sqrt_recipm1_alphas_cumprod"], t, nd)
    if objective == "x_T":
        x_0 = A * x_t - Bm * pred
        x_0 = x_0.clamp(-1, 1) if clip_x0 else x_0
        x_T = pred
    else:
        x_0 = pred.clamp(-1, 1) if clip_x0 else pred
        x_T = (A * x_t - x_0) / Bm
    if cold_diffusion:
        T = tabs["betas"].shape[0]
        x_T_est = (A * x_t - x_0.clamp(-1, 1)) / Bm                 # estimate_x_T with its default clip_x0=True (:90)
        x_t_est = estimate_x_t(tabs, x_0, t, x_T_est, T)
        x_t_prior = estimate_x_t(tabs, x_0, t - 1, x_T_est, T)
        return x_t - (x_t_est - x_t_prior), x_0, x_T
    mean = _ext(tabs["posterior_mean_coef1"], t, nd) * x_0 + _ext(tabs["posterior_mean_coef2"], t, nd) * x_t
    lo = torch.log(_ext(tabs["posterior_variance"], t, nd).clamp(min=1e-20))
    hi = torch.log(_ext(tabs["betas"], t, nd).clamp(min=1e-20))
    logvar = var_scale * hi + (1 - var_scale) * lo                    # gaussian_scheduler.py:110-116
    std = torch.exp(0.5 * logvar)
    std[t == 0] = 0.0
    return mean + std * noise, x_0, x_T


def ddim_renoise(tabs, x_0, x_T, t, t_next, noise):
    """diffusion_pipeline.py:297-304 with eta == 1 (t, t_next: 0-dim long tensors)"""
    a, an = tabs["alphas_cumprod"][t], tabs["alphas_cumprod"][t_next]
    sigma = ((1 - a / an) * (1 - an) / (1 - a)).sqrt()
    c = (1 - an - sigma ** 2).sqrt()
    return x_0 * an.sqrt() + c * x_T + sigma * noise


def denoise(unet_fn, tabs, x_T, noises, steps, use_ddim=True, guidance_scale=1.0, cond=None, un_cond=None,
            objective="x_T", clip_x0=False, T=1000, estimate_variance=False, use_self_conditioning=False,
            cold_diffusion=False):
    """DiffusionPipeline.denoise without the latent decode (diffusion_pipeline.py:278-304; forward :232-275).
    `noises` is an iterator yielding the successive randn_like draws (scheduler draw, then DDIM draw).
    unet_fn(x_t, t, cond, self_cond) -> prediction ([B, 2C, ...] when estimate_variance)."""
    noises = iter(noises)
    if use_ddim:
        ts_arr = torch.linspace(0, T - 1, steps, dtype=torch.long)
    else:
        ts_arr = torch.linspace(0, T - 1, T, dtype=torch.long)[:steps]
        steps = len(ts_arr)
    x_t = x_T
    B = x_t.shape[0]
    self_cond = None
    for i, t in enumerate(ts_arr.flip(0)):
        tb = t.expand(B)
        var_scale = 0
        if cond is not None and guidance_scale != 1.0:                  # diffusion_pipeline.py:240-249
            pu = unet_fn(x_t, tb, un_cond, self_cond)
            pc = unet_fn(x_t, tb, cond, self_cond)
            pred = pu + guidance_scale * (pc - pu)
            if estimate_variance:   # (the reference forgets to chunk `pred` on this branch and fails; the intent is kept)
                pred, pv = pred.chunk(2, dim=1)
                var_scale = pv / 2 + 0.5
        else:
            pred = unet_fn(x_t, tb, cond, self_cond)
            if estimate_variance:                                        # :250-256
                pred, pv = pred.chunk(2, dim=1)
                var_scale = pv / 2 + 0.5
        noise = None if cold_diffusion else next(noises)
        x_t, x_0, x_Te = sched_step(tabs, x_t, tb, pred, noise, objective, clip_x0, var_scale, cold_diffusion)
        self_cond = (x_Te if objective == "x_0" else x_0) if use_self_conditioning else None   # :266,271,292
        if use_ddim and (steps - i - 1 > 0):
            x_t = ddim_renoise(tabs, x_0, x_Te, t, ts_arr[steps - i - 2], next(noises))
    return x_t
