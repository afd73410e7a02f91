"""GPU: kernel-level parity through the C ABI against the CPU oracle ops (rtol=1e-3, atol=1e-5)."""
import pytest
import torch
import torch.nn.functional as F

from util import ATOL, RTOL, assert_close, load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rnd(g, *shape, scale=1.0):
    return torch.randn(*shape, generator=g) * scale


# every distinct conv shape of the canonical UNet + the VAE decoder levels, small batch
TC_SHAPES = [
    # (N, H, W, C0, C1, Cout, k)
    (2, 32, 32, 256, 0, 256, 3), (2, 32, 32, 256, 256, 256, 3), (2, 32, 32, 256, 256, 256, 1),
    (2, 16, 16, 256, 0, 512, 3), (2, 16, 16, 256, 0, 256, 3), (2, 16, 16, 512, 0, 512, 3),
    (2, 16, 16, 512, 256, 256, 3), (2, 16, 16, 512, 512, 512, 3), (2, 16, 16, 256, 0, 512, 1),
    (2, 16, 16, 512, 256, 256, 1), (2, 16, 16, 512, 512, 512, 1),
    (3, 8, 8, 512, 0, 1024, 3), (3, 8, 8, 512, 0, 512, 3), (3, 8, 8, 1024, 0, 1024, 3),
    (3, 8, 8, 1024, 512, 512, 3), (3, 8, 8, 1024, 1024, 1024, 3), (3, 8, 8, 512, 0, 1024, 1),
    (3, 8, 8, 1024, 512, 512, 1), (3, 8, 8, 1024, 1024, 1024, 1),
    (1, 32, 32, 512, 0, 512, 3), (1, 64, 64, 512, 0, 256, 3), (1, 64, 64, 256, 0, 256, 3),
    (1, 128, 128, 256, 0, 128, 3), (1, 128, 128, 128, 0, 128, 3), (1, 256, 256, 128, 0, 64, 3),
    (1, 256, 256, 64, 0, 64, 3),
]


@pytest.mark.parametrize("shape", TC_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_conv_tc_matches_oracle(shape):
    from medfusion_b200 import ops
    N, H, W, C0, C1, Cout, k = shape
    g = torch.Generator().manual_seed(hash(shape) % 2 ** 31)
    C = C0 + C1
    x = _rnd(g, N, C, H, W)
    w = _rnd(g, Cout, C, k, k, scale=1.0 / (C * k * k) ** 0.5)
    b = _rnd(g, Cout, scale=0.1)
    ref = F.conv2d(x, w, b, padding=k // 2)
    xd, wd, bd = x.to(DEV), w.to(DEV), b.to(DEV)
    s0 = ops.pack_split(xd[:, :C0].contiguous())
    s1 = ops.pack_split(xd[:, C0:].contiguous()) if C1 else None
    assert ops.conv_tc_supported(N, H, W, C0, C1, Cout, k)
    out, stats = ops.conv_tc(s0, ops.prep_weight_tc(wd), bd, k, src1=s1, want_stats=True)
    assert_close(ops.unpack_nchw(out).cpu(), ref, what=f"conv_tc {shape}")
    # GroupNorm partial sums written by the epilogue: (sum, sumsq) over 8-channel slabs
    r8 = ref.double().view(N, Cout // 8, 8, -1)
    st = stats.double().sum(dim=1).cpu()
    assert_close(st[..., 0], r8.sum(dim=(2, 3)), 1e-3, 1e-2, "stats sum")
    assert_close(st[..., 1], (r8 * r8).sum(dim=(2, 3)), 1e-4, 1e-3, "stats sumsq")
    # split-plane (fp16 hi/lo) output variant carries the same values to 22 bits
    out2, _ = ops.conv_tc(s0, ops.prep_weight_tc(wd), bd, k, src1=s1, split_out=True)
    a, b2 = ops.unpack_nchw(out2), ops.unpack_nchw(out)
    assert bool(((a - b2).abs() <= 2.0 ** -21 * b2.abs() + 1.2e-7).all())   # 22 bits, or fp16-subnormal lo (6e-8 steps)


@pytest.mark.parametrize("shape", [(2, 32, 32, 256, 256), (3, 16, 16, 512, 512), (1, 64, 64, 64, 128)],
                         ids=lambda s: "x".join(map(str, s)))
def test_conv_tc_stride2_matches_oracle(shape):
    """BasicDown (conv_blocks.py:43-52,66-70): 3x3, stride 2, pad 1 — four input-parity TMA views, same kernel."""
    from medfusion_b200 import ops
    N, H, W, Cin, Cout = shape
    g = torch.Generator().manual_seed(sum(shape))
    x = _rnd(g, N, Cin, H, W)
    w = _rnd(g, Cout, Cin, 3, 3, scale=1.0 / (Cin * 9) ** 0.5)
    b = _rnd(g, Cout, scale=0.1)
    ref = F.conv2d(x, w, b, stride=2, padding=1)
    out, stats = ops.conv_tc(ops.pack_split(x.to(DEV)), ops.prep_weight_tc(w.to(DEV)), b.to(DEV), 3, stride=2,
                             want_stats=True)
    assert_close(ops.unpack_nchw(out).cpu(), ref, what=f"conv_tc stride 2 {shape}")
    r8 = ref.double().view(N, Cout // 8, 8, -1)
    assert_close(stats.double().sum(dim=1)[..., 0].cpu(), r8.sum(dim=(2, 3)), 1e-3, 1e-2, "stats sum")


@pytest.mark.parametrize("shape", [(2, 16, 16, 256, 256), (3, 8, 8, 512, 512), (1, 64, 64, 128, 64), (1, 32, 32, 512, 256)],
                         ids=lambda s: "x".join(map(str, s)))
def test_upconv_fold_matches_oracle(shape):
    """BasicUp (conv_blocks.py:121-131): conv3x3(nearest-exact x2) as four 2x2 phase convolutions with pre-summed taps."""
    from medfusion_b200 import ops
    N, H, W, Cin, Cout = shape
    g = torch.Generator().manual_seed(sum(shape) + 1)
    x = _rnd(g, N, Cin, H, W)
    w = _rnd(g, Cout, Cin, 3, 3, scale=1.0 / (Cin * 9) ** 0.5)
    b = _rnd(g, Cout, scale=0.1)
    ref = F.conv2d(F.interpolate(x, size=(2 * H, 2 * W), mode="nearest-exact"), w, b, padding=1)
    out = ops.upconv_tc(ops.pack_split(x.to(DEV)), w.to(DEV), b.to(DEV))
    assert_close(ops.unpack_nchw(out).cpu(), ref, what=f"folded upconv {shape}")


def test_conv_tc_ragged_batch_tail():
    """8x8 maps pack two samples per 128-row tile; an odd batch exercises the zero-filled / masked tail."""
    from medfusion_b200 import ops
    g = torch.Generator().manual_seed(5)
    x, w, b = _rnd(g, 5, 64, 8, 8), _rnd(g, 64, 64, 3, 3, scale=0.05), _rnd(g, 64)
    ref = F.conv2d(x, w, b, padding=1)
    out, stats = ops.conv_tc(ops.pack_split(x.to(DEV)), ops.prep_weight_tc(w.to(DEV)), b.to(DEV), 3, want_stats=True)
    assert_close(ops.unpack_nchw(out).cpu(), ref, what="ragged tail")
    assert_close(stats.sum(dim=1)[..., 0].cpu().double(), ref.double().view(5, 8, 8, -1).sum(dim=(2, 3)), 1e-3, 1e-2)


SIMT_SHAPES = [
    # (N, Cin, H, W, Cout, k, stride, in_layout, out_layout)
    (2, 8, 32, 32, 256, 3, 1, 0, 2), (2, 256, 32, 32, 256, 3, 2, 2, 2), (2, 512, 16, 16, 512, 3, 2, 2, 2),
    (2, 256, 32, 32, 8, 1, 1, 2, 0), (1, 8, 32, 32, 512, 3, 1, 0, 1), (1, 8, 32, 32, 512, 1, 1, 0, 1),
    (1, 64, 64, 64, 3, 1, 1, 2, 0), (3, 24, 7, 5, 40, 3, 1, 0, 1),
]


@pytest.mark.parametrize("shape", SIMT_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_conv_simt_matches_oracle(shape):
    from medfusion_b200 import ops
    N, Cin, H, W, Cout, k, stride, inl, outl = shape
    g = torch.Generator().manual_seed(7)
    x = _rnd(g, N, Cin, H, W)
    w = _rnd(g, Cout, Cin, k, k, scale=1.0 / (Cin * k * k) ** 0.5)
    b = _rnd(g, Cout, scale=0.1)
    ref = F.conv2d(x, w, b, stride=stride, padding=(k - stride + 1) // 2 if k > 1 else 0)
    xd = x.to(DEV)
    xin = xd if inl == 0 else ops.pack_split(xd)
    out = ops.conv_simt(xin, inl, ops.prep_weight_simt(w.to(DEV)), b.to(DEV), Cin, k, stride, outl)
    got = out if outl == 0 else ops.unpack_nchw(out)
    assert_close(got.cpu(), ref, what=f"conv_simt {shape}")


@pytest.mark.parametrize("C,G,H,W", [(256, 32, 32, 32), (1024, 32, 8, 8), (512, 8, 32, 32), (64, 8, 64, 64)])
def test_groupnorm_swish_residual_emb(C, G, H, W):
    from medfusion_b200 import ops
    g = torch.Generator().manual_seed(3)
    N = 2
    x = _rnd(g, N, C, H, W) * 2 + 0.3
    gamma, beta = 1 + 0.1 * _rnd(g, C), 0.1 * _rnd(g, C)
    r, emb = _rnd(g, N, C, H, W), _rnd(g, N, C)
    y = F.group_norm(x, G, gamma, beta, 1e-5)
    ref = y * torch.sigmoid(y) + r + emb[:, :, None, None]
    raw = x.to(DEV).permute(0, 2, 3, 1).contiguous()
    mr = ops.gn_finalize(ops.gn_partial(raw), C, G, H * W)
    out = ops.gn_apply(raw, mr, gamma.to(DEV), beta.to(DEV), G, res=ops.pack_split(r.to(DEV)), emb=emb.to(DEV))
    assert_close(ops.unpack_nchw(out).cpu(), ref, what="gn+swish+res+emb")
    out = ops.gn_apply(raw, mr, gamma.to(DEV), beta.to(DEV), G, res=None, emb=None)
    assert_close(ops.unpack_nchw(out).cpu(), y * torch.sigmoid(y), what="gn+swish")


def test_upsample_nearest_exact():
    from medfusion_b200 import ops
    x = _rnd(torch.Generator().manual_seed(4), 2, 64, 8, 16)
    p = ops.pack_split(x.to(DEV))
    ref = F.interpolate(ops.unpack_nchw(p).cpu(), size=(16, 32), mode="nearest-exact")   # exact on the carried values
    assert torch.equal(ops.unpack_nchw(ops.upsample2x(p)).cpu(), ref)
    assert_close(ref, F.interpolate(x, size=(16, 32), mode="nearest-exact"), 1e-6, 1e-7, "split round trip")


def test_split_planes_carry_22_bits():
    """x = hi + lo with fp16 planes: relative error <= 2^-21 (11 + 11 significant bits), saturation instead of inf."""
    from medfusion_b200 import ops
    x = _rnd(torch.Generator().manual_seed(9), 1, 32, 4, 8) * 1e3
    p = ops.pack_split(x.to(DEV))
    assert p.dtype == torch.float16
    back = ops.unpack_nchw(p).cpu()
    assert float(((back - x).abs() / x.abs().clamp_min(1e-3)).max()) <= 2.0 ** -21
    big = torch.full((1, 8, 2, 2), 1e6)
    sat = ops.unpack_nchw(ops.pack_split(big.to(DEV)))
    assert bool(torch.isfinite(sat).all()) and float(sat.max()) <= 2 * 65504.0


def test_scheduler_step_matches_reference_fixture():
    from medfusion_b200.models import GaussianNoiseScheduler
    g = load_golden("sched.pt")
    s = GaussianNoiseScheduler(**g["sched"]).to(DEV)
    x_t, pred, noise, t = (g[k].to(DEV) for k in ("x_t", "pred", "noise", "t"))
    for clip in (False, True):
        ref = g["out"][f"xT_clip{int(clip)}"]
        o = s.step(x_t, t, pred, noise=noise, objective="x_T", clip_x0=clip, want=("x_prior", "x_0"))
        assert_close(o["x_prior"].cpu(), ref["prior"], what="prior (x_T objective)")
        assert_close(o["x_0"].cpu(), ref["x_0"], what="x_0 (x_T objective)")
        ref = g["out"][f"x0_clip{int(clip)}"]
        o = s.step(x_t, t, pred, noise=noise, objective="x_0", clip_x0=clip, want=("x_prior", "x_0", "x_T"))
        assert_close(o["x_prior"].cpu(), ref["prior"], what="prior (x_0 objective)")
        assert_close(o["x_T"].cpu(), ref["x_T"], what="x_T (x_0 objective)")
    # t == 0 rows carry no noise (std[t==0] = 0, gaussian_scheduler.py:98)
    o = s.step(x_t, t, pred, noise=noise * 1e6, objective="x_T", clip_x0=False, want=("x_prior",))
    o2 = s.step(x_t, t, pred, noise=None, objective="x_T", clip_x0=False, want=("x_prior",))
    assert torch.equal(o["x_prior"][4], o2["x_prior"][4])


def test_scheduler_cfg_and_ddim_against_oracle():
    import medfusion_oracle as O
    from medfusion_b200.models import GaussianNoiseScheduler
    g = load_golden("sched.pt")
    sc = g["sched"]
    tabs = O.scheduler_tables(sc["timesteps"], sc["schedule_strategy"], sc["beta_start"], sc["beta_end"])
    s = GaussianNoiseScheduler(**sc).to(DEV)
    gen = torch.Generator().manual_seed(8)
    x_t, pu, pc, n1, n2 = (torch.randn(4, 8, 16, 16, generator=gen) for _ in range(5))
    for tval, tnext in ((999, 749), (500, 250), (250, 0)):
        t = torch.full((4,), tval)
        pred = pu + 3.0 * (pc - pu)
        prior, x0, xT = O.sched_step(tabs, x_t, t, pred, n1, "x_T", False)
        nxt = O.ddim_renoise(tabs, x0, xT, torch.tensor(tval), torch.tensor(tnext), n2)
        o = s.step(x_t.to(DEV), t.to(DEV), pc.to(DEV), pred_uncond=pu.to(DEV), guidance_scale=3.0, noise=n1.to(DEV),
                   t_next=torch.tensor(tnext), noise_ddim=n2.to(DEV), objective="x_T", clip_x0=False,
                   want=("x_prior", "x_0", "x_next"))
        assert_close(o["x_prior"].cpu(), prior, what="cfg prior")
        assert_close(o["x_0"].cpu(), x0, what="cfg x_0")
        assert_close(o["x_next"].cpu(), nxt, what="ddim re-noise")


# the last three shapes have enough (batch, head, query-block) work to take the many-queries-per-warp variants
# (N in {64,128,192,256}, d in {64,128}) take the tcgen05 kernel, everything else the CUDA-core one
@pytest.mark.parametrize("B,N,heads,d", [(2, 64, 8, 32), (1, 256, 8, 128), (3, 100, 4, 64), (32, 256, 8, 128),
                                         (16, 100, 8, 64), (24, 70, 8, 32), (3, 64, 8, 64), (2, 64, 8, 128),
                                         (2, 128, 4, 64), (1, 192, 2, 128), (5, 256, 8, 64)])
def test_attention_core_matches_oracle(B, N, heads, d):
    import ctypes
    import medfusion_oracle as O
    from medfusion_b200 import _lib, ops
    g = torch.Generator().manual_seed(N + d)
    C = heads * d
    qkv = _rnd(g, B, N, 3 * C)
    q, k, v = (qkv[..., i * C:(i + 1) * C].transpose(1, 2).contiguous() for i in range(3))   # [B, C, N]
    ref = O.compute_attention(q, k, v, heads, d ** -0.25).transpose(1, 2)                      # [B, N, C]
    dq = qkv.to(DEV).contiguous()
    out = torch.empty((2, B, N, 1, C), device=DEV, dtype=torch.float16)
    _lib.check(_lib.load().mf_op_attention(dq.data_ptr(), dq.data_ptr() + 4 * C, dq.data_ptr() + 8 * C, 3 * C,
                                           out.data_ptr(), out[0].numel(), B, N, heads, d,
                                           torch.cuda.current_stream().cuda_stream), "attention")
    got = (out[0].float() + out[1].float()).reshape(B, N, C).cpu()
    assert_close(got, ref, what="attention core")


def test_layernorm_and_geglu_match_oracle():
    from medfusion_b200 import _lib, ops
    g = torch.Generator().manual_seed(12)
    T, C = 192, 256
    x = _rnd(g, T, C) * 3 + 1
    gamma, beta = 1 + 0.1 * _rnd(g, C), 0.1 * _rnd(g, C)
    ref = F.layer_norm(x, (C,), gamma, beta, 1e-5)
    xs = ops.pack_split(x.t().reshape(1, C, T, 1).contiguous().to(DEV))                      # [2,1,T,1,C]
    out = torch.empty_like(xs)
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    dg, db = gamma.to(DEV), beta.to(DEV)   # keep the device copies alive across the asynchronous launch
    _lib.check(lib.mf_op_layernorm(xs.data_ptr(), xs[0].numel(), dg.data_ptr(), db.data_ptr(),
                                   out.data_ptr(), out[0].numel(), T, C, 1e-5, st), "layernorm")
    assert_close((out[0].float() + out[1].float()).reshape(T, C).cpu(), ref, what="layernorm")
    z = _rnd(g, T, 2 * C)
    refg = z[:, :C] * F.gelu(z[:, C:])
    zo = torch.empty((2, T, C), device=DEV, dtype=torch.float16)
    dz = z.to(DEV)
    _lib.check(lib.mf_op_geglu(dz.data_ptr(), zo.data_ptr(), zo[0].numel(), T, C, st), "geglu")
    assert_close((zo[0].float() + zo[1].float()).cpu(), refg, what="geglu")


def _heavy_tailed_case(shape, dof, out_frac, out_scale):
    N, H, W, Cin, Cout, k = shape
    g = torch.Generator().manual_seed(Cin + 7 * k + dof)
    x = torch.randn(N, Cin, H, W, generator=g)
    x = x * torch.sigmoid(x) + 0.3
    out_mask = torch.rand(N, Cin, H, W, generator=g) < out_frac
    x = torch.where(out_mask, out_scale * torch.randn(N, Cin, H, W, generator=g), x)
    chi = torch.randn(dof, Cout, Cin, k, k, generator=g).pow(2).sum(0) / dof
    w = torch.randn(Cout, Cin, k, k, generator=g) / chi.sqrt() / (Cin * k * k) ** 0.5     # t_dof / sqrt(fan_in)
    b = 0.1 * torch.randn(Cout, generator=g)
    return x, w, b


@pytest.mark.parametrize("drain", [1, 3])
@pytest.mark.parametrize("profile", ["trained-like", "extreme"])
@pytest.mark.parametrize("shape", [(2, 32, 32, 256, 256, 3), (2, 8, 8, 1024, 1024, 3), (1, 16, 16, 512, 512, 1)])
def test_conv_tc_heavy_tailed_operands_keep_fp32_parity(shape, profile, drain):
    """the accumulator de-bias was calibrated on Gaussian / Swish-like operands.  Trained networks have
    heavy-tailed weights and activation outliers.
    * "trained-like": Student-t (5 degrees of freedom: kurtosis 9) weights, Swish-like activations with a non-zero mean and
      0.1 % outliers of 10x the scale — must hold the plain tolerance against the fp64 convolution.
    * "extreme": t_3 weights (infinite kurtosis), 1 % outliers of 30x the scale.  Here sum |x||w| is far larger than the
      result (cancellation), and ANY fp32 evaluation is only accurate relative to sum |x||w|: the reference's own fp32
      convolution (torch CPU) is measured against fp64 on the same data, and this kernel must stay within the plain
      tolerance + 2^-20 * sum |x||w| — the scale of the rounding noise between two fp32 implementations that sum in a
      different order (sqrt(K) * 2^-24 * sum |x||w| ~ 3e-6 * sum |x||w| at K = 2304), and it is reported next to the
      reference's own error."""
    from medfusion_b200 import ops
    N, H, W, Cin, Cout, k = shape
    x, w, b = _heavy_tailed_case(shape, 5, 0.001, 10.0) if profile == "trained-like" else _heavy_tailed_case(shape, 3, 0.01, 30.0)
    ref64 = F.conv2d(x.double(), w.double(), b.double(), padding=k // 2)
    ref = ref64.float()
    ref32 = F.conv2d(x, w, b, padding=k // 2)                      # the reference's own arithmetic (torch CPU fp32)
    mag = F.conv2d(x.abs().double(), w.abs().double(), None, padding=k // 2)   # sum |x||w| per output
    xs = ops.pack_split(x.to(DEV))
    wp = ops.prep_weight_tc(w.to(DEV))
    out, _ = ops.conv_tc(xs, wp, b.to(DEV), k, drain_interval=drain)
    got = ops.unpack_nchw(out).cpu()
    tol = ATOL + RTOL * ref64.abs()
    d = (got.double() - ref64).abs()
    d32 = (ref32.double() - ref64).abs()
    worst, worst32 = float((d / tol).max()), float((d32 / tol).max())
    fwd = float((d / (tol + 2.0 ** -20 * mag)).max())
    print(f"heavy-tailed conv {shape} {profile} drain {drain}: worst |err|/tol = {worst:.3f} (torch CPU fp32 on the same data: "
          f"{worst32:.3f}); against tol + 2^-20 sum|x||w|: {fwd:.3f}; max sum|x||w| = {float(mag.max()):.1f}, |ref|max = "
          f"{float(ref64.abs().max()):.1f}")
    if profile == "trained-like":
        assert worst <= 1.0, f"worst |err|/tol = {worst:.3f}"
    else:
        assert fwd <= 1.0, f"worst |err| / (tol + 2^-20 sum|x||w|) = {fwd:.3f}"


@pytest.mark.parametrize("shape", [(3, 128, 128, 128, 0, 128, 3), (1, 128, 128, 128, 128, 128, 3), (2, 256, 256, 64, 0, 64, 3),
                                   (1, 256, 256, 128, 0, 64, 3)], ids=lambda s: "x".join(map(str, s)))
def test_conv_tc_row_patch_equals_per_tap_tiles(shape):
    """Row-patch mode (mf_set_row_patch, the default for 3x3 layers tiled by image rows: the VAE's 128x128 / 256x256 levels):
    the 130-pixel patch of an input row is staged once for its three taps and addressed by shifted descriptors.  Same
    products as the per-tap tiles (equal up to the fp32 summation order of the stream-K split); and equal to the oracle.  Covers an
    odd batch, the image borders (zero-filled halo) and a two-source (concat) K loop."""
    from medfusion_b200 import _lib, ops
    lib = _lib.load()
    N, H, W, C0, C1, Cout, k = shape
    g = torch.Generator().manual_seed(sum(shape))
    C = C0 + C1
    x = _rnd(g, N, C, H, W)
    w = _rnd(g, Cout, C, k, k, scale=1.0 / (C * k * k) ** 0.5)
    b = _rnd(g, Cout, scale=0.1)
    ref = F.conv2d(x, w, b, padding=k // 2)
    xd, wd, bd = x.to(DEV), w.to(DEV), b.to(DEV)
    s0 = ops.pack_split(xd[:, :C0].contiguous())
    s1 = ops.pack_split(xd[:, C0:].contiguous()) if C1 else None
    wp = ops.prep_weight_tc(wd)
    outs = {}
    try:
        for mode in (0, 1):
            lib.mf_set_row_patch(mode)
            out, _ = ops.conv_tc(s0, wp, bd, k, src1=s1)
            outs[mode] = ops.unpack_nchw(out).cpu()
    finally:
        lib.mf_set_row_patch(1)
    assert_close(outs[1], ref, what=f"row-patch conv {shape}")
    # same products; the stream-K unit boundaries (and with them the fp32 summation order) differ between the two plans
    assert float((outs[0] - outs[1]).abs().max()) <= 4e-6 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("shape", [(1, 8, 8, 1024, 1024, 1024, 3), (4, 8, 8, 1024, 0, 1024, 3), (4, 32, 32, 256, 0, 256, 3),
                                   (2, 16, 16, 512, 512, 512, 1), (1, 16, 16, 256, 0, 512, 3)], ids=lambda s: "x".join(map(str, s)))
def test_conv_tc_small_batch_fill_matches_unshared_tiles(shape):
    """Small-batch fill (mf_set_split_fill): a layer with fewer tiles than half the SM pairs gets narrower tiles and / or
    several CTA pairs per tile (partial tiles through the stream-K scratch, added by the owner in group order).  Against the
    oracle, and against the one-pair-per-tile plan within fp32 re-association noise; run twice: the result is deterministic
    and the flags are re-armed."""
    from medfusion_b200 import _lib, ops
    lib = _lib.load()
    N, H, W, C0, C1, Cout, k = shape
    g = torch.Generator().manual_seed(sum(shape) + 1)
    C = C0 + C1
    x = _rnd(g, N, C, H, W)
    w = _rnd(g, Cout, C, k, k, scale=1.0 / (C * k * k) ** 0.5)
    b = _rnd(g, Cout, scale=0.1)
    ref = F.conv2d(x, w, b, padding=k // 2)
    xd, wd, bd = x.to(DEV), w.to(DEV), b.to(DEV)
    s0 = ops.pack_split(xd[:, :C0].contiguous())
    s1 = ops.pack_split(xd[:, C0:].contiguous()) if C1 else None
    wp = ops.prep_weight_tc(wd)
    outs = {}
    try:
        for mode in (0, 8, 4):
            lib.mf_set_split_fill(mode)
            a = ops.unpack_nchw(ops.conv_tc(s0, wp, bd, k, src1=s1, want_stats=True)[0]).cpu()
            b2 = ops.unpack_nchw(ops.conv_tc(s0, wp, bd, k, src1=s1, want_stats=True)[0]).cpu()
            assert torch.equal(a, b2), f"split_fill={mode}: not deterministic"
            outs[mode] = a
    finally:
        lib.mf_set_split_fill(8)
    for mode, o in outs.items():
        assert_close(o, ref, what=f"conv {shape} split_fill={mode}")
    assert float((outs[8] - outs[0]).abs().max()) <= 4e-6 * max(1.0, float(ref.abs().max()))
