/* medfusion_b200 — C ABI of the B200-native Medfusion sampling hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  Every
 * pointer argument named `d_*` is a DEVICE pointer owned by the caller (PyTorch's caching allocator
 * in the shipped Python host layer); the library borrows it for the duration of the stream-ordered
 * call.  Weights are copied / re-laid-out once into library-owned device memory by *_set_param.
 * All functions return 0 on success; on failure a thread-local message is available from
 * mf_last_error().  Nothing in the forward/decode/step calls allocates or synchronises, so they
 * can be captured into a CUDA graph after one warm-up call for the same (B, H, W, workspace).
 *
 * Reference interfaces replaced (paths relative to /root/reference):
 *   mf_unet_forward   <- medical_diffusion/models/estimators/unet2.py:222-269        UNet.forward
 *   mf_vae_decode     <- medical_diffusion/models/embedders/latent_embedders.py:764-769  VAE.decode
 *   mf_sched_step     <- medical_diffusion/models/noise_schedulers/gaussian_scheduler.py:80-124
 *                        (estimate_x_t_prior_from_x_T / _x_0, estimate_x_0, estimate_mean_t,
 *                        estimate_variance_t) + diffusion_pipeline.py:240-244 (CFG combine)
 *                        + diffusion_pipeline.py:297-304 (DDIM-form re-noise)
 *   mf_unet_forward_step <- the per-timestep body of DiffusionPipeline.denoise, pipelines/diffusion_pipeline.py:290-304
 *                        (estimator + forward()'s scheduler dispatch :264-273 + DDIM re-noise) in one call
 *   mf_sched_step_opts <- the learned-variance / cold-diffusion branches, diffusion_pipeline.py:246-262 and
 *                        gaussian_scheduler.py:61-77,88-116
 *   mf_vae_decode_u8  <- VAE.decode + the uint8 conversion of scripts/helpers/sample_dataset.py:47-50
 *   mf_vae_encode     <- medical_diffusion/models/embedders/latent_embedders.py:756-762 VAE.encode (+ :20-33 quantizer)
 *   mf_op_*           <- the individual torch ops those functions are made of (test surface)
 *
 * ABI version 2 (mf_abi_version): v1 + mf_vae_config.in_channels, mf_sched_step_opts, mf_vae_encode*, mf_set_pdl.
 * ABI version 3: v2 + mf_saturation_count, mf_vqvae_* (VQVAE.decode), mf_unet_forward_step2 (CFG as one 2B batch).
 * Tuning knobs (mf_set_*) are not part of the versioned data-path ABI: later builds add knobs (mf_set_split_fill, mf_set_row_patch)
 * without a version bump; a binding should resolve them optionally.
 */
#ifndef MEDFUSION_B200_H_
#define MEDFUSION_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* mf_stream_t; /* cudaStream_t */

#define MF_MAX_LEVELS 8

/* -------------------------------------------------------------------------------------------------
 * Errors / build info
 * ---------------------------------------------------------------------------------------------- */
const char* mf_last_error(void);
int mf_abi_version(void);
/* Sticky saturation counter (ABI v3).  Activations travel between kernels as fp16 hi/lo planes; the reference computes
 * in fp32 range.  Every value that had to be clamped to +-65504 (or was not finite) when a kernel wrote split planes is
 * counted on the device (the streaming GroupNorm-apply kernel counts once per thread that clamped: a lower bound).
 * Synchronises `stream`, returns the count since the last reset on the CURRENT device; a
 * non-zero count means results from that point on are not the reference's (the Python layer raises). */
int mf_saturation_count(unsigned long long* out_count, int reset, mf_stream_t stream);
/* Default TMEM drain interval used by the engines' tensor-core convolutions (see mf_op_conv_tc). */
int mf_set_drain_interval(int k_blocks);
/* tcgen05 convolutions: 1 = one CTA per 128-pixel tile, 2 = CTA pairs (cta_group::2, 256-row MMA), 0 = auto.
 * Affects plans built afterwards. */
int mf_set_cta_group(int cta_group);
/* Output channels per tcgen05 tile (0 = auto, 64, 128, 256; reduced automatically until it divides Cout). */
int mf_set_block_n(int block_n);
/* 1 (default): persistent stream-K schedule — one CTA group per SM, (tile, K block) units split evenly, split tiles
 * reduced through a scratch buffer; 0: one output tile per CTA group. */
int mf_set_stream_k(int enable);
/* Small-batch fill (default 8): a convolution with fewer output tiles than half the SM pairs (batch 1..16 of
 * scripts/sample.py) lets several CTA pairs share the K range of one tile, each keeping at least `min_k_blocks` 64-channel
 * K blocks; their partial sums reach the tile's owner through the stream-K scratch.  0 = off (one pair per tile at most).
 * Takes effect at the next plan build. */
int mf_set_split_fill(int min_k_blocks);
/* Row-patch mode (default 1): 3x3 stride-1 convolutions whose tile is 128 pixels of one image row and at most 128 output
 * channels wide (the 128x128 / 256x256 levels of the VAE) stage the 130-pixel patch of an input row once for its three taps
 * (shifted shared-memory descriptors) instead of once per tap: a third of the L2 -> SM activation traffic.  0 = one
 * activation tile per tap everywhere.  Takes effect at the next plan build. */
int mf_set_row_patch(int enable);
/* Schedule the library would choose for a tensor-core convolution of this shape (INPUT geometry N x H x W, C0 (+ C1
 * concatenated) -> Cout, ksize 1 | 3, stride 1 | 2, up2 = folded nearest-x2) on a device with `sm_count` SMs (<= 0: 148), under
 * the current knobs.  Pure host arithmetic — no device is touched, so the planning heuristics are testable without a GPU.
 * out8 = {supported, block_n, cta_group, CTA groups of the persistent grid, row-patch mode, output tiles, K blocks per tile,
 * tile groups along M}. */
int mf_op_conv_tc_plan(int N, int H, int W, int C0, int C1, int Cout, int ksize, int stride, int up2, int sm_count, int* out8);
/* Relative correction applied to every drained TMEM partial sum, per K block of the drain interval, compensating the
 * round-toward-zero bias of the tcgen05 accumulator (0 disables, negative = built-in calibrated table, the default). */
int mf_set_debias_eps(float eps_per_kblock);
/* BasicUp (nearest x2 + conv3x3, conv_blocks.py:121-131): 1 = four 2x2 phase convolutions on the low-resolution
 * input with pre-summed weights (default), 0 = explicit upsample kernel followed by the 3x3 convolution. */
int mf_set_fold_upsample(int enable);
/* Cin < 64 stem convolutions (UNet in_conv, VAE inc_dec): 1 = tcgen05 path through a zero-padded 64-channel copy of the
 * NCHW input (default), 0 = exact-fp32 CUDA-core kernel. */
int mf_set_stem_on_tc(int enable);
/* Attention core (compute_attention, attention_blocks.py:35-43): 1 (default) = tcgen05 kernel (S = QK^T and PV as fp16x3
 * tensor-core products, accumulators in TMEM, fp32 softmax in registers) for N in {64,128,192,256} tokens and head
 * dim in {64,128}; 0 = the CUDA-core online-softmax kernel for every shape. */
int mf_set_attn_tc(int enable);
/* Res-block halves (conv -> GroupNorm -> Swish -> + residual -> + embedding, conv_blocks.py:184-192,236-240,362) whose
 * output tile spans whole samples (H*W <= 128 per CTA, or H*W == 256 per CTA pair with the statistics exchanged through
 * distributed shared memory): 1 = normalisation applied in the convolution's epilogue (no raw fp32 tensor, no GroupNorm
 * launch: 65 instead of 89 launches per canonical UNet step), 0 (default) = separate GroupNorm-apply kernel everywhere.
 * Measured slower end to end on B200 (the fused work runs exposed at the tail of a one-tile-per-CTA kernel;
 * profiles/r02_gn_fusion.md), hence opt-in.  Takes effect at the next plan build. */
int mf_set_fuse_gn(int enable);
/* VAE / VQVAE image head (latent_embedders.py:743, 1x1 conv hid_chs[0] -> out_channels): 1 (default) = evaluated inside the
 * last GroupNorm-apply kernel (the 1.07 GB activation at B=64, 256x256 is neither written nor re-read), 0 = own kernel. */
int mf_set_fold_head(int enable);
/* GroupNorm-apply kernel used by the engine plans (tuning knob): 3 (default) finalises the statistics inside the apply
 * kernel (one launch per GroupNorm, 8 channels per thread); 0 / 1 / 2 keep a separate finalize launch with a flat
 * grid-stride / fixed channel quad per thread / one channel quad per thread mapping.  Takes effect at the next plan build. */
int mf_set_gn_variant(int v);
/* 1 (default): conv_tc and the fused GroupNorm-apply kernel are launched with programmatic stream serialization
 * (griddepcontrol): a kernel's set-up overlaps the tail of its predecessor, also inside captured CUDA graphs. */
int mf_set_pdl(int enable);

/* scheduler tables: fp32[T] device arrays as registered by gaussian_scheduler.py:44-58 */
typedef struct {
  const float* sqrt_recip_alphas_cumprod;
  const float* sqrt_recipm1_alphas_cumprod;
  const float* posterior_mean_coef1;
  const float* posterior_mean_coef2;
  const float* posterior_variance;
  const float* betas;
  const float* alphas_cumprod;
} mf_sched_tables;

/* -------------------------------------------------------------------------------------------------
 * UNet noise estimator (unet2.py:15-219 constructor arguments, restricted to the 2-D res-block
 * configuration the hot path uses)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int in_ch;                       /* unet2.py:18  in_ch (x2 if self-conditioning; not supported) */
  int out_ch;                      /* unet2.py:19  out_ch */
  int depth;                       /* len(hid_chs) */
  int hid_chs[MF_MAX_LEVELS];      /* unet2.py:21 */
  int kernel_sizes[MF_MAX_LEVELS]; /* unet2.py:22 */
  int strides[MF_MAX_LEVELS];      /* unet2.py:23 (last stride ignored, like the reference) */
  int num_res_blocks;              /* unet2.py:37 */
  int emb_dim;                     /* TimeEmbbeding emb_dim (time_embedder.py:54); 0 = no time embedder */
  int pos_emb_dim;                 /* sinusoidal width (emb_dim // 4 by default, time_embedder.py:62) */
  int num_classes;                 /* LabelEmbedder num_classes (cond_embedders.py:6); 0 = none */
  int norm_groups;                 /* GroupNorm groups (32) */
  int attention[MF_MAX_LEVELS];    /* per level: 0 = 'none', 1 = 'linear', 2 = 'spatial' (attention_blocks.py:291-335) */
  int deep_supervision;            /* number of deep-supervision heads `outc_ver` (unet2.py:214-217; True = depth-2) (ABI v3) */
  int ds_out_ch;                   /* their output channels (0 = out_ch; the reference uses the plain out_ch even when
                                      estimate_variance doubles the main head) (ABI v3) */
} mf_unet_config;

typedef struct mf_unet mf_unet;

int mf_unet_create(const mf_unet_config* cfg, mf_unet** out);
void mf_unet_destroy(mf_unet* h);
/* Enumerate the state_dict entries the engine expects (same names/shapes as the reference module). */
int mf_unet_param_count(const mf_unet* h);
const char* mf_unet_param_name(const mf_unet* h, int i);
int mf_unet_param_shape(const mf_unet* h, int i, int64_t shape[4], int* ndim);
/* Copy one fp32 parameter from a device buffer (contiguous, reference layout: conv OIHW, linear [out,in]). */
int mf_unet_set_param(mf_unet* h, const char* name, const float* d_data, const int64_t* shape, int ndim,
                      mf_stream_t stream);
/* Sinusoidal frequency table exp(-ln(max_period) * k / (half - shift)), k < pos_emb_dim/2, computed by
 * the host exactly like time_embedder.py:17-18 so sin/cos arguments match the reference bit for bit. */
int mf_unet_set_time_freqs(mf_unet* h, const float* d_freqs, int n, mf_stream_t stream);
size_t mf_unet_workspace_bytes(mf_unet* h, int B, int H, int W);
/* y[B,out_ch,H,W] = UNet(x_t[B,in_ch,H,W], t[B] (int64), cond[B] (int64) or NULL).  NCHW fp32. */
int mf_unet_forward(mf_unet* h, const float* d_x_t, const int64_t* d_t, const int64_t* d_cond, float* d_y, int B,
                    int H, int W, void* d_workspace, size_t workspace_bytes, mf_stream_t stream);
/* mf_unet_forward + (ABI v3) timesteps as fp32 (d_t_float != NULL takes precedence; the reference's sinusoid accepts any
 * dtype, time_embedder.py:15-28) and the deep-supervision outputs: d_y_ver[k] (k < n_ver) = [B, out_ch, h_k, w_k] at the
 * resolution of level k+1, or NULL to skip that head (unet2.py:258-269 returns them as the second output). */
int mf_unet_forward_ex(mf_unet* h, const float* d_x_t, const int64_t* d_t, const float* d_t_float, const int64_t* d_cond,
                       float* d_y, float* const* d_y_ver, int n_ver, int B, int H, int W, void* d_workspace,
                       size_t workspace_bytes, mf_stream_t stream);
/* UNet.forward with the scheduler update fused into the epilogue of the output head (unet2.py:267 +
 * gaussian_scheduler.py:80-124 + diffusion_pipeline.py:244,297-304): the estimator output never round-trips through a
 * separate elementwise kernel.  d_y may be NULL when only the step outputs are wanted.  Requires out_ch <= 8. */
typedef struct {
  const mf_sched_tables* tables;
  const float* d_pred_uncond;   /* NULL: no classifier-free guidance */
  float guidance_scale;
  const float* d_noise;         /* scheduler draw, NULL = 0 */
  const int64_t* d_t_next;      /* NULL: no DDIM-form re-noise */
  const float* d_noise_ddim;
  int objective_is_x0;
  int clip_x0;
  float* d_x_prior; float* d_x_0; float* d_x_T; float* d_x_next;   /* any may be NULL */
  int uniform_t;                /* 1: all entries of d_t are equal (the sampling loop) -> the embedding MLP is evaluated
                                   once per class instead of once per sample */
} mf_step_args;
int mf_unet_forward_step(mf_unet* h, const float* d_x_t, const int64_t* d_t, const int64_t* d_cond, float* d_y, int B,
                         int H, int W, void* d_workspace, size_t workspace_bytes, const mf_step_args* step,
                         mf_stream_t stream);
/* Classifier-free guidance (diffusion_pipeline.py:240-244: two estimator passes at batch B) as ONE pass at batch 2B
 * (ABI v3): samples [0,B) run without / with the un_cond label, samples [B,2B) with the condition; the head combines
 * pred_u + guidance_scale * (pred_c - pred_u) and applies the scheduler update.  d_cond2: int64[2B], the unconditional
 * half first, where the value num_classes means "no label" (un_cond=None).  Needs step->uniform_t == 1 and
 * step->d_pred_uncond == NULL; workspace from mf_unet_workspace_bytes(h, 2*B, H, W).  Outputs have batch B. */
int mf_unet_forward_step_cfg(mf_unet* h, const float* d_x_t, const int64_t* d_t, const int64_t* d_cond2, int B, int H,
                             int W, void* d_workspace, size_t workspace_bytes, const mf_step_args* step,
                             mf_stream_t stream);
/* Same as mf_unet_forward but with a CUDA-event pair around every kernel launch of the plan (synchronises).
 * ms[i]: device time of launch i; kinds[i]: 0 conv_tc, 1 conv_simt, 2 GroupNorm family, 3 other;
 * flops[i]: algorithmic FLOPs of launch i (2*MACs of the reference formulation; 0 for non-GEMM work). */
int mf_unet_profile(mf_unet* h, const float* d_x_t, const int64_t* d_t, const int64_t* d_cond, float* d_y, int B, int H,
                    int W, void* d_workspace, size_t workspace_bytes, mf_stream_t stream, float* ms, int* kinds,
                    double* flops, int max_ops, int* n_ops);
/* Path census of the last prepared plan: number of convs on the tcgen05 path / on the SIMT path. */
int mf_unet_plan_info(const mf_unet* h, int* n_tc_convs, int* n_simt_convs, int* n_launches);

/* -------------------------------------------------------------------------------------------------
 * VAE decoder (latent_embedders.py:718-743 ctor, :764-769 decode)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int emb_channels;              /* latent channels (8) */
  int out_channels;              /* image channels (3) */
  int depth;                     /* len(hid_chs) */
  int hid_chs[MF_MAX_LEVELS];    /* [64,128,256,512] */
  int strides[MF_MAX_LEVELS];    /* [1,2,2,2]; levels 1.. must be 2 (nearest x2 + conv3x3) */
  int norm_groups;               /* 8 */
  int in_channels;               /* image channels of the ENCODER half (3); 0 = decoder-only handle (ABI v2) */
  int num_embeddings;            /* > 0: VQVAE (latent_embedders.py:180-320) — decode() first snaps z to the nearest row of
                                    `quantizer.embedder.weight` [num_embeddings][emb_channels] (VectorQuantizer.forward
                                    :50-69, z_q = z + (e - z)); needs in_channels == 0 (VQVAE.encode is not built) (ABI v3) */
} mf_vae_config;

typedef struct mf_vae mf_vae;

int mf_vae_create(const mf_vae_config* cfg, mf_vae** out);
void mf_vae_destroy(mf_vae* h);
int mf_vae_param_count(const mf_vae* h);
const char* mf_vae_param_name(const mf_vae* h, int i);
int mf_vae_param_shape(const mf_vae* h, int i, int64_t shape[4], int* ndim);
int mf_vae_set_param(mf_vae* h, const char* name, const float* d_data, const int64_t* shape, int ndim,
                     mf_stream_t stream);
size_t mf_vae_workspace_bytes(mf_vae* h, int B, int H, int W);
/* x[B,out_channels,H*2^(depth-1),W*2^(depth-1)] = decode(z[B,emb_channels,H,W]) */
int mf_vae_decode(mf_vae* h, const float* d_z, float* d_x, int B, int H, int W, void* d_workspace,
                  size_t workspace_bytes, mf_stream_t stream);
/* decode + the image post-processing of scripts/helpers/sample_dataset.py:47-50 fused into the output head:
 * d_x_u8[B][H'][W'][out_channels] = uint8((clip(x,-1,1)+1)/2*255); d_x (fp32 NCHW) may be NULL. */
int mf_vae_decode_u8(mf_vae* h, const float* d_z, float* d_x, uint8_t* d_x_u8, int B, int H, int W, void* d_workspace,
                     size_t workspace_bytes, mf_stream_t stream);
/* VAE.encode (latent_embedders.py:756-762): x [B,in_channels,H,W] -> z [B,emb,H/f,W/f] = mean + exp(0.5*clamp(logvar,-30,20))
 * * noise (DiagonalGaussianDistribution, :20-33).  d_noise: the caller's standard-normal draw of z's shape (the reference
 * calls torch.randn(mean.shape)); NULL gives z = mean.  d_moments (optional): [B,2*emb,h,w] = (mean | logvar).
 * The encoder has its own plan and workspace (mf_vae_encode_workspace_bytes). */
size_t mf_vae_encode_workspace_bytes(mf_vae* h, int B, int H, int W);
int mf_vae_encode(mf_vae* h, const float* d_x, const float* d_noise, float* d_z, float* d_moments, int B, int H, int W,
                  void* d_workspace, size_t workspace_bytes, mf_stream_t stream);
int mf_vae_profile(mf_vae* h, const float* d_z, float* d_x, int B, int H, int W, void* d_workspace,
                   size_t workspace_bytes, mf_stream_t stream, float* ms, int* kinds, double* flops, int max_ops,
                   int* n_ops);
int mf_vae_plan_info(const mf_vae* h, int* n_tc_convs, int* n_simt_convs, int* n_launches);

/* -------------------------------------------------------------------------------------------------
 * Scheduler step (all tensors [B, chw] fp32 contiguous; tables fp32[T] as registered by
 * gaussian_scheduler.py:44-58; t int64[B]; t_next int64 scalar or NULL)
 * ---------------------------------------------------------------------------------------------- */
/* mf_sched_tables is declared above (before the UNet section). */

int mf_sched_step(const mf_sched_tables* tables, const float* d_x_t, const float* d_pred,
                  const float* d_pred_uncond /* NULL: no guidance */, float guidance_scale, const int64_t* d_t,
                  const float* d_noise /* scheduler draw, NULL = 0 */, const int64_t* d_t_next /* NULL: no DDIM */,
                  const float* d_noise_ddim, int objective_is_x0, int clip_x0, float* d_x_prior, float* d_x_0,
                  float* d_x_T, float* d_x_next, int B, int chw, mf_stream_t stream);

/* The optional paths of DiffusionPipeline.forward (diffusion_pipeline.py:240-262, gaussian_scheduler.py:61-77,88-116):
 *  - learned variance (estimate_variance=True): the estimator returns [B, 2*C, H, W]; pass the tensor itself as d_pred
 *    (and d_pred_uncond), d_pred_var = d_pred + chw (the second chunk) and pred_batch_stride = 2*chw;
 *    std = exp(0.5 * (s*log(beta_t) + (1-s)*log(post_var_t))), s = v/2 + 0.5, v = guided variance channels;
 *  - cold_diffusion: x_prior = x_t - (x_t_est(t) - x_t_est(t-1)), no noise draw. */
typedef struct {
  const float* d_pred_var;          /* NULL: fixed small variance */
  const float* d_pred_var_uncond;   /* NULL unless guidance is active */
  int64_t pred_batch_stride;        /* 0 = chw */
  int cold_diffusion;
  const float* sqrt_alphas_cumprod;             /* fp32[T], cold diffusion only */
  const float* sqrt_one_minus_alphas_cumprod;   /* fp32[T], cold diffusion only */
  int T;
} mf_sched_opts;
int mf_sched_step_opts(const mf_sched_tables* tables, const float* d_x_t, const float* d_pred, const float* d_pred_uncond,
                       float guidance_scale, const int64_t* d_t, const float* d_noise, const int64_t* d_t_next,
                       const float* d_noise_ddim, int objective_is_x0, int clip_x0, float* d_x_prior, float* d_x_0,
                       float* d_x_T, float* d_x_next, int B, int chw, const mf_sched_opts* opts, mf_stream_t stream);

/* -------------------------------------------------------------------------------------------------
 * Kernel-level ops (test / profiling surface).  Layouts: 0 = NCHW fp32, 1 = NHWC fp32 ("raw"), 2 = NHWC "split":
 * two fp16 planes [2][N,H,W,C], hi = fp16(x), lo = fp16(x - hi), `plane` = elements between the planes.  Split
 * tensors are passed as void*.
 * ---------------------------------------------------------------------------------------------- */
int mf_op_pack_split(const float* d_x_nchw, void* d_out, int64_t plane, int N, int C, int H, int W, mf_stream_t s);
int mf_op_unpack_nchw(const void* d_in, int64_t plane, int layout, float* d_out_nchw, int N, int C, int H, int W,
                      mf_stream_t s);
/* OIHW fp32 -> fp16 [2][Cout][K] (K = 64-channel-slab major, tap minor), pre-scaled by 2^S; d_scales: device float[4]
 * receiving {2^S, 2^-S, scratch}.  Cin % 64 == 0. */
int mf_op_prep_weight_tc(const float* d_w_oihw, void* d_out, float* d_scales, int Cout, int Cin, int kh, int kw,
                         mf_stream_t s);
int mf_op_prep_weight_simt(const float* d_w_oihw, float* d_out, int Cout, int Cin, int kh, int kw, mf_stream_t s);
int mf_op_conv_tc_supported(int N, int H, int W, int C0, int C1, int Cout, int ksize, int stride);
/* 'same'-padded conv on the tcgen05 path (N,H,W = input size); stride 1 (1x1 / 3x3, optional second source
 * src1 concatenated along channels) or stride 2 (3x3, single source, even H/W).  src1 may be NULL (C1 = 0).
 * d_out: float NHWC (out_layout 1) or fp16 split planes (out_layout 2).  d_stats: [N][chunks][Cout/8][2] or NULL.
 * drain_interval: K blocks (of 64 channels) summed inside TMEM before the round-to-nearest fp32 register
 * accumulation; 0 = library default (2; 1 is the most exact, larger is faster). */
int mf_op_conv_tc(const void* d_src0, int64_t src0_plane, int C0, const void* d_src1, int64_t src1_plane, int C1,
                  int N, int H, int W, const void* d_w_planes, const float* d_scales, int Cout, int ksize,
                  const float* d_bias, void* d_out, int64_t out_plane, int out_layout, float* d_stats,
                  int drain_interval, int stride, mf_stream_t s);
int mf_op_conv_tc_stats_chunks(int H, int W);
/* conv3x3(nearest_x2(src)) + bias -> split [N,2H,2W,Cout]; weights from mf_op_prep_weight_up_tc (fp16 [2][4*Cout][4*C]) */
int mf_op_prep_weight_up_tc(const float* d_w_oihw, void* d_out, float* d_scales, int Cout, int Cin, mf_stream_t s);
int mf_op_upconv_tc(const void* d_src, int64_t src_plane, int C, int N, int H, int W, const void* d_w_up_planes,
                    const float* d_scales, int Cout, const float* d_bias, void* d_out, int64_t out_plane,
                    mf_stream_t s);
int mf_op_conv_simt(const void* d_in, int64_t in_plane, int in_layout, int N, int Cin, int Hin, int Win,
                    const float* d_w_kc, const float* d_bias, int Cout, int ksize, int stride, void* d_out,
                    int64_t out_plane, int out_layout, mf_stream_t s);
int mf_op_gn_partial(const float* d_raw, float* d_partial, int N, int HW, int C, mf_stream_t s);
int mf_op_gn_finalize(const float* d_partial, float* d_mean_rstd, int N, int chunks, int C, int G, int HW, float eps,
                      mf_stream_t s);
int mf_op_gn_apply(const float* d_raw, const float* d_mean_rstd, const float* d_gamma, const float* d_beta,
                   const void* d_res, int64_t res_plane, int res_kind /* 0 none, 1 split, 2 raw */,
                   const float* d_emb, int emb_stride, void* d_out, int64_t out_plane, int N, int HW, int C, int G,
                   mf_stream_t s);
/* attention-block kernels (attention_blocks.py): softmax((q s)^T (k s)) v per head with s = d^-0.25, q/k/v raw rows of
 * row_stride floats per token, out split [B*N][heads*d]; LayerNorm over channels (split -> split); GEGLU gate
 * (raw [tokens][2*Ch] -> split [tokens][Ch]) */
int mf_op_attention(const float* d_q, const float* d_k, const float* d_v, int row_stride, void* d_out, int64_t out_plane,
                    int B, int N, int heads, int d, mf_stream_t s);
int mf_op_layernorm(const void* d_in, int64_t in_plane, const float* d_gamma, const float* d_beta, void* d_out,
                    int64_t out_plane, int64_t tokens, int C, float eps, mf_stream_t s);
int mf_op_geglu(const float* d_in, void* d_out, int64_t out_plane, int64_t tokens, int Ch, mf_stream_t s);
int mf_op_upsample2x(const void* d_in, int64_t in_plane, void* d_out, int64_t out_plane, int N, int H, int W, int C,
                     mf_stream_t s);
/* VectorQuantizer.forward (latent_embedders.py:40-72), inference half: z [B,C,HW] fp32 NCHW -> z_q (same layout) =
 * z + (e[argmin_k ||z||^2 + ||e_k||^2 - 2 z.e_k] - z); d_idx (optional) receives the int32 code per latent vector. */
int mf_op_vq_quantize(const float* d_z, const float* d_codebook, float* d_zq, int* d_idx, int B, int C, int HW, int K,
                      mf_stream_t s);

#ifdef __cplusplus
}
#endif
#endif /* MEDFUSION_B200_H_ */
