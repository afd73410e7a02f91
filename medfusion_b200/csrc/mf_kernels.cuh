// Non-tensor-core kernels of the hot path: layout packing, exact-fp32 SIMT convolution for the
// narrow layers (Cin = 8 stems, Cout = 3/8 heads, stride-2 downsamplers), GroupNorm finalize/apply
// with fused Swish + residual + embedding add, nearest x2 upsample, timestep/label embedding MLP,
// and the scheduler update.  Reference call sites are cited at each launcher in mf_kernels.cu.
#pragma once

#include "mf_common.cuh"

namespace mf {

enum ActLayout : int {
  kNCHW = 0,       // contiguous fp32 [N,C,H,W]        (reference tensor layout at the API boundary)
  kNHWCRaw = 1,    // fp32 [N,H,W,C], one plane
  kNHWCSplit = 2,  // fp16 [2][N,H,W,C]: hi = fp16(x) plane + lo = fp16(x - hi) plane
};

// ---- layout / weight preparation -----------------------------------------------------------------
// Cpad > C: channels C..Cpad-1 of the NHWC output are zero (stem convolutions on the tensor-core path)
// n_mod > 0: sample n of the output reads sample n % n_mod of x (a 2B batch fed from B inputs)
int pack_nchw_to_split(const float* x, __half* out, long long plane, int N, int C, int H, int W, cudaStream_t s,
                       int Cpad = 0, int n_mod = 0);
int unpack_to_nchw(const void* in, long long plane, int in_layout, float* out, int N, int C, int H, int W,
                   cudaStream_t s);
// OIHW -> fp16 [2][Cout][K], K = ((c/64)*kh*kw + r*kw+s)*64 + c%64, pre-scaled by 2^S  (tensor-core path; Cin % 64 == 0)
// scales: device float[4], receives {2^S, 2^-S, scratch}
int prep_weight_tc(const float* w_oihw, __half* out, float* scales, int Cout, int Cin, int kh, int kw, cudaStream_t s,
                   int cin_pad = 0);
// OIHW 3x3 -> [2][4*Cout][4*Cin]: the four 2x2 phase kernels of conv3x3(nearest_x2(.)) with pre-summed taps
int prep_weight_up_tc(const float* w_oihw, __half* out, float* scales, int Cout, int Cin, cudaStream_t s);
// OIHW -> [K][Cout] fp32 (SIMT path)
int prep_weight_simt(const float* w_oihw, float* out, int Cout, int Cin, int kh, int kw, cudaStream_t s);

// ---- exact fp32 convolution on CUDA cores ---------------------------------------------------------
struct ConvSimtDesc {
  const void* in; long long in_plane; int in_layout;   // float (NCHW / raw) or __half planes (split)
  const void* in1; long long in1_plane; int C1;        // optional second source concatenated along channels (NHWC raw /
                                                       // split, same layout kind as `in`); Cin counts BOTH sources
  int N, Cin, Hin, Win;
  const float* w_kc;      // [K][Cout]
  const float* bias;      // [Cout] or nullptr
  int Cout, ksize, stride;  // pad = ksize/2
  void* out; long long out_plane; int out_layout;
};
int conv_simt(const ConvSimtDesc& d, cudaStream_t s);

// ---- GroupNorm ------------------------------------------------------------------------------------
// partial stats layout (shared with conv_tc epilogue): [N][chunks][C/8][2] = (sum, sumsq) over 8 channels
// plane != 0: `raw` is a split tensor (x = hi + lo)
int gn_partial_from_raw(const void* raw, float* partial, int N, int HW, int C, cudaStream_t s, long long plane = 0);
// partial -> (mean, rstd) per (n, group): out [N][G][2]
int gn_finalize(const float* partial, float* mean_rstd, int N, int chunks, int C, int G, int HW, float eps,
                cudaStream_t s);
enum ResKind : int { kResNone = 0, kResSplit = 1, kResRaw = 2 };
struct GnApplyDesc {
  const void* raw;          // [N,HW,C] conv output (+bias), float; or a split tensor (__half planes) when raw_plane != 0
  long long raw_plane;      // != 0: input is a split tensor (hi + lo)
  int act;                  // 1: Swish after the affine (conv blocks), 0: none (attention-block norms)
  const float* mean_rstd;   // [N][G][2]  (unused when `partial` is given)
  const float* partial; int chunks; float eps;   // != nullptr: per-(sample, chunk, 8-channel slab) (sum, sumsq); the kernel
                                                 // finalises the statistics itself (no gn_finalize launch)
  const float* gamma; const float* beta;  // [C]
  const void* res; long long res_plane; int res_kind;   // split: __half planes, raw: float
  const float* emb; int emb_stride;       // emb[row*emb_stride + c] or nullptr; row = emb_index ? emb_index[n] : n
  const long long* emb_index;             // optional row indirection (deduplicated embeddings: one row per class)
  __half* out; long long out_plane;       // split planes (nullptr with a folded head: the activation is never written)
  int N, HW, C, G;
  // optional folded narrow 1x1 head (fused variant only; latent_embedders.py:743 `outc`, 64 -> 3): the C/8 threads that hold
  // one pixel's channels reduce head_cout dot products among themselves and write NCHW fp32 (and / or the uint8 HWC image
  // of scripts/helpers/sample_dataset.py:47-50) — the last activation (1.07 GB at B=64, 256x256) is neither written nor re-read
  const float* head_w; const float* head_b;   // [head_cout][C], [head_cout]
  float* head_out; unsigned char* head_out_u8;
  int head_cout;                          // 0: no head
};
int gn_apply(const GnApplyDesc& d, cudaStream_t s);
// Any channel count / any channels-per-group (the fused path needs C/G % 8 == 0): statistics straight from the raw conv
// output, one block per (sample, group), fp64 combine -> mean_rstd [N][G][2]; then gn_apply_generic (one channel per
// thread-iteration).  Used by the small-width configurations (e.g. VQVAE defaults: 32 channels in 32 groups).
int gn_stats_generic(const float* raw, float* mean_rstd, int N, int HW, int C, int G, float eps, cudaStream_t s);
int gn_apply_generic(const GnApplyDesc& d, cudaStream_t s);
// VectorQuantizer.forward (latent_embedders.py:40-72), inference half: z [B,C,HW] NCHW -> nearest codebook row by
// ||z||^2 + ||e||^2 - 2 z.e (first minimum wins, like torch.argmin), z_q = z + (e - z) as the reference evaluates it.
// codebook [K][C]; idx_out optional [B*HW] int32
int vq_quantize(const float* z, const float* codebook, float* z_q, int* idx_out, int B, int C, int HW, int K,
                cudaStream_t s);
extern int g_pdl;         // 1: conv_tc and the fused gn_apply are launched with programmatic stream serialization
extern int g_gn_variant;  // engine plans: 3 (default) = statistics finalised inside gn_apply (one launch per GroupNorm);
                          // 0/1/2 = separate gn_finalize + flat grid-stride / fixed quad per thread / one quad per thread

// nearest x2 upsample of a split tensor (reference: conv_blocks.py:123-125, F.interpolate nearest-exact)
int upsample2x_split(const __half* in, long long in_plane, __half* out, long long out_plane, int N, int H, int W, int C,
                     cudaStream_t s);

// ---- embedding MLP --------------------------------------------------------------------------------
// out[b][j] = post( sum_k W[j][k] * in[b][k] + bias[j] + add[b][j] );  W row-major [J][K]
//   in_mode 0: in is a float [B][K] matrix;  in_mode 1: in is built on the fly as sinusoidal(t[b]) with freqs[K/2]
//   post 0: identity, 1: swish;  out2 (optional) receives swish(out)
struct LinearDesc {
  const float* in; const long long* t; int t_stride; const float* freqs; int in_mode;  // t[b * t_stride]
  const float* t_float;   // != nullptr: timesteps given as fp32 (the reference accepts any dtype, time_embedder.py:15-28)
  const float* W; const float* bias;
  const float* add_table; const long long* add_idx;  // optional embedding-table add: add_table[add_idx ? add_idx[b] : b][j]
  float* out; float* out2; int post;
  int B, J, K;
};
int linear_small(const LinearDesc& d, cudaStream_t s);

// ---- scheduler step -------------------------------------------------------------------------------
struct SchedTables {  // device pointers, length T each (fp32), computed in fp64 on the host like the reference
  const float* sqrt_recip_ac; const float* sqrt_recipm1_ac; const float* coef1; const float* coef2;
  const float* post_var; const float* betas; const float* alphas_cumprod;
};
struct SchedStepDesc {
  const float* x_t; const float* pred; const float* pred_uncond; float guidance;  // pred_uncond nullable
  const long long* t;        // [B]
  const float* noise;        // randn_like(x_t) (scheduler draw); nullable -> treated as 0
  const long long* t_next;   // scalar device int64 (DDIM re-noise) or nullptr
  const float* noise2;       // DDIM draw
  int objective_x0;          // 0: estimator predicts x_T (noise), 1: predicts x_0
  int clip_x0;
  float* x_prior; float* x_0; float* x_T; float* x_next;  // outputs (x_next only with t_next)
  int B, CHW;
  SchedTables tab;
  // optional paths of DiffusionPipeline.forward (diffusion_pipeline.py:240-262, gaussian_scheduler.py:88-116)
  const float* pred_var; const float* pred_var_uncond;   // learned variance channels (estimate_variance=True) or nullptr
  long long pred_bstride;    // elements between samples of pred / pred_uncond / pred_var (0: CHW; 2*CHW for chunk(2, dim=1))
  int cold;                  // cold-diffusion update (no noise draw)
  const float* sqrt_ac; const float* sqrt_1mac; int T;   // tables estimate_x_t needs (cold diffusion only)
};
int sched_step(const SchedStepDesc& d, cudaStream_t s);
// z = mean + exp(0.5*clamp(logvar,-30,20)) * noise from NCHW moments [B,2E,HW] (noise nullptr: z = mean)
int vae_reparam(const float* moments, const float* noise, float* z, float* moments_out, int B, int EHW, cudaStream_t s);

// ---- narrow 1x1 head (Cout <= 8): split NHWC -> NCHW fp32, optional fused scheduler step ------------------------
struct HeadDesc {
  const __half* in; long long in_plane;  // split [N][HW][C]
  const float* w; const float* bias;     // [Cout][C] (reference OIHW layout of a 1x1 conv), [Cout]
  float* out;                            // NCHW [N][Cout][HW] or nullptr (only the fused step outputs are wanted)
  unsigned char* out_u8;                 // optional NHWC uint8 image [N][HW][Cout]: clip(-1,1) -> (x+1)/2*255 -> truncate
  int N, HW, C, Cout;
  int fuse_step;
  int cfg_pair;                          // 1: `in` holds 2N samples (uncond | cond); outputs N samples of the CFG combine
  float cfg_guidance;
};
// step != nullptr: every output element is fed to the scheduler update as `pred` (step->pred is ignored)
int head1x1(const HeadDesc& h, const SchedStepDesc* step, cudaStream_t s);

// saturation counters of the three kernel translation units (see mf_common.cuh); *total is incremented
int sat_read_kernels(unsigned long long* total, int reset, cudaStream_t s);
int sat_read_conv_tc(unsigned long long* total, int reset, cudaStream_t s);
int sat_read_attn(unsigned long long* total, int reset, cudaStream_t s);

}  // namespace mf
