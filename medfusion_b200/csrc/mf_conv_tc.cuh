// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a), error-compensated fp16x3 (fp32-accurate).
//
// Replaces the reference's nn.Conv2d call sites on the hot path
//   /root/reference/medical_diffusion/models/utils/conv_blocks.py:185   (BasicBlock.conv, 3x3 s1 p1)
//   /root/reference/medical_diffusion/models/utils/conv_blocks.py:224   (BasicResBlock.conv_res, 1x1)
//   /root/reference/medical_diffusion/models/utils/conv_blocks.py:104   (BasicUp.up_op, 3x3 after nearest x2)
// and the torch.cat of unet2.py:259 (two-source K loop instead of a materialised concat).
//
// GEMM view:  D[M = N*H*W output pixels][Cout] = sum_{tap, c} A[pixel shifted by tap][c] * Wt[Cout][tap*Cin + c]
// Activations live in HBM as NHWC fp16 in two planes (hi = fp16(x), lo = fp16(x - hi)); weights as
// [plane][Cout][K] fp16 (pre-scaled by a power of two) with K = 64-channel-slab major, tap minor.  One CTA computes a 128-pixel x BLOCK_N tile:
//   warp 0      : TMA producer  (5-D activation boxes with zero-filled halo, 3-D weight boxes)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (3 MMAs per K-step: hi*hi, hi*lo, lo*hi)
//   warps 4..11 : drain (tcgen05.ld of each finished TMEM partial sum -> round-to-nearest fp32 register
//                 accumulation) and epilogue (+bias -> GroupNorm partial sums -> global store)
#pragma once

#include "mf_common.cuh"

namespace mf {

constexpr int kTcBlockM = 128;  // output pixels per tile (UMMA M)
constexpr int kTcBlockK = 64;   // fp16 elements per K block = 128 B = one swizzle row
constexpr int kTcDrainWarps = 8;
constexpr int kTcThreads = 128 + 32 * kTcDrainWarps;  // warpgroup 0: TMA warp, MMA warp, 2 idle; warpgroups 1-2: drain
constexpr int kTcMaxTaps = 9;

enum ConvOutMode : int {
  kOutRaw = 0,    // single fp32 plane (input of GroupNorm)
  kOutSplit = 1,  // fp16 hi/lo planes (direct input of the next conv)
};

struct ConvTcParams {
  int N, H, W;                   // OUTPUT geometry
  int bw, bh, bn;                // pixel box of one tile, bw*bh*bn == 128
  int tiles_w, tiles_h, tiles_n;
  int m_groups, n_tiles, num_tiles;  // CTA-group tiles along M, tiles along N, total (x4 phases with up2)
  float* sk_partials;            // stream-K scratch [grid CTAs][128][256] fp32
  int* sk_flags;                 // stream-K flags   [grid CTAs]
  int C0, C1;                    // channels of source 0 / 1 (C1 == 0: single source)
  int Cout;
  int ntaps;
  int dy[kTcMaxTaps], dx[kTcMaxTaps];
  int per_tap_map;               // 0: source chosen by channel (concat); 1: source map chosen per tap (stride 2)
  int tap_map[kTcMaxTaps];       // stride 2: which input-parity tensor map a tap reads
  int up2;                       // 1: nearest-x2 upsample folded in (4 phases in blockIdx.z, 2x2 taps each)
  int drain_interval;            // K blocks accumulated inside TMEM before the fp32 register add (1 = most exact)
  float partial_scale;           // (1 + eps): de-biases the truncating TMEM accumulation when a partial is drained
  const float* bias;             // [Cout] or nullptr
  void* out;                     // NHWC [N,H,W,Cout]: float (raw mode) or __half hi plane (split mode)
  long long out_plane;           // elements between hi and lo plane (split mode)
  const float* w_inv_scale;      // device scalar 2^-S undoing the power-of-two weight pre-scale
  int out_mode;
  float* stats;                  // [N][chunks][Cout/8][2] partial (sum, sumsq) or nullptr
  const void* res;               // optional residual with the output's geometry; res_kind 1: split (__half) planes, 2: raw float
  long long res_plane;
  int res_kind;
  const float* emb;              // optional per-sample channel vector emb[n*emb_stride + c]
  int emb_stride;
  int chunks_per_sample;         // tiles per sample (1 if a tile spans >= 1 whole samples)
  int rows_per_sample;           // min(128, H*W)
  // ---- GroupNorm + Swish + residual (+ embedding) applied in the epilogue (conv_blocks.py:184-192,236-240,362): possible
  //      when the finalising CTA (gn_mode 1, H*W <= 128) or CTA pair (gn_mode 2, H*W == 256: the two CTAs swap their
  //      per-slab sums through distributed shared memory) holds every pixel of a sample.  The raw fp32 conv output and
  //      the separate gn_apply launch disappear; `out` receives the split planes of the block's result.
  int gn_mode;                   // 0: off
  int gn_cpg;                    // channels per group (multiple of 8, divides BLOCK_N)
  float gn_eps;
  const float* gn_gamma; const float* gn_beta;   // [Cout]
  const void* gn_res; long long gn_res_plane; int gn_res_kind;   // residual added after the activation (1 split, 2 raw fp32)
  const float* gn_emb; int gn_emb_stride; const long long* gn_emb_index;   // per-sample channel vector added last
};

struct TcMaps {
  CUtensorMap a[4];  // stride 1: a[0] = source 0, a[1] = source 1 (concat);  stride 2: a[2*ph + pw] = input parity grid
  CUtensorMap w;
  CUtensorMap o[4];  // output (TMA store): o[0]; folded upsample: o[2*oa + ob] = the output-parity grid of phase (oa, ob).
                     // 5-D {C, W, H, N, plane}: fp32 x 1 plane (raw, 128-byte swizzle) or fp16 x 2 planes (split, 64-byte)
};

struct ConvTcPlan {
  TcMaps maps;
  ConvTcParams p;
  int block_n;
  int cta_group;  // 1: one CTA per tile; 2: CTA pairs (cluster of 2) sharing the weight tile
  int row3;       // 1: row-patch mode (one stage = the 130-pixel patch of one input row + the weight tiles of its 3 taps)
  dim3 grid;
};

// Host API -------------------------------------------------------------------------------------
struct ConvTcDesc {
  // sources: NHWC split tensors (fp16 hi plane pointer, lo = hi + plane elements)
  const __half* src0; long long src0_plane; int C0;
  const __half* src1; long long src1_plane; int C1;  // C1 == 0 -> none
  int N, H, W;              // INPUT spatial size; output is H/stride x W/stride
  int stride;               // 1 (default when 0) or 2 (3x3, single source, even H and W)
  const __half* w_planes;   // [2][Cout][K] fp16 prepared by prep_weight_tc (pre-scaled by 2^S)
  const float* w_inv_scale; // device scalar 2^-S written by prep_weight_tc
  int Cout, ksize;          // ksize 1 or 3 (pad = ksize/2, stride 1)
  const float* bias;
  void* out; long long out_plane; int out_mode;   // float* (raw) or __half* (split)
  float* stats;             // optional
  const void* res; long long res_plane; int res_kind;   // optional fused residual add (1 split fp16, 2 raw fp32)
  const float* emb; int emb_stride;                     // optional fused per-sample channel add
  int drain_interval;       // 0 -> default (1)
  int cta_group;            // 0 -> default (auto), 1 or 2
  int block_n;              // 0 -> default (auto), 64 / 128 / 256 output channels per tile
  int up2;                  // 1: input is HxW, output 2Hx2W = conv3x3(nearest_x2(input)); w_planes from prep_weight_up_tc
  // fused GroupNorm epilogue (see ConvTcParams): gn_groups > 0 requests it; conv_tc_gn_fusable() says whether the
  // geometry allows it.  The residual / embedding are those of the res block half (added AFTER norm + Swish).
  int gn_groups; const float* gn_gamma; const float* gn_beta; float gn_eps;
  const void* gn_res; long long gn_res_plane; int gn_res_kind;
  const float* gn_emb; int gn_emb_stride;
  // stream-K scratch owned by the caller (an engine keeps its own, so two engines on different streams or devices never
  // share partial-sum tiles); nullptr: the per-device scratch of the library (stand-alone mf_op_* calls)
  const struct StreamKScratch* scratch;
};

// one [128][256] fp32 partial tile and one flag per CTA of the persistent grid
struct StreamKScratch {
  float* partials = nullptr;
  int* flags = nullptr;
  int max_ctas = 0;   // SM count of the device the scratch lives on
  int device = -1;
};
// allocate on the CURRENT device (flags zeroed); release with streamk_scratch_free
int streamk_scratch_alloc(StreamKScratch* sc);
void streamk_scratch_free(StreamKScratch* sc);

extern int g_default_drain_interval;
extern int g_default_cta_group;
extern int g_default_block_n;
extern int g_stream_k;
extern int g_split_fill;
extern int g_row_patch;
extern int g_pdl;   // defined in mf_kernels.cu
extern float g_debias_eps_per_kblock;
int conv_tc_supported(int N, int H, int W, int C0, int C1, int Cout, int ksize, int stride);
// can GroupNorm(groups) + Swish + residual be applied inside the epilogue of this convolution (OUTPUT geometry H, W)?
int conv_tc_gn_fusable(int H, int W, int Cout, int groups);
extern int g_fuse_gn;   // 1: engines use the fused GroupNorm epilogue wherever conv_tc_gn_fusable() allows (default 0)
struct ConvTcShapePlan { int cta_group, block_n, groups, row3, m_groups, n_tiles, num_tiles, nkb; };
// the schedule of a layer as a pure function of its shape and the knobs (no device access)
int conv_tc_plan_shape(int m_tiles, int bw, int bh, int bn_box, int C0, int C1, int Cout, int ksize, int stride, int up2,
                       int gn_mode, int forced_cg, int forced_bn, int drain_interval, int max_ctas, ConvTcShapePlan* out);
int conv_tc_plan_query(int N, int H, int W, int C0, int C1, int Cout, int ksize, int stride, int up2, int sm_count, int* out8);
int conv_tc_build(const ConvTcDesc& d, ConvTcPlan* plan);
// emb_dedup: the fused-GroupNorm embedding rows are deduplicated for this call (row = emb_index[n], or one shared row)
int conv_tc_launch(const ConvTcPlan& plan, cudaStream_t stream, int emb_dedup = 0, const long long* emb_index = nullptr);
// number of chunks (tiles per sample) the stats buffer must provide for this geometry
int conv_tc_stats_chunks(int H, int W);

}  // namespace mf
