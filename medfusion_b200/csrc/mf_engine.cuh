// Host-side execution engine shared by the UNet and VAE-decoder C-ABI objects: parameter registry
// (reference state_dict names), workspace arena, and a static launch plan per (B, H, W, workspace).
#pragma once

#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "mf_attn.cuh"
#include "mf_conv_tc.cuh"
#include "mf_kernels.cuh"

namespace mf {

struct DevBuf {
  float* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  int alloc(size_t elems) {
    if (elems == n && p) return 0;
    release();
    if (cudaMalloc(&p, elems * sizeof(float)) != cudaSuccess) {
      set_error("cudaMalloc failed for " + std::to_string(elems * sizeof(float)) + " bytes");
      return 1;
    }
    n = elems;
    return 0;
  }
};

struct Param {
  std::string name;
  std::vector<int64_t> shape;
  DevBuf data;  // reference layout (conv OIHW, linear [out,in], vectors)
  bool is_set = false;
  size_t numel() const {
    size_t n = 1;
    for (auto d : shape) n *= static_cast<size_t>(d);
    return n;
  }
};

struct ConvLayer {
  Param* w = nullptr;
  Param* b = nullptr;
  int Cout = 0, Cin = 0, k = 1, stride = 1;
  DevBuf w_tc, w_simt, w_up;  // derived layouts, built lazily by the plan builder (w_tc / w_up hold fp16 planes)
  DevBuf tc_scales, up_scales;  // device float[4] each: {2^S, 2^-S, scratch} of the power-of-two weight pre-scale
  int tc_version = -1, simt_version = -1, up_version = -1;
};
struct NormLayer {
  Param* g = nullptr;
  Param* b = nullptr;
  int C = 0;
};
struct ResBlockLayer {
  ConvLayer conv1, conv2, conv_res;
  NormLayer norm1, norm2;
  bool has_res_conv = false;
  int emb_offset = -1;  // row offset in the fused local-embedder matrix, -1: no embedding
  Param* emb_w = nullptr;
  Param* emb_b = nullptr;
  int Cin = 0, Cout = 0;
};

// LinearTransformer (attention_blocks.py:128-195): GroupNorm + q/k/v/out 1x1 projections
struct LinAttnLayer {
  NormLayer norm_x;
  ConvLayer to_q, to_k, to_v, to_out;
  int C = 0, kv_in = 0;                 // kv_in: channels feeding k/v (C for self-attention, emb_dim for cross)
  // fused [3C][C] q|k|v projection (self-attention only), engine-owned copies of the three registered tensors
  std::unique_ptr<Param> qkv_w, qkv_b;
  ConvLayer qkv;
  int qkv_version = -1;
};
// SpatialTransformer with one BasicTransformerBlock (attention_blocks.py:200-288)
struct SpatialAttnLayer {
  int kind = 0;                         // 0 none, 1 'linear', 2 'spatial'
  int C = 0, heads = 8, d = 0, emb_dim = 0;
  NormLayer norm;                       // .norm
  ConvLayer proj_in, proj_out;          // .proj_in / .proj_out (1x1)
  LinAttnLayer self_atn, cros_atn;      // .transformer_blocks.0.{self_atn,cros_atn}  ('linear': cros_atn only)
  NormLayer ln;                         // .transformer_blocks.0.proj_out.0.norm (LayerNorm over C)
  ConvLayer ff_in, ff_out;              // GEGLU Linear C -> 8C, then 1x1 conv 4C -> C
};

// ---- workspace arena (static planning; single-stream ordering makes reuse safe) ------------------
struct Arena {
  struct Block { size_t off, size; };
  std::vector<Block> free_list;
  size_t top = 0, peak = 0;
  size_t alloc(size_t bytes);
  void release(size_t off, size_t bytes);
};

struct Tens {
  size_t off = 0, bytes = 0;
  int N = 0, H = 0, W = 0, C = 0;
  int layout = kNHWCSplit;
  long long plane = 0;  // elements between hi and lo plane
  float* ptr = nullptr; // raw tensors: float data; split tensors: fp16 planes (use hptr())
  __half* hptr() const { return reinterpret_cast<__half*>(ptr); }
  long long elems() const { return static_cast<long long>(N) * H * W * C; }
};

extern int g_fold_head;
extern int g_fold_upsample;
extern int g_stem_on_tc;
using Launch = std::function<int(cudaStream_t)>;
enum OpKind : int { kOpConvTc = 0, kOpConvSimt = 1, kOpNorm = 2, kOpOther = 3 };
struct OpMeta { int kind; double flops; };

class EngineBase {
 public:
  virtual ~EngineBase() { streamk_scratch_free(&sk_scratch); }
  // stream-K partial tiles of THIS engine's convolutions, on the device the engine first ran on (a process-
  // global scratch was shared by every engine, stream and device)
  StreamKScratch sk_scratch;
  int device = -1;     // device the parameters / plans live on (set by the first set_param / build)
  int bind_device();   // records the current device on first use, errors if a later call runs on another one
  const StreamKScratch* scratch();
  std::vector<std::unique_ptr<Param>> params;
  std::map<std::string, Param*> by_name;
  int version = 0;  // bumped by set_param; invalidates derived weight layouts and plans

  Param* add_param(const std::string& name, std::vector<int64_t> shape);
  int set_param(const char* name, const float* d_data, const int64_t* shape, int ndim, cudaStream_t s);
  int check_all_set() const;

  // ---- plan state
  struct PlanKey { int B = -1, H = -1, W = -1; void* base = nullptr; int version = -1; } key;
  std::vector<Launch> ops;
  std::vector<OpMeta> op_meta;  // parallel to ops: kernel family + algorithmic FLOPs (profiling / roofline)
  void push_op(Launch l, int kind, double flops = 0.0) {
    ops.push_back(std::move(l));
    op_meta.push_back({kind, flops});
  }
  std::vector<std::unique_ptr<ConvTcPlan>> tc_plans;
  int n_tc = 0, n_simt = 0;

  // optional scheduler update fused into the narrow output head (set per call by the C ABI)
  SchedStepDesc io_step{};
  bool io_step_on = false;
  // embedding rows deduplicated for this call (all samples share t): row index = class id, or one shared row
  unsigned char* io_out_u8 = nullptr;   // optional uint8 HWC copy of the narrow head's output (VAE images)
  bool io_emb_dedup = false;
  const long long* io_emb_index = nullptr;
  // classifier-free guidance as ONE 2B batch (diffusion_pipeline.py:240-244 runs two B passes): the stem packs every input
  // sample twice, the label index of the first half points at the all-zero "no label" row, the head combines the halves
  bool io_cfg_pair = false;
  float io_cfg_guidance = 1.f;

  // ---- builder state (valid during build())
  bool dry = false;
  char* base = nullptr;
  Arena arena;
  cudaStream_t prep_stream = nullptr;

  Tens new_tensor(int N, int H, int W, int C, int layout);
  Tens new_floats(size_t n);
  void free_tensor(const Tens& t);

  // conv: in0 (+ optional in1 channel-concatenated) -> out.  If stats != nullptr the (sum, sumsq)
  // partials of the output are produced too and *chunks receives their chunk count.
  // gn != nullptr: GroupNorm + Swish + gn->res + gn->emb applied in the conv's epilogue (tensor-core path, geometry
  // permitting: conv_gn_fusable()); `out` is then the split result of the res-block half.
  struct GnFuse { const NormLayer* nl; int groups; const Tens* res; const float* emb; int emb_stride; };
  bool conv_gn_fusable(const ConvLayer& L, const Tens& in0, const Tens* in1, int Ho, int Wo, int groups) const;
  int add_conv(ConvLayer& L, const Tens& in0, const Tens* in1, const Tens& out, const Tens* stats, int* chunks,
               const Tens* res = nullptr, const float* emb = nullptr, int emb_stride = 0, const GnFuse* gn = nullptr);
  // GroupNorm (no activation) of a split tensor -> split
  int add_group_norm_split(const NormLayer& nl, int groups, const Tens& x, const Tens& out);
  // attention block applied to x (split) -> *out (split).  emb: raw [B, emb_dim] embedding (may be null)
  int add_attention(SpatialAttnLayer& A, int groups, const Tens& x, const Tens* emb, Tens* out);
  int ensure_qkv(LinAttnLayer& L);
  // conv reading an external NCHW fp32 pointer that is only known at call time
  int add_conv_nchw_in(ConvLayer& L, const float* const* src, int N, int Cin, int H, int W, const Tens& out,
                       const Tens* stats, int* chunks);
  // conv writing an external NCHW fp32 pointer only known at call time
  // in1: optional second source (channel concat).  A NULL *dst at launch time skips the op (optional outputs).
  int add_conv_nchw_out(ConvLayer& L, const Tens& in0, float* const* dst, const Tens* in1 = nullptr);
  // BasicUp: out[2H,2W] = conv3x3(nearest_x2(in)) as four 2x2 phase convolutions on the tensor-core path, or
  // (shapes the tensor-core path cannot take) an explicit upsample followed by add_conv.
  int add_upconv2x(ConvLayer& L, const Tens& in, Tens* out);
  // head != nullptr: the narrow 1x1 conv `head` (Cout <= 8) is evaluated inside the apply kernel and written to *head_dst
  // (NCHW fp32, resolved at launch) and/or io_out_u8; `out` is then not written (pass a Tens with ptr == nullptr)
  int add_gn_apply(const NormLayer& nl, int groups, const Tens& raw, const Tens& stats, int chunks, const Tens* res,
                   const float* emb, int emb_stride, const Tens& out, int act = 1, ConvLayer* head = nullptr,
                   float* const* head_dst = nullptr);
  // can the head be folded into the GroupNorm-apply of a C-channel tensor?
  bool can_fold_head(const ConvLayer& head, int C, int groups) const;
  // full res block: in0 (+in1 concat) -> returns output tensor (split).  head != nullptr (and foldable): the block's
  // output only feeds that head, which is folded into the last GroupNorm-apply; *out is then an empty tensor.
  int add_resblock(ResBlockLayer& rb, int groups, const Tens& in0, const Tens* in1, const Tens* embT, int emb_stride,
                   Tens* out, ConvLayer* head = nullptr, float* const* head_dst = nullptr);
  int ensure_w_tc(ConvLayer& L, int cin_pad = 0);
  int ensure_w_simt(ConvLayer& L);
  int run(cudaStream_t s);
  // run with a CUDA event pair around every op; synchronises. ms[i] receives the device time of op i.
  int run_profiled(cudaStream_t s, float* ms, int* kinds, double* flops, int max_ops, int* n_ops);
};

bool gn_needs_generic(int C, int groups);   // channels per group not a multiple of 8: generic statistics / apply kernels
void init_conv(EngineBase& e, ConvLayer& L, const std::string& prefix, int Cout, int Cin, int k, int stride);
void init_norm(EngineBase& e, NormLayer& L, const std::string& prefix, int C);
void init_conv_shape(EngineBase& e, ConvLayer& L, const std::string& prefix, std::vector<int64_t> wshape, int Cout,
                     int Cin);
void init_attention(EngineBase& e, SpatialAttnLayer& A, const std::string& prefix, int kind, int C, int emb_dim);
void init_resblock(EngineBase& e, ResBlockLayer& rb, const std::string& prefix, int Cin, int Cout, int k, int emb_dim);

}  // namespace mf
