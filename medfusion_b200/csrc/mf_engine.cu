#include "mf_engine.cuh"

#include <algorithm>

namespace mf {

int g_fold_head = 1;        // VAE image head (64 -> 3) folded into the last GroupNorm-apply
int g_fold_upsample = 1;
int g_stem_on_tc = 1;      // Cin < 64 stem convolutions on the tensor core through a zero-padded 64-channel input  // BasicUp as four phase convolutions (0: explicit nearest-x2 kernel + conv3x3)

// ---- error string --------------------------------------------------------------------------------
static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }
const char* get_error() { return g_error.c_str(); }

// ---- arena ---------------------------------------------------------------------------------------
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

size_t Arena::alloc(size_t bytes) {
  bytes = align_up(std::max<size_t>(bytes, 1), 1024);
  // best fit among the free blocks
  int best = -1;
  for (int i = 0; i < static_cast<int>(free_list.size()); ++i)
    if (free_list[i].size >= bytes && (best < 0 || free_list[i].size < free_list[best].size)) best = i;
  if (best >= 0) {
    const size_t off = free_list[best].off;
    if (free_list[best].size == bytes) free_list.erase(free_list.begin() + best);
    else { free_list[best].off += bytes; free_list[best].size -= bytes; }
    return off;
  }
  const size_t off = top;
  top += bytes;
  peak = std::max(peak, top);
  return off;
}

void Arena::release(size_t off, size_t bytes) {
  bytes = align_up(std::max<size_t>(bytes, 1), 1024);
  free_list.push_back({off, bytes});
  std::sort(free_list.begin(), free_list.end(), [](const Block& a, const Block& b) { return a.off < b.off; });
  std::vector<Block> merged;
  for (const Block& b : free_list) {
    if (!merged.empty() && merged.back().off + merged.back().size == b.off) merged.back().size += b.size;
    else merged.push_back(b);
  }
  if (!merged.empty() && merged.back().off + merged.back().size == top) {
    top = merged.back().off;
    merged.pop_back();
  }
  free_list.swap(merged);
}

// ---- device binding -------------------------------------------------------------------------------
int EngineBase::bind_device() {
  int dev = 0;
  MF_CUDA_OK(cudaGetDevice(&dev));
  if (device < 0) device = dev;
  MF_REQUIRE(device == dev, "this engine handle lives on device " + std::to_string(device) + " but the current device is " +
                                std::to_string(dev) + " (one handle per device; set the device before calling)");
  return 0;
}
const StreamKScratch* EngineBase::scratch() {
  if (sk_scratch.partials == nullptr && streamk_scratch_alloc(&sk_scratch) != 0) return nullptr;
  return &sk_scratch;
}

// ---- parameters ----------------------------------------------------------------------------------
Param* EngineBase::add_param(const std::string& name, std::vector<int64_t> shape) {
  params.emplace_back(new Param());
  Param* p = params.back().get();
  p->name = name;
  p->shape = std::move(shape);
  by_name[name] = p;
  return p;
}

int EngineBase::set_param(const char* name, const float* d_data, const int64_t* shape, int ndim, cudaStream_t s) {
  auto it = by_name.find(name);
  if (it == by_name.end()) {
    set_error(std::string("unknown parameter '") + name + "'");
    return 2;
  }
  if (int rcd = bind_device()) return rcd;
  Param* p = it->second;
  bool same = static_cast<int>(p->shape.size()) == ndim;
  for (int i = 0; same && i < ndim; ++i) same = p->shape[i] == shape[i];
  if (!same) {
    std::string want, got;
    for (auto d : p->shape) want += std::to_string(d) + ",";
    for (int i = 0; i < ndim; ++i) got += std::to_string(shape[i]) + ",";
    set_error("shape mismatch for '" + p->name + "': expected [" + want + "] got [" + got + "]");
    return 2;
  }
  if (p->data.alloc(p->numel())) return 1;
  MF_CUDA_OK(cudaMemcpyAsync(p->data.p, d_data, p->numel() * sizeof(float), cudaMemcpyDeviceToDevice, s));
  p->is_set = true;
  ++version;
  return 0;
}

int EngineBase::check_all_set() const {
  for (const auto& p : params)
    if (!p->is_set) {
      set_error("parameter '" + p->name + "' was never set (load_state_dict before forward)");
      return 2;
    }
  return 0;
}

void init_conv(EngineBase& e, ConvLayer& L, const std::string& prefix, int Cout, int Cin, int k, int stride) {
  L.w = e.add_param(prefix + ".weight", {Cout, Cin, k, k});
  L.b = e.add_param(prefix + ".bias", {Cout});
  L.Cout = Cout; L.Cin = Cin; L.k = k; L.stride = stride;
}
// conv-like layer whose registered weight has a non-OIHW shape with the same memory layout
// (Conv1d [Cout,Cin,1], Linear [Cout,Cin]); used as a 1x1 convolution
void init_conv_shape(EngineBase& e, ConvLayer& L, const std::string& prefix, std::vector<int64_t> wshape, int Cout,
                     int Cin) {
  L.w = e.add_param(prefix + ".weight", std::move(wshape));
  L.b = e.add_param(prefix + ".bias", {Cout});
  L.Cout = Cout; L.Cin = Cin; L.k = 1; L.stride = 1;
}
static void init_lin_attn(EngineBase& e, LinAttnLayer& L, const std::string& prefix, int C, int kv_in) {
  L.C = C; L.kv_in = kv_in;
  init_norm(e, L.norm_x, prefix + ".norm_x", C);
  init_conv_shape(e, L.to_q, prefix + ".to_q", {C, C, 1}, C, C);
  init_conv_shape(e, L.to_k, prefix + ".to_k", {C, kv_in, 1}, C, kv_in);
  init_conv_shape(e, L.to_v, prefix + ".to_v", {C, kv_in, 1}, C, kv_in);
  init_conv_shape(e, L.to_out, prefix + ".to_out.0", {C, C, 1}, C, C);
  if (kv_in == C) {  // self-attention: engine-owned fused q|k|v projection (not part of the state_dict)
    L.qkv_w.reset(new Param());
    L.qkv_b.reset(new Param());
    L.qkv_w->name = prefix + ".to_qkv(fused).weight";
    L.qkv_w->shape = {3 * C, C, 1, 1};
    L.qkv_b->shape = {3 * C};
    L.qkv.w = L.qkv_w.get(); L.qkv.b = L.qkv_b.get();
    L.qkv.Cout = 3 * C; L.qkv.Cin = C; L.qkv.k = 1; L.qkv.stride = 1;
  }
}
// reference: attention_blocks.py:291-335 (Attention dispatcher), :233-288 (SpatialTransformer), :128-195
void init_attention(EngineBase& e, SpatialAttnLayer& A, const std::string& prefix, int kind, int C, int emb_dim) {
  A.kind = kind; A.C = C; A.heads = 8; A.d = C / 8; A.emb_dim = emb_dim;
  if (kind == 1) {
    init_lin_attn(e, A.cros_atn, prefix, C, emb_dim > 0 ? emb_dim : C);
  } else if (kind == 2) {
    init_norm(e, A.norm, prefix + ".norm", C);
    init_conv(e, A.proj_in, prefix + ".proj_in", C, C, 1, 1);
    const std::string tb = prefix + ".transformer_blocks.0";
    init_lin_attn(e, A.self_atn, tb + ".self_atn", C, C);
    if (emb_dim > 0) init_lin_attn(e, A.cros_atn, tb + ".cros_atn", C, emb_dim);
    init_norm(e, A.ln, tb + ".proj_out.0.norm", C);
    init_conv_shape(e, A.ff_in, tb + ".proj_out.0.proj", {8 * C, C}, 8 * C, C);
    init_conv(e, A.ff_out, tb + ".proj_out.2", C, 4 * C, 1, 1);
    init_conv(e, A.proj_out, prefix + ".proj_out", C, C, 1, 1);
  }
}

void init_norm(EngineBase& e, NormLayer& L, const std::string& prefix, int C) {
  L.g = e.add_param(prefix + ".weight", {C});
  L.b = e.add_param(prefix + ".bias", {C});
  L.C = C;
}
// reference: conv_blocks.py:305-364 UnetResBlock = 2 x BasicResBlock (+ local_embedder Linear)
void init_resblock(EngineBase& e, ResBlockLayer& rb, const std::string& prefix, int Cin, int Cout, int k, int emb_dim) {
  rb.Cin = Cin; rb.Cout = Cout;
  init_conv(e, rb.conv1, prefix + ".block_seq.0.basic_block.conv", Cout, Cin, k, 1);
  init_norm(e, rb.norm1, prefix + ".block_seq.0.basic_block.norm", Cout);
  rb.has_res_conv = Cin != Cout;
  if (rb.has_res_conv) init_conv(e, rb.conv_res, prefix + ".block_seq.0.conv_res", Cout, Cin, 1, 1);
  init_conv(e, rb.conv2, prefix + ".block_seq.1.basic_block.conv", Cout, Cout, k, 1);
  init_norm(e, rb.norm2, prefix + ".block_seq.1.basic_block.norm", Cout);
  if (emb_dim > 0) {
    rb.emb_w = e.add_param(prefix + ".local_embedder.1.weight", {Cout, emb_dim});
    rb.emb_b = e.add_param(prefix + ".local_embedder.1.bias", {Cout});
  }
}

// ---- tensors -------------------------------------------------------------------------------------
Tens EngineBase::new_tensor(int N, int H, int W, int C, int layout) {
  Tens t;
  t.N = N; t.H = H; t.W = W; t.C = C; t.layout = layout;
  const size_t esize = layout == kNHWCSplit ? 2 : 4;  // split tensors are two fp16 planes
  const size_t plane_bytes = align_up(static_cast<size_t>(t.elems()) * esize, 1024);
  t.plane = static_cast<long long>(plane_bytes / esize);
  t.bytes = layout == kNHWCSplit ? 2 * plane_bytes : plane_bytes;
  t.off = arena.alloc(t.bytes);
  t.ptr = dry ? nullptr : reinterpret_cast<float*>(base + t.off);
  return t;
}
Tens EngineBase::new_floats(size_t n) {
  Tens t;
  t.N = 1; t.H = 1; t.W = 1; t.C = static_cast<int>(n); t.layout = kNHWCRaw;
  t.bytes = align_up(n * 4, 1024);
  t.off = arena.alloc(t.bytes);
  t.ptr = dry ? nullptr : reinterpret_cast<float*>(base + t.off);
  return t;
}
void EngineBase::free_tensor(const Tens& t) { arena.release(t.off, t.bytes); }

// ---- derived weight layouts ----------------------------------------------------------------------
int EngineBase::ensure_w_tc(ConvLayer& L, int cin_pad) {
  if (L.tc_version == version) return 0;
  const int cp = cin_pad > 0 ? cin_pad : L.Cin;
  // 2 fp16 planes of Cout x k x k x cp elements == that many floats
  if (L.w_tc.alloc(static_cast<size_t>(L.Cout) * cp * L.k * L.k) || L.tc_scales.alloc(4)) return 1;
  int rc = prep_weight_tc(L.w->data.p, reinterpret_cast<__half*>(L.w_tc.p), L.tc_scales.p, L.Cout, L.Cin, L.k, L.k,
                          prep_stream, cp);
  if (rc) return rc;
  L.tc_version = version;
  return 0;
}
int EngineBase::ensure_w_simt(ConvLayer& L) {
  if (L.simt_version == version) return 0;
  if (L.w_simt.alloc(L.w->numel())) return 1;
  int rc = prep_weight_simt(L.w->data.p, L.w_simt.p, L.Cout, L.Cin, L.k, L.k, prep_stream);
  if (rc) return rc;
  L.simt_version = version;
  return 0;
}

// ---- op builders ---------------------------------------------------------------------------------
bool EngineBase::conv_gn_fusable(const ConvLayer& L, const Tens& in0, const Tens* in1, int Ho, int Wo, int groups) const {
  const int C1 = in1 ? in1->C : 0;
  return L.stride == 1 && in0.layout == kNHWCSplit && (!in1 || in1->layout == kNHWCSplit) && g_gn_variant == 3 &&
         conv_tc_supported(in0.N, Ho, Wo, in0.C, C1, L.Cout, L.k, 1) && conv_tc_gn_fusable(Ho, Wo, L.Cout, groups) != 0;
}

int EngineBase::add_conv(ConvLayer& L, const Tens& in0, const Tens* in1, const Tens& out, const Tens* stats,
                         int* chunks, const Tens* res, const float* emb, int emb_stride, const GnFuse* gn) {
  const int C1 = in1 ? in1->C : 0;
  MF_REQUIRE(in0.C + C1 == L.Cin, "conv input channels do not match the weight (" + L.w->name + ")");
  MF_REQUIRE(out.C == L.Cout, "conv output channels do not match the weight (" + L.w->name + ")");
  const bool tc = in0.layout == kNHWCSplit && (!in1 || in1->layout == kNHWCSplit) && out.layout != kNCHW &&
                  (L.stride == 1 || (in0.H % 2 == 0 && in0.W % 2 == 0)) &&
                  conv_tc_supported(in0.N, out.H, out.W, in0.C, C1, L.Cout, L.k, L.stride);
  if (tc) {
    ++n_tc;
    if (chunks) *chunks = conv_tc_stats_chunks(out.H, out.W);
    if (dry) return 0;
    int rc = ensure_w_tc(L);
    if (rc) return rc;
    ConvTcDesc d{};
    d.src0 = in0.hptr(); d.src0_plane = in0.plane; d.C0 = in0.C;
    d.src1 = in1 ? in1->hptr() : nullptr; d.src1_plane = in1 ? in1->plane : 0; d.C1 = C1;
    d.N = in0.N; d.H = in0.H; d.W = in0.W; d.stride = L.stride;
    d.w_planes = reinterpret_cast<const __half*>(L.w_tc.p); d.w_inv_scale = L.tc_scales.p + 1;
    d.Cout = L.Cout; d.ksize = L.k;
    d.bias = L.b->data.p;
    d.out = out.ptr; d.out_plane = out.plane; d.out_mode = out.layout == kNHWCSplit ? kOutSplit : kOutRaw;
    d.stats = stats ? stats->ptr : nullptr;
    if (res) {
      d.res = res->ptr; d.res_plane = res->plane; d.res_kind = res->layout == kNHWCSplit ? 1 : 2;
    }
    d.emb = emb; d.emb_stride = emb_stride;
    if (gn != nullptr) {
      MF_REQUIRE(stats == nullptr && res == nullptr && emb == nullptr && out.layout == kNHWCSplit,
                 "fused GroupNorm conv: split output, no separate statistics");
      d.gn_groups = gn->groups; d.gn_gamma = gn->nl->g->data.p; d.gn_beta = gn->nl->b->data.p; d.gn_eps = 1e-5f;
      if (gn->res) {
        d.gn_res = gn->res->ptr; d.gn_res_plane = gn->res->plane; d.gn_res_kind = gn->res->layout == kNHWCSplit ? 1 : 2;
      }
      d.gn_emb = gn->emb; d.gn_emb_stride = gn->emb_stride;
    }
    d.scratch = scratch();
    MF_REQUIRE(d.scratch != nullptr, "stream-K scratch allocation failed");
    tc_plans.emplace_back(new ConvTcPlan());
    ConvTcPlan* plan = tc_plans.back().get();
    rc = conv_tc_build(d, plan);
    if (rc) return rc;
    push_op([this, plan](cudaStream_t s) { return conv_tc_launch(*plan, s, io_emb_dedup ? 1 : 0, io_emb_index); }, kOpConvTc,
            2.0 * out.N * out.H * out.W * L.Cout * static_cast<double>(L.Cin) * L.k * L.k);
    return 0;
  }
  MF_REQUIRE(gn == nullptr, "fused GroupNorm exists on the tensor-core path only (" + L.w->name + ")");
  // exact fp32 SIMT path (single source only)
  MF_REQUIRE(res == nullptr && emb == nullptr,
             "fused residual / embedding epilogues exist on the tensor-core path only (" + L.w->name + ")");
  MF_REQUIRE(in1 == nullptr || (in1->layout == in0.layout && in0.layout != kNCHW),
             "two-source convolution: both sources must be NHWC tensors of the same kind (" + L.w->name + ")");
  ++n_simt;
  if (chunks) *chunks = 1;
  if (dry) return 0;
  int rc = ensure_w_simt(L);
  if (rc) return rc;
  ConvSimtDesc d{};
  d.in = in0.ptr; d.in_plane = in0.plane; d.in_layout = in0.layout;
  d.in1 = in1 ? in1->ptr : nullptr; d.in1_plane = in1 ? in1->plane : 0; d.C1 = C1;
  d.N = in0.N; d.Cin = in0.C + C1; d.Hin = in0.H; d.Win = in0.W;
  d.w_kc = L.w_simt.p; d.bias = L.b->data.p; d.Cout = L.Cout; d.ksize = L.k; d.stride = L.stride;
  d.out = out.ptr; d.out_plane = out.plane; d.out_layout = out.layout;
  push_op([d](cudaStream_t s) { return conv_simt(d, s); }, kOpConvSimt,
          2.0 * out.N * out.H * out.W * L.Cout * static_cast<double>(L.Cin) * L.k * L.k);
  if (stats) {
    MF_REQUIRE(out.layout == kNHWCRaw, "GroupNorm statistics need a raw NHWC conv output");
    const float* raw = out.ptr;
    float* part = stats->ptr;
    const int N = out.N, HW = out.H * out.W, C = out.C;
    push_op([raw, part, N, HW, C](cudaStream_t s) { return gn_partial_from_raw(raw, part, N, HW, C, s); }, kOpNorm);
  }
  return 0;
}

int EngineBase::add_upconv2x(ConvLayer& L, const Tens& in, Tens* out) {
  MF_REQUIRE(in.C == L.Cin && L.k == 3 && L.stride == 1, "BasicUp conv must be 3x3 stride 1 (" + L.w->name + ")");
  const bool fold = g_fold_upsample && in.layout == kNHWCSplit &&
                    conv_tc_supported(in.N, in.H, in.W, in.C, 0, L.Cout, 3, 1);
  Tens o = new_tensor(in.N, in.H * 2, in.W * 2, L.Cout, kNHWCSplit);
  *out = o;
  if (fold) {
    ++n_tc;
    if (dry) return 0;
    if (L.up_version != version) {
      if (L.w_up.alloc(16 * static_cast<size_t>(L.Cout) * L.Cin) || L.up_scales.alloc(4)) return 1;
      int rc = prep_weight_up_tc(L.w->data.p, reinterpret_cast<__half*>(L.w_up.p), L.up_scales.p, L.Cout, L.Cin,
                                 prep_stream);
      if (rc) return rc;
      L.up_version = version;
    }
    ConvTcDesc d{};
    d.src0 = in.hptr(); d.src0_plane = in.plane; d.C0 = in.C;
    d.N = in.N; d.H = in.H; d.W = in.W; d.stride = 1; d.up2 = 1;
    d.w_planes = reinterpret_cast<const __half*>(L.w_up.p); d.w_inv_scale = L.up_scales.p + 1;
    d.Cout = L.Cout; d.ksize = 3;
    d.bias = L.b->data.p;
    d.out = o.ptr; d.out_plane = o.plane; d.out_mode = kOutSplit;
    d.scratch = scratch();
    MF_REQUIRE(d.scratch != nullptr, "stream-K scratch allocation failed");
    tc_plans.emplace_back(new ConvTcPlan());
    ConvTcPlan* plan = tc_plans.back().get();
    int rc = conv_tc_build(d, plan);
    if (rc) return rc;
    // algorithmic FLOPs are those of the reference formulation (9 taps at the high resolution); the fold issues 4/9
    push_op([plan](cudaStream_t s) { return conv_tc_launch(*plan, s); }, kOpConvTc,
            2.0 * o.N * o.H * o.W * L.Cout * static_cast<double>(L.Cin) * 9);
    return 0;
  }
  Tens up = new_tensor(in.N, in.H * 2, in.W * 2, in.C, in.layout);
  if (!dry) {
    MF_REQUIRE(in.layout == kNHWCSplit, "explicit upsample expects a split tensor");
    const __half* ip = in.hptr(); __half* op = up.hptr();
    const long long ipl = in.plane, opl = up.plane;
    const int N = in.N, H = in.H, W = in.W, C = in.C;
    push_op([ip, ipl, op, opl, N, H, W, C](cudaStream_t st) { return upsample2x_split(ip, ipl, op, opl, N, H, W, C, st); },
            kOpOther);
  }
  int rc = add_conv(L, up, nullptr, o, nullptr, nullptr);
  free_tensor(up);
  return rc;
}

int EngineBase::add_conv_nchw_in(ConvLayer& L, const float* const* src, int N, int Cin, int H, int W, const Tens& out,
                                 const Tens* stats, int* chunks) {
  MF_REQUIRE(Cin == L.Cin && out.C == L.Cout, "stem conv channel mismatch (" + L.w->name + ")");
  // Narrow stems (Cin = 8) still go to the tensor core: the NCHW input is packed into a zero-padded 64-channel split
  // tensor and the weights get zero columns.  7/8 of the issued MMAs multiply zeros, yet at ~400 TFLOP/s that is 5x
  // faster than the exact-fp32 CUDA-core kernel (0.05 ms instead of 0.25 ms per UNet step at B = 64).
  if (g_stem_on_tc && L.stride == 1 && Cin < 64 && out.layout != kNCHW && out.H == H && out.W == W &&
      conv_tc_supported(N, H, W, 64, 0, L.Cout, L.k, 1)) {
    ++n_tc;
    if (chunks) *chunks = conv_tc_stats_chunks(H, W);
    Tens xp = new_tensor(N, H, W, 64, kNHWCSplit);
    if (!dry) {
      __half* xpp = xp.hptr();
      const long long xpl = xp.plane;
      push_op([this, src, xpp, xpl, N, Cin, H, W](cudaStream_t s) {
        return pack_nchw_to_split(*src, xpp, xpl, N, Cin, H, W, s, 64, io_cfg_pair ? N / 2 : 0);
      }, kOpOther);
      int rc = ensure_w_tc(L, 64);
      if (rc) return rc;
      ConvTcDesc d{};
      d.src0 = xp.hptr(); d.src0_plane = xp.plane; d.C0 = 64;
      d.N = N; d.H = H; d.W = W; d.stride = 1;
      d.w_planes = reinterpret_cast<const __half*>(L.w_tc.p); d.w_inv_scale = L.tc_scales.p + 1;
      d.Cout = L.Cout; d.ksize = L.k;
      d.bias = L.b->data.p;
      d.out = out.ptr; d.out_plane = out.plane; d.out_mode = out.layout == kNHWCSplit ? kOutSplit : kOutRaw;
      d.stats = stats ? stats->ptr : nullptr;
      d.scratch = scratch();
      MF_REQUIRE(d.scratch != nullptr, "stream-K scratch allocation failed");
      tc_plans.emplace_back(new ConvTcPlan());
      ConvTcPlan* plan = tc_plans.back().get();
      rc = conv_tc_build(d, plan);
      if (rc) return rc;
      push_op([plan](cudaStream_t s) { return conv_tc_launch(*plan, s); }, kOpConvTc,
              2.0 * N * H * W * L.Cout * static_cast<double>(L.Cin) * L.k * L.k);
    }
    free_tensor(xp);
    return 0;
  }
  ++n_simt;
  if (chunks) *chunks = 1;
  if (dry) return 0;
  int rc = ensure_w_simt(L);
  if (rc) return rc;
  ConvSimtDesc d{};
  d.in = nullptr; d.in_plane = 0; d.in_layout = kNCHW;
  d.N = N; d.Cin = Cin; d.Hin = H; d.Win = W;
  d.w_kc = L.w_simt.p; d.bias = L.b->data.p; d.Cout = L.Cout; d.ksize = L.k; d.stride = L.stride;
  d.out = out.ptr; d.out_plane = out.plane; d.out_layout = out.layout;
  push_op([d, src](cudaStream_t s) {
    ConvSimtDesc dd = d;
    dd.in = *src;
    return conv_simt(dd, s);
  }, kOpConvSimt, 2.0 * out.N * out.H * out.W * L.Cout * static_cast<double>(L.Cin) * L.k * L.k);
  if (stats) {
    MF_REQUIRE(out.layout == kNHWCRaw, "GroupNorm statistics need a raw NHWC conv output");
    const float* raw = out.ptr;
    float* part = stats->ptr;
    const int HW = out.H * out.W, C = out.C;
    push_op([raw, part, N, HW, C](cudaStream_t s) { return gn_partial_from_raw(raw, part, N, HW, C, s); }, kOpNorm);
  }
  return 0;
}

int EngineBase::add_conv_nchw_out(ConvLayer& L, const Tens& in0, float* const* dst, const Tens* in1) {
  const int C1 = in1 ? in1->C : 0;
  MF_REQUIRE(in0.C + C1 == L.Cin, "head conv channel mismatch (" + L.w->name + ")");
  MF_REQUIRE(in1 == nullptr || (in1->layout == in0.layout && in0.layout != kNCHW), "two-source head: NHWC sources of one kind");
  if (in1 == nullptr && L.k == 1 && L.stride == 1 && L.Cout <= 8 && in0.layout == kNHWCSplit && in0.C % 64 == 0) {
    // narrow 1x1 head straight from the reference-layout weights; the scheduler update can ride in its epilogue
    ++n_simt;
    if (dry) return 0;
    HeadDesc hd{};
    hd.in = in0.hptr(); hd.in_plane = in0.plane;
    hd.w = L.w->data.p; hd.bias = L.b->data.p;
    hd.N = in0.N; hd.HW = in0.H * in0.W; hd.C = in0.C; hd.Cout = L.Cout;
    push_op([this, hd, dst](cudaStream_t s) {
      HeadDesc h2 = hd;
      h2.out = *dst;
      h2.out_u8 = io_out_u8;
      if (io_cfg_pair) {          // plan batch = 2N: uncond | cond halves, one guided output per sample pair
        h2.N = hd.N / 2;
        h2.cfg_pair = 1;
        h2.cfg_guidance = io_cfg_guidance;
      }
      return head1x1(h2, io_step_on ? &io_step : nullptr, s);
    }, kOpConvSimt, 2.0 * in0.N * in0.H * in0.W * L.Cout * static_cast<double>(L.Cin));
    return 0;
  }
  ++n_simt;
  if (dry) return 0;
  int rc = ensure_w_simt(L);
  if (rc) return rc;
  ConvSimtDesc d{};
  d.in = in0.ptr; d.in_plane = in0.plane; d.in_layout = in0.layout;
  d.in1 = in1 ? in1->ptr : nullptr; d.in1_plane = in1 ? in1->plane : 0; d.C1 = C1;
  d.N = in0.N; d.Cin = in0.C + C1; d.Hin = in0.H; d.Win = in0.W;
  d.w_kc = L.w_simt.p; d.bias = L.b->data.p; d.Cout = L.Cout; d.ksize = L.k; d.stride = L.stride;
  d.out = nullptr; d.out_plane = 0; d.out_layout = kNCHW;
  push_op([d, dst](cudaStream_t s) {
    if (*dst == nullptr) return 0;      // optional output not requested by this call
    ConvSimtDesc dd = d;
    dd.out = *dst;
    return conv_simt(dd, s);
  }, kOpConvSimt, 2.0 * in0.N * in0.H * in0.W * L.Cout * static_cast<double>(L.Cin) * L.k * L.k);
  return 0;
}

bool gn_needs_generic(int C, int groups) { return groups <= 0 || C % groups != 0 || (C / groups) % 8 != 0; }

bool EngineBase::can_fold_head(const ConvLayer& head, int C, int groups) const {
  const int c8 = C / 8;
  return g_gn_variant == 3 && head.k == 1 && head.stride == 1 && head.Cout <= 8 && head.Cin == C && C % 8 == 0 &&
         !gn_needs_generic(C, groups) && groups <= 128 && c8 <= 32 && (c8 & (c8 - 1)) == 0 && 256 % c8 == 0;
}

int EngineBase::add_gn_apply(const NormLayer& nl, int groups, const Tens& raw, const Tens& stats, int chunks,
                             const Tens* res, const float* emb, int emb_stride, const Tens& out, int act, ConvLayer* head,
                             float* const* head_dst) {
  MF_REQUIRE(head == nullptr || can_fold_head(*head, raw.C, groups), "this head cannot be folded into gn_apply");
  MF_REQUIRE(groups > 0 && raw.C % groups == 0, "GroupNorm: channels must be divisible by the group count (" + nl.g->name + ")");
  Tens mr = new_floats(static_cast<size_t>(raw.N) * groups * 2);
  if (gn_needs_generic(raw.C, groups)) {
    // small widths (e.g. 32 channels in 32 groups, the VQVAE / reference-test defaults): statistics straight from the raw
    // conv output, one channel per thread-iteration in the apply; `stats` is not used
    MF_REQUIRE(raw.layout == kNHWCRaw, "generic GroupNorm expects the raw fp32 conv output (" + nl.g->name + ")");
    if (!dry) {
      const float* rp = raw.ptr; float* mrp = mr.ptr;
      const int N = raw.N, C = raw.C, HW = raw.H * raw.W;
      push_op([rp, mrp, N, HW, C, groups](cudaStream_t s) { return gn_stats_generic(rp, mrp, N, HW, C, groups, 1e-5f, s); },
              kOpNorm);
      GnApplyDesc d{};
      d.raw = raw.ptr; d.mean_rstd = mr.ptr; d.gamma = nl.g->data.p; d.beta = nl.b->data.p; d.raw_plane = 0; d.act = act;
      if (res) {
        d.res = res->ptr; d.res_plane = res->plane; d.res_kind = res->layout == kNHWCSplit ? kResSplit : kResRaw;
      }
      d.emb = emb; d.emb_stride = emb_stride;
      d.out = out.hptr(); d.out_plane = out.plane;
      d.N = N; d.HW = HW; d.C = C; d.G = groups;
      push_op([this, d](cudaStream_t s) {
        GnApplyDesc dd = d;
        if (dd.emb != nullptr && io_emb_dedup) {
          dd.emb_index = io_emb_index;
          if (io_emb_index == nullptr) dd.emb_stride = 0;
        }
        return gn_apply_generic(dd, s);
      }, kOpNorm);
    }
    free_tensor(mr);
    return 0;
  }
  if (!dry) {
    const float* part = stats.ptr;
    float* mrp = mr.ptr;
    const int N = raw.N, C = raw.C, HW = raw.H * raw.W;
    const bool fused = g_gn_variant == 3 && groups <= 128 && raw.N <= 65535 && 256 % (raw.C / 8) == 0;
    if (!fused)
      push_op([part, mrp, N, chunks, C, groups, HW](cudaStream_t s) {
        return gn_finalize(part, mrp, N, chunks, C, groups, HW, 1e-5f, s);
      }, kOpNorm);
    GnApplyDesc d{};
    if (fused) { d.partial = part; d.chunks = chunks; d.eps = 1e-5f; }
    d.raw = raw.ptr; d.mean_rstd = mr.ptr; d.gamma = nl.g->data.p; d.beta = nl.b->data.p;
    d.raw_plane = raw.layout == kNHWCSplit ? raw.plane : 0; d.act = act;
    if (res) {
      d.res = res->ptr; d.res_plane = res->plane;
      d.res_kind = res->layout == kNHWCSplit ? kResSplit : kResRaw;
    } else {
      d.res = nullptr; d.res_kind = kResNone;
    }
    d.emb = emb; d.emb_stride = emb_stride;
    d.out = out.hptr(); d.out_plane = out.plane;
    d.N = raw.N; d.HW = raw.H * raw.W; d.C = raw.C; d.G = groups;
    if (head != nullptr) {
      MF_REQUIRE(fused, "the folded head needs the fused gn_apply variant");
      d.head_w = head->w->data.p; d.head_b = head->b->data.p; d.head_cout = head->Cout;
      d.out = nullptr;
    }
    push_op([this, d, head_dst](cudaStream_t s) {
      GnApplyDesc dd = d;
      if (dd.emb != nullptr && io_emb_dedup) {   // deduplicated embedding rows: one per class (or a single row)
        dd.emb_index = io_emb_index;
        if (io_emb_index == nullptr) dd.emb_stride = 0;
      }
      if (dd.head_cout > 0) {
        dd.head_out = head_dst ? *head_dst : nullptr;
        dd.head_out_u8 = io_out_u8;
      }
      return gn_apply(dd, s);
    }, kOpNorm, head ? 2.0 * raw.N * raw.H * raw.W * head->Cout * static_cast<double>(head->Cin) : 0.0);
  }
  free_tensor(mr);
  return 0;
}

int EngineBase::add_group_norm_split(const NormLayer& nl, int groups, const Tens& x, const Tens& out) {
  MF_REQUIRE(x.layout == kNHWCSplit && x.C % 8 == 0, "group norm input must be a split tensor");
  Tens part = new_floats(static_cast<size_t>(x.N) * (x.C / 8) * 2);
  if (!dry) {
    const float* xp = x.ptr; float* pp = part.ptr;
    const long long plane = x.plane;
    const int N = x.N, HW = x.H * x.W, C = x.C;
    push_op([xp, pp, N, HW, C, plane](cudaStream_t s) { return gn_partial_from_raw(xp, pp, N, HW, C, s, plane); }, kOpNorm);
  }
  int rc = add_gn_apply(nl, groups, x, part, 1, nullptr, nullptr, 0, out, /*act=*/0);
  free_tensor(part);
  return rc;
}

int EngineBase::ensure_qkv(LinAttnLayer& L) {
  if (L.qkv_version == version) return 0;
  const size_t C = L.C;
  MF_REQUIRE(L.qkv_w != nullptr, "fused qkv projection exists for self-attention only");
  if (L.qkv_w->data.alloc(3 * C * C) || L.qkv_b->data.alloc(3 * C)) return 1;
  const ConvLayer* src[3] = {&L.to_q, &L.to_k, &L.to_v};
  for (int i = 0; i < 3; ++i) {
    MF_CUDA_OK(cudaMemcpyAsync(L.qkv_w->data.p + i * C * C, src[i]->w->data.p, C * C * 4, cudaMemcpyDeviceToDevice,
                               prep_stream));
    MF_CUDA_OK(cudaMemcpyAsync(L.qkv_b->data.p + i * C, src[i]->b->data.p, C * 4, cudaMemcpyDeviceToDevice,
                               prep_stream));
  }
  L.qkv_w->is_set = L.qkv_b->is_set = true;
  L.qkv_version = version;
  L.qkv.tc_version = -1;
  return 0;
}

// Attention dispatcher (attention_blocks.py:331-335): 'none' -> identity; 'linear' -> LinearTransformer with the
// embedding as its single key/value token; 'spatial' -> SpatialTransformer.
int EngineBase::add_attention(SpatialAttnLayer& A, int groups, const Tens& x, const Tens* emb, Tens* out) {
  MF_REQUIRE(A.kind == 1 || A.kind == 2, "add_attention on an identity block");
  MF_REQUIRE(x.layout == kNHWCSplit && x.C == A.C, "attention input must be a split tensor with C channels");
  const int B = x.N, H = x.H, W = x.W, C = A.C, HW = H * W;
  // ---- one-token cross attention: softmax over a single key is 1, so the block adds W_o (W_v emb + b_v) + b_o to every
  //      position (attention_blocks.py:160-195 with embedding [B, E, 1]); q, k and the norm do not influence the result
  Tens cb;
  bool have_cb = false;
  if (emb != nullptr && (A.kind == 1 || A.emb_dim > 0)) {
    LinAttnLayer& X = A.cros_atn;
    MF_REQUIRE(X.kv_in % 32 == 0 && C % 32 == 0 && X.kv_in <= 2048 && C <= 2048, "cross-attention widths must be multiples of 32");
    Tens vv = new_floats(static_cast<size_t>(B) * C);
    cb = new_floats(static_cast<size_t>(B) * C);
    have_cb = true;
    if (!dry) {
      LinearDesc l1{};
      l1.in_mode = 0; l1.in = emb->ptr; l1.W = X.to_v.w->data.p; l1.bias = X.to_v.b->data.p;
      l1.out = vv.ptr; l1.post = 0; l1.B = B; l1.J = C; l1.K = X.kv_in;
      push_op([l1](cudaStream_t s) { return linear_small(l1, s); }, kOpOther, 2.0 * B * C * X.kv_in);
      LinearDesc l2{};
      l2.in_mode = 0; l2.in = vv.ptr; l2.W = X.to_out.w->data.p; l2.bias = X.to_out.b->data.p;
      l2.out = cb.ptr; l2.post = 0; l2.B = B; l2.J = C; l2.K = C;
      push_op([l2](cudaStream_t s) { return linear_small(l2, s); }, kOpOther, 2.0 * B * C * C);
    }
    free_tensor(vv);
  }
  if (A.kind == 1) {
    MF_REQUIRE(have_cb, "'linear' attention without an embedding (self-attention form) is not implemented");
    Tens o = new_tensor(B, H, W, C, kNHWCSplit);
    if (!dry) {
      const __half* ip = x.hptr(); const long long ipl = x.plane; const float* bp = cb.ptr;
      __half* op = o.hptr(); const long long opl = o.plane;
      push_op([ip, ipl, bp, C, op, opl, B, HW](cudaStream_t s) {
        return add_channel_bias_split(ip, ipl, bp, C, op, opl, B, HW, C, s);
      }, kOpOther);
    }
    free_tensor(cb);
    *out = o;
    return 0;
  }
  // ---- SpatialTransformer (attention_blocks.py:276-288)
  int rc = 0;
  Tens h0 = new_tensor(B, H, W, C, kNHWCSplit);
  rc = add_group_norm_split(A.norm, groups, x, h0);
  if (rc) return rc;
  Tens h1 = new_tensor(B, H, W, C, kNHWCSplit);
  rc = add_conv(A.proj_in, h0, nullptr, h1, nullptr, nullptr);
  if (rc) return rc;
  free_tensor(h0);
  // self attention (attention_blocks.py:160-195 with embedding None)
  Tens xn = new_tensor(B, H, W, C, kNHWCSplit);
  rc = add_group_norm_split(A.self_atn.norm_x, groups, h1, xn);
  if (rc) return rc;
  if (!dry) {
    rc = ensure_qkv(A.self_atn);
    if (rc) return rc;
  }
  Tens qkv = new_tensor(B, H, W, 3 * C, kNHWCRaw);
  rc = add_conv(A.self_atn.qkv, xn, nullptr, qkv, nullptr, nullptr);
  if (rc) return rc;
  free_tensor(xn);
  Tens ao = new_tensor(B, H, W, C, kNHWCSplit);
  if (!dry) {
    const float* qp = qkv.ptr; __half* op = ao.hptr(); const long long opl = ao.plane;
    const int heads = A.heads, d = A.d;
    push_op([qp, C, op, opl, B, HW, heads, d](cudaStream_t s) {
      return attention_core(qp, qp + C, qp + 2 * C, 3 * C, op, opl, B, HW, heads, d, s);
    }, kOpOther, 4.0 * B * heads * static_cast<double>(HW) * HW * d);
  }
  free_tensor(qkv);
  // h2 = h1 + to_out(attn)  (+ cross-attention contribution, a per-sample channel vector)
  Tens h2 = new_tensor(B, H, W, C, kNHWCSplit);
  rc = add_conv(A.self_atn.to_out, ao, nullptr, h2, nullptr, nullptr, &h1, have_cb && !dry ? cb.ptr : nullptr, C);
  if (rc) return rc;
  free_tensor(ao);
  free_tensor(h1);
  if (have_cb) free_tensor(cb);
  // feed-forward: LayerNorm -> Linear C->8C -> x*gelu(gate) -> conv1x1 4C->C, + residual (attention_blocks.py:17-25,214-231)
  Tens ln = new_tensor(B, H, W, C, kNHWCSplit);
  if (!dry) {
    const __half* ip = h2.hptr(); const long long ipl = h2.plane; __half* op = ln.hptr(); const long long opl = ln.plane;
    const float* g = A.ln.g->data.p; const float* bt = A.ln.b->data.p;
    const long long tokens = static_cast<long long>(B) * HW;
    push_op([ip, ipl, g, bt, op, opl, tokens, C](cudaStream_t s) {
      return layernorm_split(ip, ipl, g, bt, op, opl, tokens, C, 1e-5f, s);
    }, kOpNorm);
  }
  Tens gg = new_tensor(B, H, W, 8 * C, kNHWCRaw);
  rc = add_conv(A.ff_in, ln, nullptr, gg, nullptr, nullptr);
  if (rc) return rc;
  free_tensor(ln);
  Tens ge = new_tensor(B, H, W, 4 * C, kNHWCSplit);
  if (!dry) {
    const float* ip = gg.ptr; __half* op = ge.hptr(); const long long opl = ge.plane;
    const long long tokens = static_cast<long long>(B) * HW;
    const int Ch = 4 * C;
    push_op([ip, op, opl, tokens, Ch](cudaStream_t s) { return geglu_split(ip, op, opl, tokens, Ch, s); }, kOpOther);
  }
  free_tensor(gg);
  Tens h4 = new_tensor(B, H, W, C, kNHWCSplit);
  rc = add_conv(A.ff_out, ge, nullptr, h4, nullptr, nullptr, &h2);
  if (rc) return rc;
  free_tensor(ge);
  free_tensor(h2);
  Tens o = new_tensor(B, H, W, C, kNHWCSplit);
  rc = add_conv(A.proj_out, h4, nullptr, o, nullptr, nullptr, &x);
  if (rc) return rc;
  free_tensor(h4);
  *out = o;
  return 0;
}

// x1 = swish(gn(conv1(x))) + res(x) + emb ;  x2 = swish(gn(conv2(x1))) + x1      (conv_blocks.py:347-364)
int EngineBase::add_resblock(ResBlockLayer& rb, int groups, const Tens& in0, const Tens* in1, const Tens* embT,
                             int emb_stride, Tens* out, ConvLayer* head, float* const* head_dst) {
  const int N = in0.N, H = in0.H, W = in0.W;
  const float* emb = (embT != nullptr && rb.emb_offset >= 0 && !dry) ? embT->ptr + rb.emb_offset : nullptr;
  int rc = 0;
  // ---- levels where a CTA (pair) holds whole samples: GroupNorm + Swish + residual + embedding ride in the conv epilogue,
  //      no raw fp32 tensor, no GroupNorm launch
  Tens x1_like = in0;            // geometry of the first half's output (what conv2 reads)
  x1_like.C = rb.Cout; x1_like.layout = kNHWCSplit;
  if (head == nullptr && conv_gn_fusable(rb.conv1, in0, in1, H, W, groups) &&
      conv_gn_fusable(rb.conv2, x1_like, nullptr, H, W, groups)) {
    Tens res_raw;
    const Tens* res = nullptr;
    if (rb.has_res_conv) {
      res_raw = new_tensor(N, H, W, rb.Cout, kNHWCRaw);
      rc = add_conv(rb.conv_res, in0, in1, res_raw, nullptr, nullptr);
      if (rc) return rc;
      res = &res_raw;
    } else {
      MF_REQUIRE(in1 == nullptr, "identity residual over a concatenated input is not representable");
      res = &in0;
    }
    Tens x1 = new_tensor(N, H, W, rb.Cout, kNHWCSplit);
    GnFuse g1{&rb.norm1, groups, res, emb, emb_stride};
    rc = add_conv(rb.conv1, in0, in1, x1, nullptr, nullptr, nullptr, nullptr, 0, &g1);
    if (rc) return rc;
    if (rb.has_res_conv) free_tensor(res_raw);
    Tens x2 = new_tensor(N, H, W, rb.Cout, kNHWCSplit);
    GnFuse g2{&rb.norm2, groups, &x1, nullptr, 0};
    rc = add_conv(rb.conv2, x1, nullptr, x2, nullptr, nullptr, nullptr, nullptr, 0, &g2);
    if (rc) return rc;
    free_tensor(x1);
    *out = x2;
    return 0;
  }
  const int max_chunks = std::max(1, conv_tc_stats_chunks(H, W));
  Tens raw = new_tensor(N, H, W, rb.Cout, kNHWCRaw);
  Tens part = new_floats(static_cast<size_t>(N) * max_chunks * ((rb.Cout + 7) / 8) * 2);
  int chunks = 1;
  const bool generic_gn = gn_needs_generic(rb.Cout, groups);   // statistics then come from the raw tensor, not the conv
  rc = add_conv(rb.conv1, in0, in1, raw, generic_gn ? nullptr : &part, &chunks);
  if (rc) return rc;
  Tens res_raw;
  const Tens* res = nullptr;
  if (rb.has_res_conv) {
    res_raw = new_tensor(N, H, W, rb.Cout, kNHWCRaw);
    rc = add_conv(rb.conv_res, in0, in1, res_raw, nullptr, nullptr);
    if (rc) return rc;
    res = &res_raw;
  } else {
    MF_REQUIRE(in1 == nullptr, "identity residual over a concatenated input is not representable");
    res = &in0;
  }
  Tens x1 = new_tensor(N, H, W, rb.Cout, kNHWCSplit);
  rc = add_gn_apply(rb.norm1, groups, raw, part, chunks, res, emb, emb_stride, x1);
  if (rc) return rc;
  if (rb.has_res_conv) free_tensor(res_raw);
  // second half reuses `raw` and `part`
  rc = add_conv(rb.conv2, x1, nullptr, raw, generic_gn ? nullptr : &part, &chunks);
  if (rc) return rc;
  Tens x2;
  if (head != nullptr) {
    x2.N = N; x2.H = H; x2.W = W; x2.C = rb.Cout; x2.layout = kNHWCSplit;   // never materialised (bytes == 0)
    rc = add_gn_apply(rb.norm2, groups, raw, part, chunks, &x1, nullptr, 0, x2, 1, head, head_dst);
  } else {
    x2 = new_tensor(N, H, W, rb.Cout, kNHWCSplit);
    rc = add_gn_apply(rb.norm2, groups, raw, part, chunks, &x1, nullptr, 0, x2);
  }
  if (rc) return rc;
  free_tensor(raw);
  free_tensor(part);
  free_tensor(x1);
  *out = x2;
  return 0;
}

int EngineBase::run_profiled(cudaStream_t s, float* ms, int* kinds, double* flops, int max_ops, int* n_ops) {
  const int n = static_cast<int>(ops.size());
  MF_REQUIRE(n <= max_ops, "profile buffers too small");
  std::vector<cudaEvent_t> ev(n + 1);
  for (auto& e : ev) MF_CUDA_OK(cudaEventCreate(&e));
  int rc = 0;
  for (int i = 0; i < n && rc == 0; ++i) {
    MF_CUDA_OK(cudaEventRecord(ev[i], s));
    rc = ops[i](s);
  }
  MF_CUDA_OK(cudaEventRecord(ev[n], s));
  MF_CUDA_OK(cudaStreamSynchronize(s));
  for (int i = 0; i < n; ++i) {
    MF_CUDA_OK(cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]));
    kinds[i] = op_meta[i].kind;
    flops[i] = op_meta[i].flops;
  }
  for (auto& e : ev) cudaEventDestroy(e);
  *n_ops = n;
  return rc;
}

int EngineBase::run(cudaStream_t s) {
  for (auto& op : ops) {
    int rc = op(s);
    if (rc) return rc;
  }
  return 0;
}

}  // namespace mf
