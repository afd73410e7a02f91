#include "mf_engine.cuh"

#include <algorithm>

namespace mf {

int g_fold_upsample = 1;  // BasicUp as four phase convolutions (0: explicit nearest-x2 kernel + conv3x3)

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
  const size_t plane_bytes = align_up(static_cast<size_t>(t.elems()) * 4, 1024);
  t.plane = static_cast<long long>(plane_bytes / 4);
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
int EngineBase::ensure_w_tc(ConvLayer& L) {
  if (L.tc_version == version) return 0;
  if (L.w_tc.alloc(2 * L.w->numel())) return 1;
  int rc = prep_weight_tc(L.w->data.p, L.w_tc.p, L.Cout, L.Cin, L.k, L.k, prep_stream);
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
int EngineBase::add_conv(ConvLayer& L, const Tens& in0, const Tens* in1, const Tens& out, const Tens* stats,
                         int* chunks) {
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
    d.src0 = in0.ptr; d.src0_plane = in0.plane; d.C0 = in0.C;
    d.src1 = in1 ? in1->ptr : nullptr; d.src1_plane = in1 ? in1->plane : 0; d.C1 = C1;
    d.N = in0.N; d.H = in0.H; d.W = in0.W; d.stride = L.stride;
    d.w_planes = L.w_tc.p; d.Cout = L.Cout; d.ksize = L.k;
    d.bias = L.b->data.p;
    d.out = out.ptr; d.out_plane = out.plane; d.out_mode = out.layout == kNHWCSplit ? kOutSplit : kOutRaw;
    d.stats = stats ? stats->ptr : nullptr;
    tc_plans.emplace_back(new ConvTcPlan());
    ConvTcPlan* plan = tc_plans.back().get();
    rc = conv_tc_build(d, plan);
    if (rc) return rc;
    push_op([plan](cudaStream_t s) { return conv_tc_launch(*plan, s); }, kOpConvTc,
            2.0 * out.N * out.H * out.W * L.Cout * static_cast<double>(L.Cin) * L.k * L.k);
    return 0;
  }
  // exact fp32 SIMT path (single source only)
  MF_REQUIRE(in1 == nullptr, "two-source convolution is only available on the tensor-core path (" + L.w->name +
                                 "): channels must be multiples of 32/64 and H*W a power of two >= 32");
  ++n_simt;
  if (chunks) *chunks = 1;
  if (dry) return 0;
  int rc = ensure_w_simt(L);
  if (rc) return rc;
  ConvSimtDesc d{};
  d.in = in0.ptr; d.in_plane = in0.plane; d.in_layout = in0.layout;
  d.N = in0.N; d.Cin = in0.C; d.Hin = in0.H; d.Win = in0.W;
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
      if (L.w_up.alloc(2 * 16 * static_cast<size_t>(L.Cout) * L.Cin)) return 1;
      int rc = prep_weight_up_tc(L.w->data.p, L.w_up.p, L.Cout, L.Cin, prep_stream);
      if (rc) return rc;
      L.up_version = version;
    }
    ConvTcDesc d{};
    d.src0 = in.ptr; d.src0_plane = in.plane; d.C0 = in.C;
    d.N = in.N; d.H = in.H; d.W = in.W; d.stride = 1; d.up2 = 1;
    d.w_planes = L.w_up.p; d.Cout = L.Cout; d.ksize = 3;
    d.bias = L.b->data.p;
    d.out = o.ptr; d.out_plane = o.plane; d.out_mode = kOutSplit;
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
    const float* ip = in.ptr; float* op = up.ptr;
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

int EngineBase::add_conv_nchw_out(ConvLayer& L, const Tens& in0, float* const* dst) {
  MF_REQUIRE(in0.C == L.Cin, "head conv channel mismatch (" + L.w->name + ")");
  ++n_simt;
  if (dry) return 0;
  int rc = ensure_w_simt(L);
  if (rc) return rc;
  ConvSimtDesc d{};
  d.in = in0.ptr; d.in_plane = in0.plane; d.in_layout = in0.layout;
  d.N = in0.N; d.Cin = in0.C; d.Hin = in0.H; d.Win = in0.W;
  d.w_kc = L.w_simt.p; d.bias = L.b->data.p; d.Cout = L.Cout; d.ksize = L.k; d.stride = L.stride;
  d.out = nullptr; d.out_plane = 0; d.out_layout = kNCHW;
  push_op([d, dst](cudaStream_t s) {
    ConvSimtDesc dd = d;
    dd.out = *dst;
    return conv_simt(dd, s);
  }, kOpConvSimt, 2.0 * in0.N * in0.H * in0.W * L.Cout * static_cast<double>(L.Cin) * L.k * L.k);
  return 0;
}

int EngineBase::add_gn_apply(const NormLayer& nl, int groups, const Tens& raw, const Tens& stats, int chunks,
                             const Tens* res, const float* emb, int emb_stride, const Tens& out) {
  MF_REQUIRE(raw.C % groups == 0 && (raw.C / groups) % 8 == 0,
             "GroupNorm: channels per group must be a multiple of 8 (" + nl.g->name + ")");
  Tens mr = new_floats(static_cast<size_t>(raw.N) * groups * 2);
  if (!dry) {
    const float* part = stats.ptr;
    float* mrp = mr.ptr;
    const int N = raw.N, C = raw.C, HW = raw.H * raw.W;
    push_op([part, mrp, N, chunks, C, groups, HW](cudaStream_t s) {
      return gn_finalize(part, mrp, N, chunks, C, groups, HW, 1e-5f, s);
    }, kOpNorm);
    GnApplyDesc d{};
    d.raw = raw.ptr; d.mean_rstd = mr.ptr; d.gamma = nl.g->data.p; d.beta = nl.b->data.p;
    if (res) {
      d.res = res->ptr; d.res_plane = res->plane;
      d.res_kind = res->layout == kNHWCSplit ? kResSplit : kResRaw;
    } else {
      d.res = nullptr; d.res_kind = kResNone;
    }
    d.emb = emb; d.emb_stride = emb_stride;
    d.out = out.ptr; d.out_plane = out.plane;
    d.N = raw.N; d.HW = raw.H * raw.W; d.C = raw.C; d.G = groups;
    push_op([d](cudaStream_t s) { return gn_apply(d, s); }, kOpNorm);
  }
  free_tensor(mr);
  return 0;
}

// x1 = swish(gn(conv1(x))) + res(x) + emb ;  x2 = swish(gn(conv2(x1))) + x1      (conv_blocks.py:347-364)
int EngineBase::add_resblock(ResBlockLayer& rb, int groups, const Tens& in0, const Tens* in1, const Tens* embT,
                             int emb_stride, Tens* out) {
  const int N = in0.N, H = in0.H, W = in0.W;
  const int max_chunks = std::max(1, conv_tc_stats_chunks(H, W));
  Tens raw = new_tensor(N, H, W, rb.Cout, kNHWCRaw);
  Tens part = new_floats(static_cast<size_t>(N) * max_chunks * (rb.Cout / 8) * 2);
  int chunks = 1;
  int rc = add_conv(rb.conv1, in0, in1, raw, &part, &chunks);
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
  const float* emb = (embT != nullptr && rb.emb_offset >= 0 && !dry) ? embT->ptr + rb.emb_offset : nullptr;
  rc = add_gn_apply(rb.norm1, groups, raw, part, chunks, res, emb, emb_stride, x1);
  if (rc) return rc;
  if (rb.has_res_conv) free_tensor(res_raw);
  // second half reuses `raw` and `part`
  rc = add_conv(rb.conv2, x1, nullptr, raw, &part, &chunks);
  if (rc) return rc;
  Tens x2 = new_tensor(N, H, W, rb.Cout, kNHWCSplit);
  rc = add_gn_apply(rb.norm2, groups, raw, part, chunks, &x1, nullptr, 0, x2);
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
