// C ABI (include/medfusion_b200.h): UNet / VAE-decoder engines built on EngineBase, scheduler step,
// and the kernel-level test surface.
#include "../../include/medfusion_b200.h"

#include <algorithm>

#include "mf_engine.cuh"

using namespace mf;

// =================================================================================================
// UNet  (reference: medical_diffusion/models/estimators/unet2.py)
// =================================================================================================
struct mf_unet : public EngineBase {
  mf_unet_config cfg{};
  // layers in reference module order
  ConvLayer in_conv;
  struct EncEntry { bool is_down = false; ResBlockLayer rb; ConvLayer down; SpatialAttnLayer attn; };
  std::vector<std::unique_ptr<EncEntry>> enc;          // in_blocks
  ResBlockLayer mid0, mid2;                            // middle_block.{0,2}
  SpatialAttnLayer mid_attn;                           // middle_block.1
  struct DecEntry { ResBlockLayer rb; SpatialAttnLayer attn; bool has_up = false; ConvLayer up; int up_factor = 1; };
  std::vector<std::unique_ptr<DecEntry>> dec;          // out_blocks (module index order)
  ConvLayer outc;
  std::vector<std::unique_ptr<ConvLayer>> outc_ver;    // deep-supervision heads (unet2.py:214-217), 1x1 on the concat input
  float* io_y_ver[MF_MAX_LEVELS] = {nullptr};          // optional outputs of this call (NULL: head not evaluated)
  const float* io_t_float = nullptr;                   // timesteps as fp32 (NULL: io_t int64)
  Param *t_w1 = nullptr, *t_b1 = nullptr, *t_w2 = nullptr, *t_b2 = nullptr, *cond_table = nullptr;
  DevBuf freqs;
  bool freqs_set = false;
  // fused local-embedder matrix [emb_total][emb_dim] + bias [emb_total]
  int emb_total = 0;
  DevBuf loc_w, loc_b;
  DevBuf cond_ext;        // [num_classes + 1][E]: the label table plus an all-zero "no label" row (CFG as one batch)
  int loc_version = -1;
  std::vector<ResBlockLayer*> all_rb;
  // per-call IO (read by the launch closures)
  const float* io_x = nullptr;
  const long long* io_t = nullptr;
  const long long* io_cond = nullptr;
  float* io_y = nullptr;
  int n_launches = 0;
  int emb_rows = 0;       // rows of the embedding tables actually computed this call (dedup mode)
  bool has_attention = false;

  int init(const mf_unet_config& c);
  int build(int B, int H, int W, char* ws, bool dry_run, cudaStream_t s);
  int ensure_local_embedder(cudaStream_t s);
};

int mf_unet::init(const mf_unet_config& c) {
  cfg = c;
  MF_REQUIRE(c.depth >= 2 && c.depth <= MF_MAX_LEVELS, "depth must be in [2, 8]");
  MF_REQUIRE(c.num_res_blocks >= 1, "num_res_blocks >= 1");
  for (int i = 0; i < c.depth; ++i) {
    MF_REQUIRE(c.attention[i] >= 0 && c.attention[i] <= 2, "attention must be 0 ('none'), 1 ('linear') or 2 ('spatial')");
    has_attention = has_attention || c.attention[i] != 0;
  }
  const int E = c.emb_dim;
  if (E > 0) {
    MF_REQUIRE(c.pos_emb_dim > 0 && c.pos_emb_dim % 2 == 0, "pos_emb_dim must be even");
    t_w1 = add_param("time_embedder.time_emb.1.weight", {E, c.pos_emb_dim});
    t_b1 = add_param("time_embedder.time_emb.1.bias", {E});
    t_w2 = add_param("time_embedder.time_emb.3.weight", {E, E});
    t_b2 = add_param("time_embedder.time_emb.3.bias", {E});
    if (c.num_classes > 0) cond_table = add_param("cond_embedder.embedding.weight", {c.num_classes, E});
  }
  const int* hid = c.hid_chs;
  init_conv(*this, in_conv, "in_conv.conv", hid[0], c.in_ch, c.kernel_sizes[0], c.strides[0]);
  // encoder (unet2.py:70-114)
  for (int i = 1; i < c.depth; ++i) {
    for (int k = 0; k < c.num_res_blocks; ++k) {
      enc.emplace_back(new EncEntry());
      const std::string pre = "in_blocks." + std::to_string(enc.size() - 1) + ".0";
      init_resblock(*this, enc.back()->rb, pre, hid[k == 0 ? i - 1 : i], hid[i], c.kernel_sizes[i], E);
      all_rb.push_back(&enc.back()->rb);
      init_attention(*this, enc.back()->attn, "in_blocks." + std::to_string(enc.size() - 1) + ".1.attention",
                     c.attention[i], hid[i], E);
    }
    if (i < c.depth - 1) {
      enc.emplace_back(new EncEntry());
      enc.back()->is_down = true;
      init_conv(*this, enc.back()->down, "in_blocks." + std::to_string(enc.size() - 1) + ".down_op", hid[i], hid[i],
                c.kernel_sizes[i], c.strides[i]);
    }
  }
  // middle (unet2.py:117-153)
  init_resblock(*this, mid0, "middle_block.0", hid[c.depth - 1], hid[c.depth - 1], c.kernel_sizes[c.depth - 1], E);
  init_attention(*this, mid_attn, "middle_block.1.attention", c.attention[c.depth - 1], hid[c.depth - 1], E);
  init_resblock(*this, mid2, "middle_block.2", hid[c.depth - 1], hid[c.depth - 1], c.kernel_sizes[c.depth - 1], E);
  all_rb.push_back(&mid0);
  all_rb.push_back(&mid2);
  // decoder (unet2.py:158-206)
  for (int i = 1; i < c.depth; ++i) {
    for (int k = 0; k <= c.num_res_blocks; ++k) {
      dec.emplace_back(new DecEntry());
      DecEntry& d = *dec.back();
      const int oc = hid[k == 0 ? i - 1 : i];
      const std::string pre = "out_blocks." + std::to_string(dec.size() - 1);
      init_resblock(*this, d.rb, pre + ".0", hid[i] + oc, oc, c.kernel_sizes[i], E);
      all_rb.push_back(&d.rb);
      init_attention(*this, d.attn, pre + ".1.attention", c.attention[i], oc, E);
      if (i > 1 && k == 0) {
        d.has_up = true;
        d.up_factor = c.strides[i];
        MF_REQUIRE(d.up_factor == 1 || d.up_factor == 2, "BasicUp supports stride 1 or 2");
        init_conv(*this, d.up, pre + ".2.up_op", oc, oc, 3, 1);
      }
    }
  }
  init_conv(*this, outc, "outc.conv.conv", c.out_ch, hid[0], 1, 1);
  // deep-supervision heads: UnetOutBlock(hid[i] + hid[i-1] -> out_ch) for i = 2 .. deep_supervision + 1 (unet2.py:214-217;
  // always the plain out_ch, also with estimate_variance)
  MF_REQUIRE(c.deep_supervision >= 0 && c.deep_supervision <= std::max(0, c.depth - 2), "deep_supervision must be in [0, depth-2]");
  for (int i = 2; i < c.deep_supervision + 2; ++i) {
    outc_ver.emplace_back(new ConvLayer());
    init_conv(*this, *outc_ver.back(), "outc_ver." + std::to_string(i - 2) + ".conv.conv", c.ds_out_ch > 0 ? c.ds_out_ch : c.out_ch,
              hid[i] + hid[i - 1], 1, 1);
  }
  // embedding offsets
  emb_total = 0;
  if (E > 0)
    for (ResBlockLayer* rb : all_rb) {
      rb->emb_offset = emb_total;
      emb_total += rb->Cout;
    }
  return 0;
}

int mf_unet::ensure_local_embedder(cudaStream_t s) {
  if (cfg.emb_dim <= 0 || loc_version == version) return 0;
  const int E = cfg.emb_dim;
  if (loc_w.alloc(static_cast<size_t>(emb_total) * E) || loc_b.alloc(emb_total)) return 1;
  for (ResBlockLayer* rb : all_rb) {
    MF_CUDA_OK(cudaMemcpyAsync(loc_w.p + static_cast<size_t>(rb->emb_offset) * E, rb->emb_w->data.p,
                               static_cast<size_t>(rb->Cout) * E * 4, cudaMemcpyDeviceToDevice, s));
    MF_CUDA_OK(cudaMemcpyAsync(loc_b.p + rb->emb_offset, rb->emb_b->data.p, static_cast<size_t>(rb->Cout) * 4,
                               cudaMemcpyDeviceToDevice, s));
  }
  if (cond_table != nullptr) {
    const size_t n = static_cast<size_t>(cfg.num_classes) * E;
    if (cond_ext.alloc(n + E)) return 1;
    MF_CUDA_OK(cudaMemcpyAsync(cond_ext.p, cond_table->data.p, n * 4, cudaMemcpyDeviceToDevice, s));
    MF_CUDA_OK(cudaMemsetAsync(cond_ext.p + n, 0, static_cast<size_t>(E) * 4, s));
  }
  loc_version = version;
  return 0;
}

// Build the launch plan for UNet.forward (unet2.py:222-269) at a fixed (B, H, W, workspace).
int mf_unet::build(int B, int H, int W, char* ws, bool dry_run, cudaStream_t s) {
  dry = dry_run;
  base = ws;
  prep_stream = s;
  arena = Arena();
  ops.clear();
  op_meta.clear();
  tc_plans.clear();
  n_tc = n_simt = 0;
  const int E = cfg.emb_dim, G = cfg.norm_groups;
  int rc = 0;
  if (!dry) {
    rc = check_all_set();
    if (rc) return rc;
    MF_REQUIRE(E <= 0 || freqs_set, "time frequency table not set (mf_unet_set_time_freqs)");
    rc = ensure_local_embedder(s);
    if (rc) return rc;
  }

  // ---- embeddings (time_embedder.py:67-75, cond_embedders.py:18-23, conv_blocks.py:16-18, :350)
  Tens emb_raw;
  bool have_emb = false;
  Tens embT;
  const Tens* embTp = nullptr;
  if (E > 0) {
    Tens h1 = new_floats(static_cast<size_t>(B) * E);
    Tens emb = new_floats(static_cast<size_t>(B) * E);
    Tens semb = new_floats(static_cast<size_t>(B) * E);
    embT = new_floats(static_cast<size_t>(B) * emb_total);
    embTp = &embT;
    if (!dry) {
      LinearDesc l1{};
      l1.in_mode = 1; l1.freqs = freqs.p; l1.W = t_w1->data.p; l1.bias = t_b1->data.p;
      l1.out = h1.ptr; l1.post = 1; l1.B = B; l1.J = E; l1.K = cfg.pos_emb_dim;
      push_op([this, l1](cudaStream_t st) {
        LinearDesc d = l1;
        d.t = io_t;
        d.t_float = io_t_float;
        d.t_stride = 1;
        if (io_emb_dedup) { d.B = emb_rows; d.t_stride = 0; }
        return linear_small(d, st);
      }, kOpOther, 2.0 * B * E * cfg.pos_emb_dim);
      LinearDesc l2{};
      l2.in_mode = 0; l2.in = h1.ptr; l2.W = t_w2->data.p; l2.bias = t_b2->data.p;
      l2.out = emb.ptr; l2.out2 = semb.ptr; l2.post = 0; l2.B = B; l2.J = E; l2.K = E;
      const float* table = cond_table ? cond_table->data.p : nullptr;
      push_op([this, l2, table](cudaStream_t st) {
        LinearDesc d = l2;
        if (io_cond != nullptr && table != nullptr) {
          d.add_table = io_cfg_pair ? cond_ext.p : table;   // pair mode: index num_classes = the all-zero row
          d.add_idx = io_cond;
        }
        if (io_emb_dedup) {            // row r is class r (identity index); a single row when there is no condition
          d.B = emb_rows;
          d.add_idx = nullptr;
        }
        return linear_small(d, st);
      }, kOpOther, 2.0 * B * E * E);
      LinearDesc l3{};
      l3.in_mode = 0; l3.in = semb.ptr; l3.W = loc_w.p; l3.bias = loc_b.p;
      l3.out = embT.ptr; l3.post = 0; l3.B = B; l3.J = emb_total; l3.K = E;
      push_op([this, l3](cudaStream_t st) {
        LinearDesc d = l3;
        if (io_emb_dedup) d.B = emb_rows;
        return linear_small(d, st);
      }, kOpOther, 2.0 * B * emb_total * E);
    }
    free_tensor(h1);
    free_tensor(semb);
    emb_raw = emb;          // raw time(+label) embedding: the key/value token of the cross-attention blocks
    have_emb = true;
  }

  // ---- encoder
  const int pad0 = cfg.kernel_sizes[0] / 2;
  int h = (H + 2 * pad0 - cfg.kernel_sizes[0]) / cfg.strides[0] + 1;
  int w = (W + 2 * pad0 - cfg.kernel_sizes[0]) / cfg.strides[0] + 1;
  std::vector<Tens> skips;
  Tens x0 = new_tensor(B, h, w, cfg.hid_chs[0], kNHWCSplit);
  rc = add_conv_nchw_in(in_conv, &io_x, B, cfg.in_ch, H, W, x0, nullptr, nullptr);
  if (rc) return rc;
  skips.push_back(x0);
  for (auto& e : enc) {
    const Tens& cur = skips.back();
    Tens nxt;
    if (e->is_down) {
      const int k = e->down.k, st = e->down.stride, pd = k / 2;
      nxt = new_tensor(B, (cur.H + 2 * pd - k) / st + 1, (cur.W + 2 * pd - k) / st + 1, e->down.Cout, kNHWCSplit);
      rc = add_conv(e->down, cur, nullptr, nxt, nullptr, nullptr);
    } else {
      rc = add_resblock(e->rb, G, cur, nullptr, embTp, emb_total, &nxt);
      if (rc == 0 && e->attn.kind != 0) {
        Tens ao;
        rc = add_attention(e->attn, G, nxt, have_emb ? &emb_raw : nullptr, &ao);
        free_tensor(nxt);
        nxt = ao;
      }
    }
    if (rc) return rc;
    skips.push_back(nxt);
  }
  // ---- middle
  Tens hcur, tmp;
  rc = add_resblock(mid0, G, skips.back(), nullptr, embTp, emb_total, &tmp);
  if (rc) return rc;
  if (mid_attn.kind != 0) {
    Tens ao;
    rc = add_attention(mid_attn, G, tmp, have_emb ? &emb_raw : nullptr, &ao);
    if (rc) return rc;
    free_tensor(tmp);
    tmp = ao;
  }
  rc = add_resblock(mid2, G, tmp, nullptr, embTp, emb_total, &hcur);
  if (rc) return rc;
  free_tensor(tmp);
  // ---- decoder: h = cat([h, skip]) -> out_blocks[i-1]  (unet2.py:258-264)
  for (int i = static_cast<int>(dec.size()); i >= 1; --i) {
    DecEntry& d = *dec[i - 1];
    Tens skip = skips.back();
    skips.pop_back();
    // deep supervision (unet2.py:262): the head of level L = (i-1)/(num_res_blocks+1) + 1 reads the concat input of that
    // level's first decoder block
    {
      const int per = cfg.num_res_blocks + 1;
      const int level = (i - 1) / per + 1, kk = (i - 1) % per;
      if (kk == 0 && level >= 2 && level - 2 < static_cast<int>(outc_ver.size())) {
        rc = add_conv_nchw_out(*outc_ver[level - 2], hcur, &io_y_ver[level - 2], &skip);
        if (rc) return rc;
      }
    }
    Tens o;
    rc = add_resblock(d.rb, G, hcur, &skip, embTp, emb_total, &o);
    if (rc) return rc;
    free_tensor(hcur);
    free_tensor(skip);
    if (d.attn.kind != 0) {
      Tens ao;
      rc = add_attention(d.attn, G, o, have_emb ? &emb_raw : nullptr, &ao);
      if (rc) return rc;
      free_tensor(o);
      o = ao;
    }
    if (d.has_up) {
      // BasicUp: nearest x2 then conv3x3 (conv_blocks.py:121-131)
      Tens u;
      if (d.up_factor == 2) {
        rc = add_upconv2x(d.up, o, &u);
      } else {
        u = new_tensor(B, o.H, o.W, d.up.Cout, kNHWCSplit);
        rc = add_conv(d.up, o, nullptr, u, nullptr, nullptr);
      }
      if (rc) return rc;
      free_tensor(o);
      o = u;
    }
    hcur = o;
  }
  // ---- head (unet2.py:267)
  rc = add_conv_nchw_out(outc, hcur, &io_y);
  if (rc) return rc;
  free_tensor(hcur);
  if (E > 0) free_tensor(embT);
  if (have_emb) free_tensor(emb_raw);
  n_launches = static_cast<int>(ops.size());
  return 0;
}

// =================================================================================================
// VAE decoder (reference: latent_embedders.py:718-743, :764-769; conv_blocks.py:444-528 UpBlock)
// =================================================================================================
// Encoder half (latent_embedders.py:680-716 ctor, :756-762 encode): its own launch plan / workspace state, while the
// parameters live in the parent mf_vae's registry (one state_dict, reference key order: inc, encoders, out_enc, ...).
struct mf_vae_enc_plan : public EngineBase {
  const mf_vae_config* cfg = nullptr;
  int in_channels = 3;
  ResBlockLayer inc;
  struct Down { ConvLayer down; ResBlockLayer rb; };
  std::vector<std::unique_ptr<Down>> encoders;  // index i == encoders.{i}
  ConvLayer out0, out1;
  const float* io_x = nullptr;      // [B, in_channels, H, W]
  const float* io_noise = nullptr;  // [B, emb, h, w] standard normal draw, or NULL (z = mean)
  float* io_z = nullptr;            // [B, emb, h, w]
  float* io_moments = nullptr;      // optional [B, 2*emb, h, w] (mean | logvar, unclamped)
  float* moments_ws = nullptr;
  int n_launches = 0;
  int build(int B, int H, int W, char* ws, bool dry_run, cudaStream_t s);
};

struct mf_vae : public EngineBase {
  mf_vae_config cfg{};
  mf_vae_enc_plan enc;
  bool has_encoder = false;
  ResBlockLayer inc_dec;
  struct Up { ConvLayer up; ResBlockLayer rb; int factor = 2; };
  std::vector<std::unique_ptr<Up>> decoders;  // index i == decoders.{i}
  ConvLayer outc;
  Param* codebook = nullptr;          // VQVAE: quantizer.embedder.weight [num_embeddings][emb_channels]
  const float* io_z = nullptr;
  const float* zq_ptr = nullptr;      // VQVAE: the quantised latent inside the workspace (what the stems read)
  float* io_x = nullptr;
  int n_launches = 0;

  int init(const mf_vae_config& c);
  int build(int B, int H, int W, char* ws, bool dry_run, cudaStream_t s);
};

int mf_vae::init(const mf_vae_config& c) {
  cfg = c;
  MF_REQUIRE(c.depth >= 1 && c.depth <= MF_MAX_LEVELS, "depth must be in [1, 8]");
  MF_REQUIRE(c.num_embeddings == 0 || c.in_channels == 0, "VQVAE handles are decoder-only (in_channels must be 0)");
  if (c.in_channels > 0) {
    // encoder parameters, registered in the PARENT registry with the reference's names
    has_encoder = true;
    enc.cfg = &cfg;
    enc.in_channels = c.in_channels;
    MF_REQUIRE(c.strides[0] == 1, "VAE encoder: strides[0] must be 1");
    init_resblock(*this, enc.inc, "inc", c.in_channels, c.hid_chs[0], 3, 0);
    for (int i = 1; i < c.depth; ++i) {
      enc.encoders.emplace_back(new mf_vae_enc_plan::Down());
      auto& d = *enc.encoders.back();
      const std::string pre = "encoders." + std::to_string(i - 1);
      init_conv(*this, d.down, pre + ".down_op.down_op", c.hid_chs[i], c.hid_chs[i - 1], 3, c.strides[i]);
      init_resblock(*this, d.rb, pre + ".conv_block", c.hid_chs[i], c.hid_chs[i], 3, 0);
    }
    init_conv(*this, enc.out0, "out_enc.0.conv", 2 * c.emb_channels, c.hid_chs[c.depth - 1], 3, 1);
    init_conv(*this, enc.out1, "out_enc.1.conv", 2 * c.emb_channels, 2 * c.emb_channels, 1, 1);
  }
  if (c.num_embeddings > 0) {
    // VQVAE (latent_embedders.py:180-320): decode() quantises z first (:315); registered before inc_dec like the reference
    MF_REQUIRE(c.emb_channels <= 16, "VQVAE: emb_channels must be <= 16");
    codebook = add_param("quantizer.embedder.weight", {c.num_embeddings, c.emb_channels});
  }
  init_resblock(*this, inc_dec, "inc_dec", c.emb_channels, c.hid_chs[c.depth - 1], 3, 0);
  for (int i = 0; i < c.depth - 1; ++i) {
    decoders.emplace_back(new Up());
    Up& u = *decoders.back();
    u.factor = c.strides[i + 1];
    MF_REQUIRE(u.factor == 2, "VAE decoder levels must upsample by 2");
    const std::string pre = "decoders." + std::to_string(i);
    init_conv(*this, u.up, pre + ".up_op.up_op", c.hid_chs[i], c.hid_chs[i + 1], 3, 1);
    init_resblock(*this, u.rb, pre + ".conv_block", c.hid_chs[i], c.hid_chs[i], 3, 0);
  }
  init_conv(*this, outc, "outc.conv", c.out_channels, c.hid_chs[0], 1, 1);
  return 0;
}

int mf_vae::build(int B, int H, int W, char* ws, bool dry_run, cudaStream_t s) {
  dry = dry_run;
  base = ws;
  prep_stream = s;
  arena = Arena();
  ops.clear();
  op_meta.clear();
  tc_plans.clear();
  n_tc = n_simt = 0;
  int rc = 0;
  if (!dry) {
    rc = check_all_set();
    if (rc) return rc;
  }
  const int G = cfg.norm_groups;
  const int Ctop = cfg.hid_chs[cfg.depth - 1];
  // ---- VQVAE: nearest codebook row per latent vector (VectorQuantizer.forward, latent_embedders.py:50-69)
  const float* const* zsrc = &io_z;
  Tens zq;
  if (codebook != nullptr) {
    zq = new_floats(static_cast<size_t>(B) * cfg.emb_channels * H * W);
    zq_ptr = zq.ptr;
    zsrc = &zq_ptr;
    if (!dry) {
      float* zqp = zq.ptr;
      const float* cb = codebook->data.p;
      const int C = cfg.emb_channels, K = cfg.num_embeddings;
      push_op([this, zqp, cb, B, C, H, W, K](cudaStream_t st) { return vq_quantize(io_z, cb, zqp, nullptr, B, C, H * W, K, st); },
              kOpOther);
    }
  }
  // ---- inc_dec: UnetResBlock(emb_channels -> Ctop), stem convs read the NCHW latent directly
  Tens raw = new_tensor(B, H, W, Ctop, kNHWCRaw);
  Tens part = new_floats(static_cast<size_t>(B) * std::max(1, conv_tc_stats_chunks(H, W)) * ((Ctop + 7) / 8) * 2);
  int chunks = 1;
  const bool generic_gn = gn_needs_generic(Ctop, G);
  rc = add_conv_nchw_in(inc_dec.conv1, zsrc, B, cfg.emb_channels, H, W, raw, generic_gn ? nullptr : &part, &chunks);
  if (rc) return rc;
  Tens res_raw;
  const Tens* res = nullptr;
  MF_REQUIRE(inc_dec.has_res_conv, "inc_dec with emb_channels == hid_chs[-1] is not supported");
  res_raw = new_tensor(B, H, W, Ctop, kNHWCRaw);
  rc = add_conv_nchw_in(inc_dec.conv_res, zsrc, B, cfg.emb_channels, H, W, res_raw, nullptr, nullptr);
  if (rc) return rc;
  res = &res_raw;
  Tens x1 = new_tensor(B, H, W, Ctop, kNHWCSplit);
  rc = add_gn_apply(inc_dec.norm1, G, raw, part, chunks, res, nullptr, 0, x1);
  if (rc) return rc;
  free_tensor(res_raw);
  if (codebook != nullptr) free_tensor(zq);
  rc = add_conv(inc_dec.conv2, x1, nullptr, raw, generic_gn ? nullptr : &part, &chunks);
  if (rc) return rc;
  Tens hcur = new_tensor(B, H, W, Ctop, kNHWCSplit);
  rc = add_gn_apply(inc_dec.norm2, G, raw, part, chunks, &x1, nullptr, 0, hcur);
  if (rc) return rc;
  free_tensor(raw);
  free_tensor(part);
  free_tensor(x1);
  // ---- decoders[depth-2 .. 0]: nearest x2 + conv3x3, then UnetResBlock
  bool head_folded = false;
  for (int i = cfg.depth - 2; i >= 0; --i) {
    Up& u = *decoders[i];
    Tens uo;
    rc = add_upconv2x(u.up, hcur, &uo);
    if (rc) return rc;
    free_tensor(hcur);
    Tens o;
    // last level: the image head (latent_embedders.py:743) rides in the block's final GroupNorm-apply
    const bool fold = (i == 0) && g_fold_head && can_fold_head(outc, u.rb.Cout, G);
    rc = add_resblock(u.rb, G, uo, nullptr, nullptr, 0, &o, fold ? &outc : nullptr, fold ? &io_x : nullptr);
    if (rc) return rc;
    free_tensor(uo);
    hcur = o;
    head_folded = fold;
  }
  if (!head_folded) {
    rc = add_conv_nchw_out(outc, hcur, &io_x);
    if (rc) return rc;
    free_tensor(hcur);
  } else {
    ++n_simt;   // census: the head still counts as one (CUDA-core) convolution, now without a launch of its own
  }
  n_launches = static_cast<int>(ops.size());
  return 0;
}

// ---- VAE encoder plan: inc -> encoders (conv3x3 s2 + res block) -> out_enc (3x3, 1x1) -> reparameterisation
int mf_vae_enc_plan::build(int B, int H, int W, char* ws, bool dry_run, cudaStream_t s) {
  dry = dry_run;
  base = ws;
  prep_stream = s;
  arena = Arena();
  ops.clear();
  op_meta.clear();
  tc_plans.clear();
  n_tc = n_simt = 0;
  const mf_vae_config& c = *cfg;
  const int G = c.norm_groups;
  const int C0 = c.hid_chs[0];
  int rc = 0;
  // ---- inc: UnetResBlock(in_channels -> C0); both stem convs read the NCHW image directly
  Tens raw = new_tensor(B, H, W, C0, kNHWCRaw);
  Tens part = new_floats(static_cast<size_t>(B) * std::max(1, conv_tc_stats_chunks(H, W)) * (C0 / 8) * 2);
  int chunks = 1;
  rc = add_conv_nchw_in(inc.conv1, &io_x, B, in_channels, H, W, raw, &part, &chunks);
  if (rc) return rc;
  MF_REQUIRE(inc.has_res_conv, "VAE encoder with in_channels == hid_chs[0] is not supported");
  Tens res_raw = new_tensor(B, H, W, C0, kNHWCRaw);
  rc = add_conv_nchw_in(inc.conv_res, &io_x, B, in_channels, H, W, res_raw, nullptr, nullptr);
  if (rc) return rc;
  Tens x1 = new_tensor(B, H, W, C0, kNHWCSplit);
  rc = add_gn_apply(inc.norm1, G, raw, part, chunks, &res_raw, nullptr, 0, x1);
  if (rc) return rc;
  free_tensor(res_raw);
  rc = add_conv(inc.conv2, x1, nullptr, raw, &part, &chunks);
  if (rc) return rc;
  Tens hcur = new_tensor(B, H, W, C0, kNHWCSplit);
  rc = add_gn_apply(inc.norm2, G, raw, part, chunks, &x1, nullptr, 0, hcur);
  if (rc) return rc;
  free_tensor(raw);
  free_tensor(part);
  free_tensor(x1);
  // ---- encoders[i]: BasicDown (conv3x3, stride s) then UnetResBlock   (conv_blocks.py:430-441)
  for (auto& e : encoders) {
    const int st = e->down.stride, pd = e->down.k / 2;
    Tens d = new_tensor(B, (hcur.H + 2 * pd - e->down.k) / st + 1, (hcur.W + 2 * pd - e->down.k) / st + 1, e->down.Cout,
                        kNHWCSplit);
    rc = add_conv(e->down, hcur, nullptr, d, nullptr, nullptr);
    if (rc) return rc;
    free_tensor(hcur);
    Tens o;
    rc = add_resblock(e->rb, G, d, nullptr, nullptr, 0, &o);
    if (rc) return rc;
    free_tensor(d);
    hcur = o;
  }
  // ---- out_enc: conv3x3 -> conv1x1 (no norm / activation), moments in NCHW
  const int E2 = 2 * c.emb_channels;
  Tens m0 = new_tensor(B, hcur.H, hcur.W, E2, kNHWCRaw);
  rc = add_conv(out0, hcur, nullptr, m0, nullptr, nullptr);
  if (rc) return rc;
  free_tensor(hcur);
  Tens mom = new_floats(static_cast<size_t>(B) * E2 * m0.H * m0.W);
  moments_ws = mom.ptr;
  rc = add_conv_nchw_out(out1, m0, &moments_ws);
  if (rc) return rc;
  free_tensor(m0);
  if (!dry) {
    const int ehw = c.emb_channels * m0.H * m0.W;
    push_op([this, B, ehw](cudaStream_t st) { return vae_reparam(moments_ws, io_noise, io_z, io_moments, B, ehw, st); },
            kOpOther);
  }
  free_tensor(mom);
  n_launches = static_cast<int>(ops.size());
  return 0;
}

// =================================================================================================
// C ABI
// =================================================================================================
template <class E>
static int prepare_plan(E* h, int B, int H, int W, void* ws, size_t ws_bytes, cudaStream_t s) {
  if (int rcd = h->bind_device()) return rcd;
  if (h->key.B == B && h->key.H == H && h->key.W == W && h->key.base == ws && h->key.version == h->version) return 0;
  // dry run for the size check
  int rc = h->build(B, H, W, nullptr, true, s);
  if (rc) return rc;
  const size_t need = h->arena.peak;
  if (ws_bytes < need) {
    set_error("workspace too small: need " + std::to_string(need) + " bytes, got " + std::to_string(ws_bytes));
    return 2;
  }
  if ((reinterpret_cast<uintptr_t>(ws) & 1023) != 0) {
    set_error("workspace must be 1024-byte aligned");
    return 2;
  }
  rc = h->build(B, H, W, static_cast<char*>(ws), false, s);
  if (rc) {
    h->key = typename E::PlanKey();
    return rc;
  }
  h->key.B = B; h->key.H = H; h->key.W = W; h->key.base = ws; h->key.version = h->version;
  return 0;
}

extern "C" {

const char* mf_last_error(void) { return mf::get_error(); }
int mf_abi_version(void) { return 3; }
int mf_saturation_count(unsigned long long* out_count, int reset, mf_stream_t stream) {
  MF_REQUIRE(out_count != nullptr, "null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  unsigned long long total = 0;
  int rc = sat_read_kernels(&total, reset, s);
  if (rc == 0) rc = sat_read_conv_tc(&total, reset, s);
  if (rc == 0) rc = sat_read_attn(&total, reset, s);
  *out_count = total;
  return rc;
}
int mf_set_debias_eps(float eps_per_kblock) {
  mf::g_debias_eps_per_kblock = eps_per_kblock;
  return 0;
}
int mf_set_pdl(int enable) {
  mf::g_pdl = enable ? 1 : 0;
  return 0;
}
int mf_set_attn_tc(int enable) {
  mf::g_attn_tc = enable ? 1 : 0;
  return 0;
}
int mf_set_fuse_gn(int enable) {
  mf::g_fuse_gn = enable ? 1 : 0;
  return 0;
}
int mf_set_fold_head(int enable) {
  mf::g_fold_head = enable ? 1 : 0;
  return 0;
}
int mf_set_gn_variant(int v) {
  mf::g_gn_variant = v;
  return 0;
}
int mf_set_stem_on_tc(int enable) {
  mf::g_stem_on_tc = enable ? 1 : 0;
  return 0;
}
int mf_set_fold_upsample(int enable) {
  mf::g_fold_upsample = enable ? 1 : 0;
  return 0;
}
int mf_set_stream_k(int enable) {
  mf::g_stream_k = enable ? 1 : 0;
  return 0;
}
int mf_set_split_fill(int min_k_blocks) {
  MF_REQUIRE(min_k_blocks >= 0, "min_k_blocks must be >= 0 (0 = off)");
  mf::g_split_fill = min_k_blocks;
  return 0;
}
int mf_op_conv_tc_plan(int N, int H, int W, int C0, int C1, int Cout, int ksize, int stride, int up2, int sm_count,
                       int* out8) {
  MF_REQUIRE(out8 != nullptr, "null argument");
  return mf::conv_tc_plan_query(N, H, W, C0, C1, Cout, ksize, stride, up2, sm_count, out8);
}
int mf_set_row_patch(int enable) {
  mf::g_row_patch = enable ? 1 : 0;
  return 0;
}
int mf_set_block_n(int block_n) {
  MF_REQUIRE(block_n == 0 || block_n == 64 || block_n == 128 || block_n == 256, "block_n must be 0, 64, 128 or 256");
  mf::g_default_block_n = block_n;
  return 0;
}
int mf_set_cta_group(int cta_group) {
  MF_REQUIRE(cta_group >= 0 && cta_group <= 2, "cta_group must be 0 (auto), 1 or 2");
  mf::g_default_cta_group = cta_group;
  return 0;
}
int mf_set_drain_interval(int k_blocks) {
  MF_REQUIRE(k_blocks >= 1, "drain interval must be >= 1");
  mf::g_default_drain_interval = k_blocks;
  return 0;
}

// ---- UNet ---------------------------------------------------------------------------------------
int mf_unet_create(const mf_unet_config* cfg, mf_unet** out) {
  if (!cfg || !out) { set_error("null argument"); return 2; }
  std::unique_ptr<mf_unet> h(new mf_unet());
  int rc = h->init(*cfg);
  if (rc) return rc;
  *out = h.release();
  return 0;
}
void mf_unet_destroy(mf_unet* h) { delete h; }
int mf_unet_param_count(const mf_unet* h) { return static_cast<int>(h->params.size()); }
const char* mf_unet_param_name(const mf_unet* h, int i) {
  return (i >= 0 && i < static_cast<int>(h->params.size())) ? h->params[i]->name.c_str() : nullptr;
}
int mf_unet_param_shape(const mf_unet* h, int i, int64_t shape[4], int* ndim) {
  if (i < 0 || i >= static_cast<int>(h->params.size())) { set_error("param index out of range"); return 2; }
  const auto& s = h->params[i]->shape;
  *ndim = static_cast<int>(s.size());
  for (size_t k = 0; k < s.size(); ++k) shape[k] = s[k];
  return 0;
}
int mf_unet_set_param(mf_unet* h, const char* name, const float* d_data, const int64_t* shape, int ndim,
                      mf_stream_t stream) {
  return h->set_param(name, d_data, shape, ndim, static_cast<cudaStream_t>(stream));
}
int mf_unet_set_time_freqs(mf_unet* h, const float* d_freqs, int n, mf_stream_t stream) {
  MF_REQUIRE(h->cfg.emb_dim > 0 && n == h->cfg.pos_emb_dim / 2, "frequency table must have pos_emb_dim/2 entries");
  if (h->freqs.alloc(n)) return 1;
  MF_CUDA_OK(cudaMemcpyAsync(h->freqs.p, d_freqs, n * sizeof(float), cudaMemcpyDeviceToDevice,
                             static_cast<cudaStream_t>(stream)));
  h->freqs_set = true;
  ++h->version;
  return 0;
}
size_t mf_unet_workspace_bytes(mf_unet* h, int B, int H, int W) {
  const auto saved = h->key;
  if (h->build(B, H, W, nullptr, true, nullptr)) return 0;
  const size_t need = h->arena.peak;
  h->key = typename mf_unet::PlanKey();  // the dry build clobbered the op list
  (void)saved;
  return need;
}
int mf_unet_forward_ex(mf_unet* h, const float* d_x_t, const int64_t* d_t, const float* d_t_float, const int64_t* d_cond,
                       float* d_y, float* const* d_y_ver, int n_ver, int B, int H, int W, void* d_workspace,
                       size_t workspace_bytes, mf_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  MF_REQUIRE(d_x_t && d_y && B > 0 && H > 0 && W > 0, "bad arguments");
  MF_REQUIRE(h->cfg.emb_dim <= 0 || d_t != nullptr || d_t_float != nullptr, "t is required when the UNet has a time embedder");
  MF_REQUIRE(n_ver >= 0 && n_ver <= static_cast<int>(h->outc_ver.size()), "more deep-supervision outputs than heads");
  int rc = prepare_plan(h, B, H, W, d_workspace, workspace_bytes, s);
  if (rc) return rc;
  h->io_x = d_x_t;
  h->io_t = reinterpret_cast<const long long*>(d_t);
  h->io_t_float = d_t_float;
  h->io_cond = reinterpret_cast<const long long*>(d_cond);
  h->io_y = d_y;
  for (int k = 0; k < n_ver; ++k) h->io_y_ver[k] = d_y_ver ? d_y_ver[k] : nullptr;
  rc = h->run(s);
  h->io_t_float = nullptr;
  for (int k = 0; k < MF_MAX_LEVELS; ++k) h->io_y_ver[k] = nullptr;
  return rc;
}
int mf_unet_forward(mf_unet* h, const float* d_x_t, const int64_t* d_t, const int64_t* d_cond, float* d_y, int B,
                    int H, int W, void* d_workspace, size_t workspace_bytes, mf_stream_t stream) {
  return mf_unet_forward_ex(h, d_x_t, d_t, nullptr, d_cond, d_y, nullptr, 0, B, H, W, d_workspace, workspace_bytes, stream);
}
// All samples share the timestep (the sampling loop calls the estimator with t.expand(B)): the embedding MLP then
// depends only on the class, so it is evaluated once per class (or once) instead of once per sample.
static void set_emb_dedup(mf_unet* h, int uniform_t, const int64_t* d_cond, int B) {
  h->io_emb_dedup = false;
  h->io_emb_index = nullptr;
  if (!uniform_t || h->cfg.emb_dim <= 0 || h->has_attention) return;
  const bool cond = d_cond != nullptr && h->cfg.num_classes > 0;
  const int rows = cond ? h->cfg.num_classes : 1;
  if (rows > B) return;   // the plan's embedding buffers hold B rows
  h->io_emb_dedup = true;
  h->emb_rows = rows;
  h->io_emb_index = cond ? reinterpret_cast<const long long*>(d_cond) : nullptr;
}
static void fill_step_desc(SchedStepDesc& d, const mf_step_args* a, const float* x_t, const int64_t* t, int B, int chw) {
  d = SchedStepDesc{};
  d.x_t = x_t; d.pred = nullptr; d.pred_uncond = a->d_pred_uncond; d.guidance = a->guidance_scale;
  d.t = reinterpret_cast<const long long*>(t);
  d.noise = a->d_noise;
  d.t_next = reinterpret_cast<const long long*>(a->d_t_next);
  d.noise2 = a->d_noise_ddim;
  d.objective_x0 = a->objective_is_x0; d.clip_x0 = a->clip_x0;
  d.x_prior = a->d_x_prior; d.x_0 = a->d_x_0; d.x_T = a->d_x_T; d.x_next = a->d_x_next;
  d.B = B; d.CHW = chw;
  d.tab.sqrt_recip_ac = a->tables->sqrt_recip_alphas_cumprod;
  d.tab.sqrt_recipm1_ac = a->tables->sqrt_recipm1_alphas_cumprod;
  d.tab.coef1 = a->tables->posterior_mean_coef1;
  d.tab.coef2 = a->tables->posterior_mean_coef2;
  d.tab.post_var = a->tables->posterior_variance;
  d.tab.betas = a->tables->betas;
  d.tab.alphas_cumprod = a->tables->alphas_cumprod;
}
int mf_unet_forward_step(mf_unet* h, const float* d_x_t, const int64_t* d_t, const int64_t* d_cond, float* d_y, int B,
                         int H, int W, void* d_workspace, size_t workspace_bytes, const mf_step_args* step,
                         mf_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  MF_REQUIRE(d_x_t && d_t && step && step->tables && B > 0 && H > 0 && W > 0, "bad arguments");
  MF_REQUIRE(h->cfg.out_ch <= 8 && h->cfg.hid_chs[0] % 64 == 0 && h->cfg.kernel_sizes[0] / 2 * 2 + 1 == h->cfg.kernel_sizes[0] &&
                 h->cfg.strides[0] == 1,
             "the fused head needs out_ch <= 8, hid_chs[0] % 64 == 0 and a stride-1 stem (use mf_unet_forward + mf_sched_step)");
  int rc = prepare_plan(h, B, H, W, d_workspace, workspace_bytes, s);
  if (rc) return rc;
  h->io_x = d_x_t;
  h->io_t = reinterpret_cast<const long long*>(d_t);
  h->io_cond = reinterpret_cast<const long long*>(d_cond);
  h->io_y = d_y;
  set_emb_dedup(h, step->uniform_t, d_cond, B);
  fill_step_desc(h->io_step, step, d_x_t, d_t, B, h->cfg.out_ch * H * W);
  h->io_step_on = true;
  rc = h->run(s);
  h->io_step_on = false;
  h->io_emb_dedup = false;
  return rc;
}
// Classifier-free guidance as one 2B batch.  d_cond2[2B]: labels of the unconditional half first (num_classes = "no
// label"), then the conditional half.  The scheduler update consumes pred_u + g (pred_c - pred_u) inside the head.
int mf_unet_forward_step_cfg(mf_unet* h, const float* d_x_t, const int64_t* d_t, const int64_t* d_cond2, int B, int H,
                             int W, void* d_workspace, size_t workspace_bytes, const mf_step_args* step,
                             mf_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  MF_REQUIRE(d_x_t && d_t && d_cond2 && step && step->tables && B > 0 && H > 0 && W > 0, "bad arguments");
  MF_REQUIRE(step->uniform_t && step->d_pred_uncond == nullptr, "the one-batch CFG step needs uniform t and no separate uncond prediction");
  MF_REQUIRE(h->cfg.num_classes > 0 && h->cfg.emb_dim > 0 && !h->has_attention, "the one-batch CFG step needs a label embedder (and no attention blocks)");
  MF_REQUIRE(h->cfg.out_ch <= 8 && h->cfg.hid_chs[0] % 64 == 0 && h->cfg.kernel_sizes[0] % 2 == 1 && h->cfg.strides[0] == 1 &&
                 h->cfg.in_ch < 64 && mf::g_stem_on_tc,
             "the one-batch CFG step needs the narrow fused head and the tensor-core stem");
  int rc = prepare_plan(h, 2 * B, H, W, d_workspace, workspace_bytes, s);
  if (rc) return rc;
  h->io_x = d_x_t;
  h->io_t = reinterpret_cast<const long long*>(d_t);
  h->io_cond = reinterpret_cast<const long long*>(d_cond2);
  h->io_y = nullptr;
  // embeddings: one row per class + the "no label" row, indexed by d_cond2
  h->io_emb_dedup = true;
  h->emb_rows = h->cfg.num_classes + 1;
  h->io_emb_index = reinterpret_cast<const long long*>(d_cond2);
  MF_REQUIRE(h->emb_rows <= 2 * B, "more classes than samples");
  h->io_cfg_pair = true;
  h->io_cfg_guidance = step->guidance_scale;
  fill_step_desc(h->io_step, step, d_x_t, d_t, B, h->cfg.out_ch * H * W);
  h->io_step_on = true;
  rc = h->run(s);
  h->io_step_on = false;
  h->io_emb_dedup = false;
  h->io_cfg_pair = false;
  return rc;
}
int mf_unet_profile(mf_unet* h, const float* d_x_t, const int64_t* d_t, const int64_t* d_cond, float* d_y, int B, int H,
                    int W, void* d_workspace, size_t workspace_bytes, mf_stream_t stream, float* ms, int* kinds,
                    double* flops, int max_ops, int* n_ops) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int rc = prepare_plan(h, B, H, W, d_workspace, workspace_bytes, s);
  if (rc) return rc;
  h->io_x = d_x_t;
  h->io_t = reinterpret_cast<const long long*>(d_t);
  h->io_cond = reinterpret_cast<const long long*>(d_cond);
  h->io_y = d_y;
  return h->run_profiled(s, ms, kinds, flops, max_ops, n_ops);
}
int mf_unet_plan_info(const mf_unet* h, int* n_tc_convs, int* n_simt_convs, int* n_launches) {
  if (n_tc_convs) *n_tc_convs = h->n_tc;
  if (n_simt_convs) *n_simt_convs = h->n_simt;
  if (n_launches) *n_launches = h->n_launches;
  return 0;
}

// ---- VAE ----------------------------------------------------------------------------------------
int mf_vae_create(const mf_vae_config* cfg, mf_vae** out) {
  if (!cfg || !out) { set_error("null argument"); return 2; }
  std::unique_ptr<mf_vae> h(new mf_vae());
  int rc = h->init(*cfg);
  if (rc) return rc;
  *out = h.release();
  return 0;
}
void mf_vae_destroy(mf_vae* h) { delete h; }
int mf_vae_param_count(const mf_vae* h) { return static_cast<int>(h->params.size()); }
const char* mf_vae_param_name(const mf_vae* h, int i) {
  return (i >= 0 && i < static_cast<int>(h->params.size())) ? h->params[i]->name.c_str() : nullptr;
}
int mf_vae_param_shape(const mf_vae* h, int i, int64_t shape[4], int* ndim) {
  if (i < 0 || i >= static_cast<int>(h->params.size())) { set_error("param index out of range"); return 2; }
  const auto& s = h->params[i]->shape;
  *ndim = static_cast<int>(s.size());
  for (size_t k = 0; k < s.size(); ++k) shape[k] = s[k];
  return 0;
}
int mf_vae_set_param(mf_vae* h, const char* name, const float* d_data, const int64_t* shape, int ndim,
                     mf_stream_t stream) {
  return h->set_param(name, d_data, shape, ndim, static_cast<cudaStream_t>(stream));
}
size_t mf_vae_workspace_bytes(mf_vae* h, int B, int H, int W) {
  if (h->build(B, H, W, nullptr, true, nullptr)) return 0;
  const size_t need = h->arena.peak;
  h->key = typename mf_vae::PlanKey();
  return need;
}
int mf_vae_decode(mf_vae* h, const float* d_z, float* d_x, int B, int H, int W, void* d_workspace,
                  size_t workspace_bytes, mf_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  MF_REQUIRE(d_z && d_x && B > 0 && H > 0 && W > 0, "bad arguments");
  int rc = prepare_plan(h, B, H, W, d_workspace, workspace_bytes, s);
  if (rc) return rc;
  h->io_z = d_z;
  h->io_x = d_x;
  return h->run(s);
}
int mf_vae_decode_u8(mf_vae* h, const float* d_z, float* d_x, uint8_t* d_x_u8, int B, int H, int W, void* d_workspace,
                     size_t workspace_bytes, mf_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  MF_REQUIRE(d_z && d_x_u8 && B > 0 && H > 0 && W > 0, "bad arguments");
  MF_REQUIRE(h->cfg.out_channels <= 8 && h->cfg.hid_chs[0] % 64 == 0, "uint8 output needs the narrow 1x1 head");
  int rc = prepare_plan(h, B, H, W, d_workspace, workspace_bytes, s);
  if (rc) return rc;
  h->io_z = d_z;
  h->io_x = d_x;            // may be NULL: only the uint8 image is written
  h->io_out_u8 = d_x_u8;
  rc = h->run(s);
  h->io_out_u8 = nullptr;
  return rc;
}
size_t mf_vae_encode_workspace_bytes(mf_vae* h, int B, int H, int W) {
  if (!h->has_encoder) { set_error("this VAE handle was created without an encoder (in_channels == 0)"); return 0; }
  h->enc.version = h->version;
  if (h->enc.build(B, H, W, nullptr, true, nullptr)) return 0;
  const size_t need = h->enc.arena.peak;
  h->enc.key = typename mf_vae_enc_plan::PlanKey();
  return need;
}
int mf_vae_encode(mf_vae* h, const float* d_x, const float* d_noise, float* d_z, float* d_moments, int B, int H, int W,
                  void* d_workspace, size_t workspace_bytes, mf_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  MF_REQUIRE(h->has_encoder, "this VAE handle was created without an encoder (in_channels == 0)");
  MF_REQUIRE(d_x && d_z && B > 0 && H > 0 && W > 0, "bad arguments");
  int rc = h->check_all_set();
  if (rc) return rc;
  h->enc.version = h->version;   // parameters (and their prepared layouts) are owned by the parent handle
  rc = prepare_plan(&h->enc, B, H, W, d_workspace, workspace_bytes, s);
  if (rc) return rc;
  h->enc.io_x = d_x;
  h->enc.io_noise = d_noise;
  h->enc.io_z = d_z;
  h->enc.io_moments = d_moments;
  return h->enc.run(s);
}
int mf_vae_profile(mf_vae* h, const float* d_z, float* d_x, int B, int H, int W, void* d_workspace,
                   size_t workspace_bytes, mf_stream_t stream, float* ms, int* kinds, double* flops, int max_ops,
                   int* n_ops) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int rc = prepare_plan(h, B, H, W, d_workspace, workspace_bytes, s);
  if (rc) return rc;
  h->io_z = d_z;
  h->io_x = d_x;
  return h->run_profiled(s, ms, kinds, flops, max_ops, n_ops);
}
int mf_vae_plan_info(const mf_vae* h, int* n_tc_convs, int* n_simt_convs, int* n_launches) {
  if (n_tc_convs) *n_tc_convs = h->n_tc;
  if (n_simt_convs) *n_simt_convs = h->n_simt;
  if (n_launches) *n_launches = h->n_launches;
  return 0;
}

// ---- scheduler ----------------------------------------------------------------------------------
int mf_sched_step_opts(const mf_sched_tables* tables, const float* d_x_t, const float* d_pred,
                       const float* d_pred_uncond, float guidance_scale, const int64_t* d_t, const float* d_noise,
                       const int64_t* d_t_next, const float* d_noise_ddim, int objective_is_x0, int clip_x0,
                       float* d_x_prior, float* d_x_0, float* d_x_T, float* d_x_next, int B, int chw,
                       const mf_sched_opts* opts, mf_stream_t stream) {
  MF_REQUIRE(tables && d_x_t && d_pred && d_t, "bad arguments");
  SchedStepDesc d{};
  d.x_t = d_x_t; d.pred = d_pred; d.pred_uncond = d_pred_uncond; d.guidance = guidance_scale;
  d.t = reinterpret_cast<const long long*>(d_t);
  d.noise = d_noise;
  d.t_next = reinterpret_cast<const long long*>(d_t_next);
  d.noise2 = d_noise_ddim;
  d.objective_x0 = objective_is_x0; d.clip_x0 = clip_x0;
  d.x_prior = d_x_prior; d.x_0 = d_x_0; d.x_T = d_x_T; d.x_next = d_x_next;
  d.B = B; d.CHW = chw;
  d.tab.sqrt_recip_ac = tables->sqrt_recip_alphas_cumprod;
  d.tab.sqrt_recipm1_ac = tables->sqrt_recipm1_alphas_cumprod;
  d.tab.coef1 = tables->posterior_mean_coef1;
  d.tab.coef2 = tables->posterior_mean_coef2;
  d.tab.post_var = tables->posterior_variance;
  d.tab.betas = tables->betas;
  d.tab.alphas_cumprod = tables->alphas_cumprod;
  if (opts != nullptr) {
    MF_REQUIRE(opts->pred_batch_stride == 0 || opts->pred_batch_stride >= chw, "pred_batch_stride < chw");
    MF_REQUIRE(!opts->cold_diffusion || (opts->sqrt_alphas_cumprod && opts->sqrt_one_minus_alphas_cumprod && opts->T > 0),
               "cold_diffusion needs sqrt_alphas_cumprod / sqrt_one_minus_alphas_cumprod / T");
    MF_REQUIRE(opts->d_pred_var_uncond == nullptr || opts->d_pred_var != nullptr, "d_pred_var_uncond without d_pred_var");
    d.pred_var = opts->d_pred_var; d.pred_var_uncond = d_pred_uncond ? opts->d_pred_var_uncond : nullptr;
    d.pred_bstride = opts->pred_batch_stride;
    d.cold = opts->cold_diffusion;
    d.sqrt_ac = opts->sqrt_alphas_cumprod; d.sqrt_1mac = opts->sqrt_one_minus_alphas_cumprod; d.T = opts->T;
  }
  return sched_step(d, static_cast<cudaStream_t>(stream));
}
int mf_sched_step(const mf_sched_tables* tables, const float* d_x_t, const float* d_pred, const float* d_pred_uncond,
                  float guidance_scale, const int64_t* d_t, const float* d_noise, const int64_t* d_t_next,
                  const float* d_noise_ddim, int objective_is_x0, int clip_x0, float* d_x_prior, float* d_x_0,
                  float* d_x_T, float* d_x_next, int B, int chw, mf_stream_t stream) {
  return mf_sched_step_opts(tables, d_x_t, d_pred, d_pred_uncond, guidance_scale, d_t, d_noise, d_t_next, d_noise_ddim,
                            objective_is_x0, clip_x0, d_x_prior, d_x_0, d_x_T, d_x_next, B, chw, nullptr, stream);
}

// ---- kernel-level ops (split tensors are fp16 hi/lo planes, passed as void*) ---------------------------------------
#define MF_H(p) reinterpret_cast<__half*>(p)
#define MF_CH(p) reinterpret_cast<const __half*>(p)
int mf_op_pack_split(const float* d_x_nchw, void* d_out, int64_t plane, int N, int C, int H, int W, mf_stream_t s) {
  return pack_nchw_to_split(d_x_nchw, MF_H(d_out), plane, N, C, H, W, static_cast<cudaStream_t>(s));
}
int mf_op_unpack_nchw(const void* d_in, int64_t plane, int layout, float* d_out_nchw, int N, int C, int H, int W,
                      mf_stream_t s) {
  return unpack_to_nchw(d_in, plane, layout, d_out_nchw, N, C, H, W, static_cast<cudaStream_t>(s));
}
int mf_op_prep_weight_tc(const float* d_w_oihw, void* d_out, float* d_scales, int Cout, int Cin, int kh, int kw,
                         mf_stream_t s) {
  return prep_weight_tc(d_w_oihw, MF_H(d_out), d_scales, Cout, Cin, kh, kw, static_cast<cudaStream_t>(s));
}
int mf_op_prep_weight_simt(const float* d_w_oihw, float* d_out, int Cout, int Cin, int kh, int kw, mf_stream_t s) {
  return prep_weight_simt(d_w_oihw, d_out, Cout, Cin, kh, kw, static_cast<cudaStream_t>(s));
}
int mf_op_conv_tc_supported(int N, int H, int W, int C0, int C1, int Cout, int ksize, int stride) {
  return conv_tc_supported(N, H, W, C0, C1, Cout, ksize, stride);
}
int mf_op_conv_tc_stats_chunks(int H, int W) { return conv_tc_stats_chunks(H, W); }
int mf_op_conv_tc(const void* d_src0, int64_t src0_plane, int C0, const void* d_src1, int64_t src1_plane, int C1,
                  int N, int H, int W, const void* d_w_planes, const float* d_scales, int Cout, int ksize,
                  const float* d_bias, void* d_out, int64_t out_plane, int out_layout, float* d_stats,
                  int drain_interval, int stride, mf_stream_t s) {
  MF_REQUIRE(out_layout == kNHWCRaw || out_layout == kNHWCSplit, "conv_tc writes NHWC");
  ConvTcDesc d{};
  d.src0 = MF_CH(d_src0); d.src0_plane = src0_plane; d.C0 = C0;
  d.src1 = MF_CH(d_src1); d.src1_plane = src1_plane; d.C1 = C1;
  d.N = N; d.H = H; d.W = W; d.stride = stride;
  d.w_planes = MF_CH(d_w_planes); d.w_inv_scale = d_scales + 1; d.Cout = Cout; d.ksize = ksize; d.bias = d_bias;
  d.out = d_out; d.out_plane = out_plane; d.out_mode = out_layout == kNHWCSplit ? kOutSplit : kOutRaw;
  d.stats = d_stats;
  d.drain_interval = drain_interval;
  ConvTcPlan plan;
  int rc = conv_tc_build(d, &plan);
  if (rc) return rc;
  return conv_tc_launch(plan, static_cast<cudaStream_t>(s));
}
int mf_op_prep_weight_up_tc(const float* d_w_oihw, void* d_out, float* d_scales, int Cout, int Cin, mf_stream_t s) {
  return prep_weight_up_tc(d_w_oihw, MF_H(d_out), d_scales, Cout, Cin, static_cast<cudaStream_t>(s));
}
int mf_op_upconv_tc(const void* d_src, int64_t src_plane, int C, int N, int H, int W, const void* d_w_up_planes,
                    const float* d_scales, int Cout, const float* d_bias, void* d_out, int64_t out_plane,
                    mf_stream_t s) {
  ConvTcDesc d{};
  d.src0 = MF_CH(d_src); d.src0_plane = src_plane; d.C0 = C;
  d.N = N; d.H = H; d.W = W; d.stride = 1; d.up2 = 1;
  d.w_planes = MF_CH(d_w_up_planes); d.w_inv_scale = d_scales + 1; d.Cout = Cout; d.ksize = 3; d.bias = d_bias;
  d.out = d_out; d.out_plane = out_plane; d.out_mode = kOutSplit;
  ConvTcPlan plan;
  int rc = conv_tc_build(d, &plan);
  if (rc) return rc;
  return conv_tc_launch(plan, static_cast<cudaStream_t>(s));
}
int mf_op_conv_simt(const void* d_in, int64_t in_plane, int in_layout, int N, int Cin, int Hin, int Win,
                    const float* d_w_kc, const float* d_bias, int Cout, int ksize, int stride, void* d_out,
                    int64_t out_plane, int out_layout, mf_stream_t s) {
  ConvSimtDesc d{};
  d.in = d_in; d.in_plane = in_plane; d.in_layout = in_layout;
  d.N = N; d.Cin = Cin; d.Hin = Hin; d.Win = Win;
  d.w_kc = d_w_kc; d.bias = d_bias; d.Cout = Cout; d.ksize = ksize; d.stride = stride;
  d.out = d_out; d.out_plane = out_plane; d.out_layout = out_layout;
  return conv_simt(d, static_cast<cudaStream_t>(s));
}
int mf_op_gn_partial(const float* d_raw, float* d_partial, int N, int HW, int C, mf_stream_t s) {
  return gn_partial_from_raw(d_raw, d_partial, N, HW, C, static_cast<cudaStream_t>(s));
}
int mf_op_gn_finalize(const float* d_partial, float* d_mean_rstd, int N, int chunks, int C, int G, int HW, float eps,
                      mf_stream_t s) {
  return gn_finalize(d_partial, d_mean_rstd, N, chunks, C, G, HW, eps, static_cast<cudaStream_t>(s));
}
int mf_op_gn_apply(const float* d_raw, const float* d_mean_rstd, const float* d_gamma, const float* d_beta,
                   const void* d_res, int64_t res_plane, int res_kind, const float* d_emb, int emb_stride,
                   void* d_out, int64_t out_plane, int N, int HW, int C, int G, mf_stream_t s) {
  GnApplyDesc d{};
  d.raw = d_raw; d.mean_rstd = d_mean_rstd; d.gamma = d_gamma; d.beta = d_beta;
  d.res = d_res; d.res_plane = res_plane; d.res_kind = res_kind;
  d.emb = d_emb; d.emb_stride = emb_stride;
  d.out = MF_H(d_out); d.out_plane = out_plane; d.N = N; d.HW = HW; d.C = C; d.G = G;
  d.raw_plane = 0; d.act = 1;
  return gn_apply(d, static_cast<cudaStream_t>(s));
}
int mf_op_attention(const float* d_q, const float* d_k, const float* d_v, int row_stride, void* d_out, int64_t out_plane,
                    int B, int N, int heads, int d, mf_stream_t s) {
  return attention_core(d_q, d_k, d_v, row_stride, MF_H(d_out), out_plane, B, N, heads, d, static_cast<cudaStream_t>(s));
}
int mf_op_layernorm(const void* d_in, int64_t in_plane, const float* d_gamma, const float* d_beta, void* d_out,
                    int64_t out_plane, int64_t tokens, int C, float eps, mf_stream_t s) {
  return layernorm_split(MF_CH(d_in), in_plane, d_gamma, d_beta, MF_H(d_out), out_plane, tokens, C, eps,
                         static_cast<cudaStream_t>(s));
}
int mf_op_geglu(const float* d_in, void* d_out, int64_t out_plane, int64_t tokens, int Ch, mf_stream_t s) {
  return geglu_split(d_in, MF_H(d_out), out_plane, tokens, Ch, static_cast<cudaStream_t>(s));
}
int mf_op_vq_quantize(const float* d_z, const float* d_codebook, float* d_zq, int* d_idx, int B, int C, int HW, int K,
                      mf_stream_t s) {
  return vq_quantize(d_z, d_codebook, d_zq, d_idx, B, C, HW, K, static_cast<cudaStream_t>(s));
}
int mf_op_upsample2x(const void* d_in, int64_t in_plane, void* d_out, int64_t out_plane, int N, int H, int W, int C,
                     mf_stream_t s) {
  return upsample2x_split(MF_CH(d_in), in_plane, MF_H(d_out), out_plane, N, H, W, C, static_cast<cudaStream_t>(s));
}

}  // extern "C"
