#include "mf_attn.cuh"

#include <algorithm>

namespace mf {

// =================================================================================================
// LayerNorm: one warp per token, three passes over an L1-resident row (mean, centred variance, write)
// =================================================================================================
__global__ void layernorm_split_kernel(const __half* __restrict__ in, long long in_plane, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, __half* __restrict__ out, long long out_plane,
                                       long long tokens, int C, float eps) {
  const long long tok = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (tok >= tokens) return;
  const __half* hi = in + tok * C;
  const __half* lo = hi + in_plane;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += join16(hi[c], lo[c]);
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  const float mean = s / static_cast<float>(C);
  float ss = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float d = join16(hi[c], lo[c]) - mean;
    ss = fmaf(d, d, ss);
  }
  for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
  const float rstd = rsqrtf(ss / static_cast<float>(C) + eps);
  __half* oh = out + tok * C;
  __half* ol = oh + out_plane;
  for (int c = lane; c < C; c += 32) {
    const float y = (join16(hi[c], lo[c]) - mean) * rstd * gamma[c] + beta[c];
    split16(y, oh[c], ol[c]);
  }
}

int layernorm_split(const __half* in, long long in_plane, const float* gamma, const float* beta, __half* out,
                    long long out_plane, long long tokens, int C, float eps, cudaStream_t s) {
  if (tokens == 0) return 0;
  const long long threads = tokens * 32;
  layernorm_split_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, s>>>(in, in_plane, gamma, beta, out,
                                                                                      out_plane, tokens, C, eps);
  MF_CUDA_OK(cudaGetLastError());
  return 0;
}

// =================================================================================================
// GEGLU gate: x * gelu(gate), exact (erf) GELU like F.gelu's default
// =================================================================================================
__global__ void geglu_split_kernel(const float* __restrict__ in, __half* __restrict__ out, long long out_plane,
                                   long long tokens, int Ch) {
  const int c4n = Ch / 4;
  const long long total = tokens * c4n;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long tok = i / c4n;
    const int c = static_cast<int>(i % c4n) * 4;
    const float4 a = *reinterpret_cast<const float4*>(in + tok * 2 * Ch + c);
    const float4 g = *reinterpret_cast<const float4*>(in + tok * 2 * Ch + Ch + c);
    const float av[4] = {a.x, a.y, a.z, a.w}, gv[4] = {g.x, g.y, g.z, g.w};
    float y[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) y[j] = av[j] * (0.5f * gv[j] * (1.0f + erff(gv[j] * 0.70710678118654752440f)));
    st_split4(out + tok * Ch + c, out + out_plane + tok * Ch + c, make_float4(y[0], y[1], y[2], y[3]));
  }
}

int geglu_split(const float* in, __half* out, long long out_plane, long long tokens, int Ch, cudaStream_t s) {
  MF_REQUIRE(Ch % 4 == 0, "geglu: channel count must be a multiple of 4");
  const long long total = tokens * (Ch / 4);
  if (total == 0) return 0;
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 32));
  geglu_split_kernel<<<blocks, 256, 0, s>>>(in, out, out_plane, tokens, Ch);
  MF_CUDA_OK(cudaGetLastError());
  return 0;
}

// =================================================================================================
// Attention core: block = 8 warps = 8 queries of one (sample, head); keys/values streamed through shared memory in
// tiles of 32; lane j scores key j of the tile, the softmax is kept online (running max / sum), fp32 throughout.
// =================================================================================================
template <int D, int QPW>
__global__ void __launch_bounds__(256) attention_core_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                              const float* __restrict__ v, int row_stride,
                                                              __half* __restrict__ out, long long out_plane, int N,
                                                              int heads, float scale2) {
  // QPW queries per warp (8 * QPW per block) share every staged K/V tile: with QPW = 1 a block re-read all keys and
  // values of its (batch, head) for only 8 queries and the kernel was L2-bandwidth bound (2.1 GB per launch at config 4).
  constexpr int KT = 32;        // keys per tile
  constexpr int DP = D + 4;     // padded key row (16-byte aligned; lane j reads row j with 128-bit loads)
  constexpr int DL = D / 32;    // output channels per lane
  constexpr int QB = 8 * QPW;   // queries per block
  extern __shared__ __align__(16) float att_smem[];
  float (*ks)[DP] = reinterpret_cast<float (*)[DP]>(att_smem);
  float (*vs)[D] = reinterpret_cast<float (*)[D]>(att_smem + KT * DP);
  float (*qs)[D] = reinterpret_cast<float (*)[D]>(att_smem + KT * DP + KT * D);
  float (*ps)[KT] = reinterpret_cast<float (*)[KT]>(att_smem + KT * DP + KT * D + QB * D);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qblocks = (N + QB - 1) / QB;
  const int qb = blockIdx.x % qblocks;
  const int h = (blockIdx.x / qblocks) % heads;
  const int b = blockIdx.x / (qblocks * heads);
  const int q0 = qb * QB + warp * QPW;          // first query of this warp
  const long long tok0 = static_cast<long long>(b) * N;
#pragma unroll
  for (int r = 0; r < QPW; ++r)
    if (q0 + r < N)
      for (int dd = lane; dd < D; dd += 32) qs[warp * QPW + r][dd] = q[(tok0 + q0 + r) * row_stride + h * D + dd];
    else
      for (int dd = lane; dd < D; dd += 32) qs[warp * QPW + r][dd] = 0.f;
  float m[QPW], l[QPW], acc[QPW][DL];
#pragma unroll
  for (int r = 0; r < QPW; ++r) {
    m[r] = -INFINITY;
    l[r] = 0.f;
#pragma unroll
    for (int i = 0; i < DL; ++i) acc[r][i] = 0.f;
  }

  for (int j0 = 0; j0 < N; j0 += KT) {
    __syncthreads();
    for (int e = threadIdx.x; e < KT * (D / 4); e += 256) {
      const int jj = e / (D / 4), d4 = (e % (D / 4)) * 4;
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (j0 + jj < N) {
        kv = *reinterpret_cast<const float4*>(k + (tok0 + j0 + jj) * row_stride + h * D + d4);
        vv = *reinterpret_cast<const float4*>(v + (tok0 + j0 + jj) * row_stride + h * D + d4);
      }
      *reinterpret_cast<float4*>(&ks[jj][d4]) = kv;
      *reinterpret_cast<float4*>(&vs[jj][d4]) = vv;
    }
    __syncthreads();
    // lane j: scores of key j0 + j for the warp's QPW queries.  (q*s).(k*s) == s^2 * (q.k); the reference scales both
    // operands by d^-0.25.  Per query the summation order over d is sequential, as in the single-query version.
    float sc[QPW];
#pragma unroll
    for (int r = 0; r < QPW; ++r) sc[r] = 0.f;
#pragma unroll 4
    for (int dd = 0; dd < D; dd += 4) {
      const float4 kk = *reinterpret_cast<const float4*>(&ks[lane][dd]);
#pragma unroll
      for (int r = 0; r < QPW; ++r) {
        const float4 qq = *reinterpret_cast<const float4*>(&qs[warp * QPW + r][dd]);
        sc[r] = fmaf(qq.x, kk.x, sc[r]);
        sc[r] = fmaf(qq.y, kk.y, sc[r]);
        sc[r] = fmaf(qq.z, kk.z, sc[r]);
        sc[r] = fmaf(qq.w, kk.w, sc[r]);
      }
    }
    float corr[QPW];
#pragma unroll
    for (int r = 0; r < QPW; ++r) {
      const float s1 = (j0 + lane < N) ? sc[r] * scale2 : -INFINITY;
      float tmax = s1;
      for (int off = 16; off > 0; off >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, off));
      const float mnew = fmaxf(m[r], tmax);
      const float p = expf(s1 - mnew);
      float psum = p;
      for (int off = 16; off > 0; off >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, off);
      corr[r] = expf(m[r] - mnew);  // 0 on the first tile (m = -inf)
      l[r] = l[r] * corr[r] + psum;
      m[r] = mnew;
      ps[warp * QPW + r][lane] = p;
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < QPW; ++r)
#pragma unroll
      for (int i = 0; i < DL; ++i) acc[r][i] *= corr[r];
#pragma unroll 2
    for (int jj = 0; jj < KT; jj += 4) {
      float vv[4][DL];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int i = 0; i < DL; ++i) vv[u][i] = vs[jj + u][lane + 32 * i];
#pragma unroll
      for (int r = 0; r < QPW; ++r) {
        const float4 pj = *reinterpret_cast<const float4*>(&ps[warp * QPW + r][jj]);
#pragma unroll
        for (int i = 0; i < DL; ++i) {
          acc[r][i] = fmaf(pj.x, vv[0][i], acc[r][i]);
          acc[r][i] = fmaf(pj.y, vv[1][i], acc[r][i]);
          acc[r][i] = fmaf(pj.z, vv[2][i], acc[r][i]);
          acc[r][i] = fmaf(pj.w, vv[3][i], acc[r][i]);
        }
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int r = 0; r < QPW; ++r) {
    if (q0 + r < N) {
      const float inv = 1.0f / l[r];
      __half* oh = out + (tok0 + q0 + r) * (static_cast<long long>(heads) * D) + h * D;
#pragma unroll
      for (int i = 0; i < DL; ++i) split16(acc[r][i] * inv, oh[lane + 32 * i], oh[out_plane + lane + 32 * i]);
    }
  }
}

template <int D, int QPW>
static int attention_launch(const float* q, const float* k, const float* v, int row_stride, __half* out,
                            long long out_plane, int B, int N, int heads, float scale2, cudaStream_t s) {
  constexpr int KT = 32, QB = 8 * QPW;
  constexpr size_t smem = sizeof(float) * (KT * (D + 4) + KT * D + QB * D + QB * KT);
  static bool attr_set = false;
  if (!attr_set) {
    MF_CUDA_OK(cudaFuncSetAttribute(attention_core_kernel<D, QPW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(smem)));
    attr_set = true;
  }
  const int grid = B * heads * ((N + QB - 1) / QB);
  attention_core_kernel<D, QPW><<<grid, 256, smem, s>>>(q, k, v, row_stride, out, out_plane, N, heads, scale2);
  MF_CUDA_OK(cudaGetLastError());
  return 0;
}

// =================================================================================================
// Attention core on tcgen05 (N in {64, 128, 192, 256} tokens, d in {64, 128}): one CTA per (sample, head, block of 128
// queries).  S = (q s)(k s)^T and O = softmax(S) v both run as fp16x3 tensor-core products (hi*hi + hi*lo + lo*hi, fp32
// accumulation in TMEM), the softmax in between is fp32 in registers, in the reference's order: probabilities are normalised
// BEFORE the second product (attention_blocks.py:38-40).
//   stage 1  q, k (scaled by d^-0.25 in fp32, like the reference scales both operands) -> fp16 hi/lo planes, K-major
//            128-byte-swizzled tiles in shared memory (written by the threads: the q|k|v projection is a raw fp32 tensor)
//   stage 2  one thread issues the S MMAs (M = 128 queries, N = keys, K = d) -> TMEM columns [0, N)
//   stage 3  8 warps = 4 TMEM lane quarters x 2 key halves: tcgen05.ld, row max / row sum exchanged between the two
//            halves through shared memory, p = exp(s - max) / sum -> hi/lo tiles (A operand of the second product)
//   stage 4  v -> V^T hi/lo tiles (keys are the K dimension), in chunks of 128 keys; O MMAs (N = d) -> TMEM columns
//            [N, N + d); each 128-key chunk is drained and added in fp32 registers (short TMEM accumulation chains)
//   stage 5  O -> hi/lo split planes [token][heads*d]
// Shared memory: stage 1 needs (d/64) * 2 * (16 KB + N * 128 B), stage 3/4 N/64 * 32 KB + 64 KB: 192 KB at N = 256, d = 128.
// =================================================================================================
constexpr int kAtQ = 128;           // queries per CTA (UMMA M)
constexpr int kAtThreads = 256;
constexpr int kAtAux = 4096;        // mbarrier, TMEM slot, softmax exchange [2][128] x 2

__device__ __forceinline__ uint32_t at_sw128(int row, int k) {   // byte offset of fp16 element (row, k < 64) in a SW128 tile
  return static_cast<uint32_t>(row) * 128u + (((static_cast<uint32_t>(k) >> 3) ^ (static_cast<uint32_t>(row) & 7u)) << 4) +
         ((static_cast<uint32_t>(k) & 7u) << 1);
}

template <int D>
__global__ void __launch_bounds__(kAtThreads)
attention_tc_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v, int row_stride,
                    __half* __restrict__ out, long long out_plane, int N, int heads, float scale) {
  constexpr int KB = D / 64;                 // K blocks of the first product
  extern __shared__ __align__(1024) uint8_t at_smem[];
  if ((smem_u32(at_smem) & 1023u) != 0u) __trap();
  const int nkb = N / 64;                    // K blocks (of 64 keys) of the second product
  // stage-1 layout: Q tiles [plane][kb] (16 KB each) then K tiles [plane][kb] (N * 128 B each)
  uint8_t* q_t = at_smem;
  uint8_t* k_t = at_smem + 2 * KB * 16384;
  // stage-3/4 layout: P tiles [plane][key block] (16 KB each) then V^T tiles [plane][key block of the chunk] (D * 128 B each)
  uint8_t* p_t = at_smem;
  uint8_t* v_t = at_smem + 2 * nkb * 16384;
  const size_t main_bytes = max(static_cast<size_t>(2 * KB * 16384 + 2 * KB * N * 128),
                                static_cast<size_t>(2 * nkb * 16384 + 2 * 2 * D * 128));
  uint8_t* aux = at_smem + ((main_bytes + 1023) & ~static_cast<size_t>(1023));
  uint64_t* mma_bar = reinterpret_cast<uint64_t*>(aux);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux + 16);
  float* xmax = reinterpret_cast<float*>(aux + 1024);        // [2 halves][128 rows]
  float* xsum = reinterpret_cast<float*>(aux + 2048);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qblocks = (N + kAtQ - 1) / kAtQ;
  const int qb = blockIdx.x % qblocks;
  const int h = (blockIdx.x / qblocks) % heads;
  const int b = blockIdx.x / (qblocks * heads);
  const long long tok0 = static_cast<long long>(b) * N;
  const int q0 = qb * kAtQ;

  if (threadIdx.x == 0) {
    mbar_init(mma_bar, 1);
    fence_mbar_init();
  }
  const uint32_t tmem_cols = (N + D <= 256) ? 256u : 512u;   // S: N columns, O: D columns behind it
  if (warp == 1) { tmem_alloc(tmem_slot, tmem_cols); tmem_relinquish(); }

  // ---- stage 1: q (this block's 128 rows, zero beyond N) and k (all N rows), scaled, split, swizzled
  // 4 independent 16-byte loads in flight per thread (8 warps per SM cannot hide the latency of one load per iteration)
  {
    constexpr int ITEMS_Q = kAtQ * (D / 4);
    for (int e0 = threadIdx.x; e0 < ITEMS_Q; e0 += 4 * kAtThreads) {
      float4 val[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = e0 + u * kAtThreads;
        const int r = e / (D / 4), d4 = (e % (D / 4)) * 4;
        val[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < ITEMS_Q && q0 + r < N) val[u] = *reinterpret_cast<const float4*>(q + (tok0 + q0 + r) * row_stride + h * D + d4);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = e0 + u * kAtThreads;
        if (e >= ITEMS_Q) break;
        const int r = e / (D / 4), d4 = (e % (D / 4)) * 4;
        uint32_t hi[2], lo[2];
        split16x2(val[u].x * scale, val[u].y * scale, hi[0], lo[0]);
        split16x2(val[u].z * scale, val[u].w * scale, hi[1], lo[1]);
        const int kb = d4 / 64, kk = d4 % 64;
        const uint32_t off = at_sw128(r, kk);
        *reinterpret_cast<uint2*>(q_t + (0 * KB + kb) * 16384 + off) = make_uint2(hi[0], hi[1]);
        *reinterpret_cast<uint2*>(q_t + (1 * KB + kb) * 16384 + off) = make_uint2(lo[0], lo[1]);
      }
    }
    const int items_k = N * (D / 4);
    for (int e0 = threadIdx.x; e0 < items_k; e0 += 4 * kAtThreads) {
      float4 val[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = e0 + u * kAtThreads;
        const int r = e / (D / 4), d4 = (e % (D / 4)) * 4;
        val[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < items_k) val[u] = *reinterpret_cast<const float4*>(k + (tok0 + r) * row_stride + h * D + d4);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = e0 + u * kAtThreads;
        if (e >= items_k) break;
        const int r = e / (D / 4), d4 = (e % (D / 4)) * 4;
        uint32_t hi[2], lo[2];
        split16x2(val[u].x * scale, val[u].y * scale, hi[0], lo[0]);
        split16x2(val[u].z * scale, val[u].w * scale, hi[1], lo[1]);
        const int kb = d4 / 64, kk = d4 % 64;
        const uint32_t off = at_sw128(r, kk);
        *reinterpret_cast<uint2*>(k_t + static_cast<size_t>(0 * KB + kb) * N * 128 + off) = make_uint2(hi[0], hi[1]);
        *reinterpret_cast<uint2*>(k_t + static_cast<size_t>(1 * KB + kb) * N * 128 + off) = make_uint2(lo[0], lo[1]);
      }
    }
  }
  fence_proxy_async_smem();          // generic-proxy writes -> visible to the tensor core's async-proxy reads
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + N;

  // ---- stage 2: S = Q K^T
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_f16(kAtQ, N);
    bool first = true;
    for (int kb = 0; kb < KB; ++kb) {
      const uint64_t a_hi = umma_smem_desc_sw128(smem_u32(q_t + (0 * KB + kb) * 16384));
      const uint64_t a_lo = umma_smem_desc_sw128(smem_u32(q_t + (1 * KB + kb) * 16384));
      const uint64_t b_hi = umma_smem_desc_sw128(smem_u32(k_t + static_cast<size_t>(0 * KB + kb) * N * 128));
      const uint64_t b_lo = umma_smem_desc_sw128(smem_u32(k_t + static_cast<size_t>(1 * KB + kb) * N * 128));
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t koff = static_cast<uint64_t>(ks * 2);
        umma_f16(tmem_base, a_lo + koff, b_hi + koff, idesc, first ? 0u : 1u);
        umma_f16(tmem_base, a_hi + koff, b_lo + koff, idesc, 1u);
        first = false;
      }
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t koff = static_cast<uint64_t>(ks * 2);
        umma_f16(tmem_base, a_hi + koff, b_hi + koff, idesc, 1u);
      }
    }
    umma_commit(mma_bar);
  }
  mbar_wait(mma_bar, 0);
  tc_fence_after();

  // ---- stage 3: softmax over the keys; thread = (row, half of the keys)
  const int qd = warp & 3, half = warp >> 2;
  const int row = qd * 32 + lane;
  const int hcols = N / 2;                      // keys per half: 32, 64, 96 or 128
  float sv[128];
  {
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(qd * 32) << 16) + half * hcols;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (c * 32 < hcols) {
        float t32[32];
        tmem_ld_32x32(taddr + c * 32, t32);
#pragma unroll
        for (int i = 0; i < 32; ++i) sv[c * 32 + i] = t32[i];
      }
    }
  }
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < 4; ++c)
    if (c * 32 < hcols) {
#pragma unroll
      for (int i = 0; i < 32; ++i) mx = fmaxf(mx, sv[c * 32 + i]);
    }
  xmax[half * 128 + row] = mx;
  tc_fence_before();
  __syncthreads();           // also: every warp has finished reading S from TMEM and Q / K tiles are dead
  mx = fmaxf(xmax[row], xmax[128 + row]);
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < 4; ++c)
    if (c * 32 < hcols) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float pv = expf(sv[c * 32 + i] - mx);
        sv[c * 32 + i] = pv;
        sum += pv;
      }
    }
  xsum[half * 128 + row] = sum;
  __syncthreads();
  const float inv = 1.0f / (xsum[row] + xsum[128 + row]);
  // P (normalised) -> hi/lo A-operand tiles [key block][128 rows x 64 keys]
#pragma unroll
  for (int c = 0; c < 4; ++c)
    if (c * 32 < hcols) {
#pragma unroll
      for (int g8 = 0; g8 < 4; ++g8) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          split16x2(sv[c * 32 + g8 * 8 + 2 * j] * inv, sv[c * 32 + g8 * 8 + 2 * j + 1] * inv, hi[j], lo[j]);
        const int key = half * hcols + c * 32 + g8 * 8;
        const int kb = key / 64, kk = key % 64;
        const uint32_t off = at_sw128(row, kk);
        *reinterpret_cast<uint4*>(p_t + (0 * nkb + kb) * 16384 + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(p_t + (1 * nkb + kb) * 16384 + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
    }

  // ---- stage 4: O = P V, 128 keys at a time; thread = (row, half of the d output columns)
  constexpr int OC = D / 2;                     // output columns per thread
  float oacc[OC];
#pragma unroll
  for (int i = 0; i < OC; ++i) oacc[i] = 0.f;
  uint32_t bar_phase = 1;
  for (int key0 = 0; key0 < N; key0 += 128) {
    const int ckeys = min(128, N - key0);
    const int cb = ckeys / 64;                  // key blocks in this chunk
    // V^T tiles: element (n = output column dd, k = key j) of block jb at v_t[plane][jb][dd*128 + swizzle(j)].  An item
    // = (group of 8 keys, one output column): 8 loads that are coalesced ACROSS the warp (lanes = consecutive columns of
    // one key row) fill one 16-byte chunk per plane; rows 128 B apart + the chunk XOR make the stores conflict-free.
    for (int e = threadIdx.x; e < (ckeys / 8) * D; e += kAtThreads) {
      const int dd = e % D, g8 = e / D;
      float vals[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) vals[u] = v[(tok0 + key0 + g8 * 8 + u) * row_stride + h * D + dd];
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) split16x2(vals[2 * u], vals[2 * u + 1], hi[u], lo[u]);
      const int jb = (g8 * 8) / 64, jk = (g8 * 8) % 64;
      const uint32_t off = at_sw128(dd, jk);
      *reinterpret_cast<uint4*>(v_t + static_cast<size_t>(0 * 2 + jb) * D * 128 + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(v_t + static_cast<size_t>(1 * 2 + jb) * D * 128 + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
    fence_proxy_async_smem();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) {
      const uint32_t idesc = umma_idesc_f16(kAtQ, D);
      bool first = true;
      for (int jb = 0; jb < cb; ++jb) {
        const int kbg = key0 / 64 + jb;
        const uint64_t a_hi = umma_smem_desc_sw128(smem_u32(p_t + (0 * nkb + kbg) * 16384));
        const uint64_t a_lo = umma_smem_desc_sw128(smem_u32(p_t + (1 * nkb + kbg) * 16384));
        const uint64_t b_hi = umma_smem_desc_sw128(smem_u32(v_t + static_cast<size_t>(0 * 2 + jb) * D * 128));
        const uint64_t b_lo = umma_smem_desc_sw128(smem_u32(v_t + static_cast<size_t>(1 * 2 + jb) * D * 128));
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t koff = static_cast<uint64_t>(ks * 2);
          umma_f16(tmem_o, a_lo + koff, b_hi + koff, idesc, first ? 0u : 1u);
          umma_f16(tmem_o, a_hi + koff, b_lo + koff, idesc, 1u);
          first = false;
        }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t koff = static_cast<uint64_t>(ks * 2);
          umma_f16(tmem_o, a_hi + koff, b_hi + koff, idesc, 1u);
        }
      }
      umma_commit(mma_bar);
    }
    mbar_wait(mma_bar, bar_phase);
    bar_phase ^= 1;
    tc_fence_after();
    {
      const uint32_t taddr = tmem_o + (static_cast<uint32_t>(qd * 32) << 16) + half * OC;
#pragma unroll
      for (int c = 0; c < OC / 32; ++c) {
        float t32[32];
        tmem_ld_32x32(taddr + c * 32, t32);
#pragma unroll
        for (int i = 0; i < 32; ++i) oacc[c * 32 + i] += t32[i];
      }
    }
    tc_fence_before();
    __syncthreads();         // V^T tiles and the O accumulator may be overwritten by the next chunk
  }

  // ---- stage 5: O -> split planes
  if (q0 + row < N) {
    __half* oh = out + (tok0 + q0 + row) * (static_cast<long long>(heads) * D) + h * D + half * OC;
#pragma unroll
    for (int i = 0; i < OC; i += 4)
      st_split4(oh + i, oh + out_plane + i, make_float4(oacc[i], oacc[i + 1], oacc[i + 2], oacc[i + 3]));
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { __syncwarp(); tmem_dealloc(tmem_base, tmem_cols); }
}

int g_attn_tc = 1;   // 1: tcgen05 attention core where the shape allows it (N in {64..256} step 64, d in {64, 128})

template <int D>
static int attention_tc_launch(const float* q, const float* k, const float* v, int row_stride, __half* out,
                               long long out_plane, int B, int N, int heads, cudaStream_t s) {
  const int KB = D / 64, nkb = N / 64;
  const size_t main_bytes = std::max(static_cast<size_t>(2 * KB * 16384 + 2 * KB * N * 128),
                                     static_cast<size_t>(2 * nkb * 16384 + 2 * 2 * D * 128));
  const size_t smem = ((main_bytes + 1023) & ~static_cast<size_t>(1023)) + kAtAux;
  MF_REQUIRE(smem <= 227 * 1024, "attention_tc: shared memory budget");
  static size_t attr_set = 0;
  if (attr_set < smem) {
    MF_CUDA_OK(cudaFuncSetAttribute(attention_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr_set = smem;
  }
  const float scale = 1.0f / sqrtf(sqrtf(static_cast<float>(D)));   // d^-0.25 (attention_blocks.py:38)
  const int grid = B * heads * ((N + kAtQ - 1) / kAtQ);
  attention_tc_kernel<D><<<grid, kAtThreads, smem, s>>>(q, k, v, row_stride, out, out_plane, N, heads, scale);
  MF_CUDA_OK(cudaGetLastError());
  return 0;
}

int attention_core(const float* q, const float* k, const float* v, int row_stride, __half* out, long long out_plane,
                   int B, int N, int heads, int d, cudaStream_t s) {
  if (B == 0 || N == 0) return 0;
  MF_REQUIRE(row_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(q) & 15) == 0 && (reinterpret_cast<uintptr_t>(k) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(v) & 15) == 0,
             "attention_core: q/k/v rows must be 16-byte aligned");
  if (g_attn_tc && (d == 64 || d == 128) && N >= 64 && N <= 256 && N % 64 == 0) {
    return d == 64 ? attention_tc_launch<64>(q, k, v, row_stride, out, out_plane, B, N, heads, s)
                   : attention_tc_launch<128>(q, k, v, row_stride, out, out_plane, B, N, heads, s);
  }
  const float scale2 = 1.0f / sqrtf(static_cast<float>(d));  // (d^-0.25)^2
  // 4-8 queries per warp (32-64 per block; 4 at d = 128 keeps two blocks per SM resident) once there are enough
  // queries to fill the GPU, else 2 per warp
  const bool wide = static_cast<long long>(B) * heads * ((N + (d == 128 ? 31 : 63)) / (d == 128 ? 32 : 64)) >= 148;
  switch (d) {
    case 32:
      return wide ? attention_launch<32, 8>(q, k, v, row_stride, out, out_plane, B, N, heads, scale2, s)
                  : attention_launch<32, 2>(q, k, v, row_stride, out, out_plane, B, N, heads, scale2, s);
    case 64:
      return wide ? attention_launch<64, 8>(q, k, v, row_stride, out, out_plane, B, N, heads, scale2, s)
                  : attention_launch<64, 2>(q, k, v, row_stride, out, out_plane, B, N, heads, scale2, s);
    case 128:
      return wide ? attention_launch<128, 4>(q, k, v, row_stride, out, out_plane, B, N, heads, scale2, s)
                  : attention_launch<128, 2>(q, k, v, row_stride, out, out_plane, B, N, heads, scale2, s);
    default:
      set_error("attention_core: head dim must be 32, 64 or 128");
      return 2;
  }
}

// =================================================================================================
// out = in + bias[n][c]
// =================================================================================================
__global__ void add_channel_bias_split_kernel(const __half* __restrict__ in, long long in_plane,
                                              const float* __restrict__ bias, int bias_stride, __half* __restrict__ out,
                                              long long out_plane, int HW, int C, long long total) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const int n = static_cast<int>(i / (static_cast<long long>(HW) * C));
    const float y = join16(in[i], in[in_plane + i]) + bias[static_cast<long long>(n) * bias_stride + c];
    split16(y, out[i], out[out_plane + i]);
  }
}

int add_channel_bias_split(const __half* in, long long in_plane, const float* bias, int bias_stride, __half* out,
                           long long out_plane, int N, int HW, int C, cudaStream_t s) {
  const long long total = static_cast<long long>(N) * HW * C;
  if (total == 0) return 0;
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 32));
  add_channel_bias_split_kernel<<<blocks, 256, 0, s>>>(in, in_plane, bias, bias_stride, out, out_plane, HW, C, total);
  MF_CUDA_OK(cudaGetLastError());
  return 0;
}

MF_DEFINE_SATURATION_READER(sat_read_attn)

}  // namespace mf
