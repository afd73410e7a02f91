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

int attention_core(const float* q, const float* k, const float* v, int row_stride, __half* out, long long out_plane,
                   int B, int N, int heads, int d, cudaStream_t s) {
  if (B == 0 || N == 0) return 0;
  MF_REQUIRE(row_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(q) & 15) == 0 && (reinterpret_cast<uintptr_t>(k) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(v) & 15) == 0,
             "attention_core: q/k/v rows must be 16-byte aligned");
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
