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
template <int D>
__global__ void __launch_bounds__(256) attention_core_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                              const float* __restrict__ v, int row_stride,
                                                              __half* __restrict__ out, long long out_plane, int N,
                                                              int heads, float scale2) {
  constexpr int KT = 32;        // keys per tile
  constexpr int DP = D + 1;     // padded row: lane j walks row j without bank conflicts
  constexpr int DL = D / 32;    // output channels per lane
  __shared__ float ks[KT][DP];
  __shared__ float vs[KT][D];
  __shared__ float qs[8][D];
  __shared__ float ps[8][KT];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qblocks = (N + 7) / 8;
  const int qb = blockIdx.x % qblocks;
  const int h = (blockIdx.x / qblocks) % heads;
  const int b = blockIdx.x / (qblocks * heads);
  const int qi = qb * 8 + warp;
  const bool qvalid = qi < N;
  const long long tok0 = static_cast<long long>(b) * N;
  if (qvalid)
    for (int dd = lane; dd < D; dd += 32) qs[warp][dd] = q[(tok0 + qi) * row_stride + h * D + dd];
  float m = -INFINITY, l = 0.f;
  float acc[DL];
#pragma unroll
  for (int i = 0; i < DL; ++i) acc[i] = 0.f;

  for (int j0 = 0; j0 < N; j0 += KT) {
    __syncthreads();
    for (int e = threadIdx.x; e < KT * D; e += 256) {
      const int jj = e / D, dd = e % D;
      const bool ok = j0 + jj < N;
      ks[jj][dd] = ok ? k[(tok0 + j0 + jj) * row_stride + h * D + dd] : 0.f;
      vs[jj][dd] = ok ? v[(tok0 + j0 + jj) * row_stride + h * D + dd] : 0.f;
    }
    __syncthreads();
    if (qvalid) {
      // lane j: score of key j0 + j.  (q*s).(k*s) == s^2 * (q.k); the reference scales both operands by d^-0.25
      float sc = 0.f;
#pragma unroll 8
      for (int dd = 0; dd < D; ++dd) sc = fmaf(qs[warp][dd], ks[lane][dd], sc);
      sc = (j0 + lane < N) ? sc * scale2 : -INFINITY;
      float tmax = sc;
      for (int off = 16; off > 0; off >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, off));
      const float mnew = fmaxf(m, tmax);
      const float p = expf(sc - mnew);
      float psum = p;
      for (int off = 16; off > 0; off >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, off);
      const float corr = expf(m - mnew);  // 0 on the first tile (m = -inf)
      l = l * corr + psum;
      m = mnew;
      ps[warp][lane] = p;
      __syncwarp();
#pragma unroll
      for (int i = 0; i < DL; ++i) acc[i] *= corr;
#pragma unroll 8
      for (int jj = 0; jj < KT; ++jj) {
        const float pj = ps[warp][jj];
#pragma unroll
        for (int i = 0; i < DL; ++i) acc[i] = fmaf(pj, vs[jj][lane + 32 * i], acc[i]);
      }
      __syncwarp();
    }
  }
  if (qvalid) {
    const float inv = 1.0f / l;
    __half* oh = out + (tok0 + qi) * (static_cast<long long>(heads) * D) + h * D;
#pragma unroll
    for (int i = 0; i < DL; ++i) split16(acc[i] * inv, oh[lane + 32 * i], oh[out_plane + lane + 32 * i]);
  }
}

int attention_core(const float* q, const float* k, const float* v, int row_stride, __half* out, long long out_plane,
                   int B, int N, int heads, int d, cudaStream_t s) {
  if (B == 0 || N == 0) return 0;
  const float scale2 = 1.0f / sqrtf(static_cast<float>(d));  // (d^-0.25)^2
  const int grid = B * heads * ((N + 7) / 8);
  switch (d) {
    case 32: attention_core_kernel<32><<<grid, 256, 0, s>>>(q, k, v, row_stride, out, out_plane, N, heads, scale2); break;
    case 64: attention_core_kernel<64><<<grid, 256, 0, s>>>(q, k, v, row_stride, out, out_plane, N, heads, scale2); break;
    case 128: attention_core_kernel<128><<<grid, 256, 0, s>>>(q, k, v, row_stride, out, out_plane, N, heads, scale2); break;
    default:
      set_error("attention_core: head dim must be 32, 64 or 128");
      return 2;
  }
  MF_CUDA_OK(cudaGetLastError());
  return 0;
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

}  // namespace mf
