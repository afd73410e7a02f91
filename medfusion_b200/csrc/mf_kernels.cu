#include "mf_kernels.cuh"

namespace mf {

static inline int ceil_div_ll(long long a, long long b) { return static_cast<int>((a + b - 1) / b); }

__device__ __forceinline__ float swish(float x) { return x / (1.0f + expf(-x)); }
// streaming variant for the GroupNorm-apply kernel (which is instruction-issue bound, not bandwidth bound): same accurate
// expf, but the IEEE division (about 14 instructions plus a slow-path call) becomes MUFU.RCP + FMUL (<= 2 ulp).
__device__ __forceinline__ float swish_stream(float x) { return __fdividef(x, 1.0f + expf(-x)); }

// =================================================================================================
// Layout packing
// =================================================================================================
__global__ void pack_nchw_to_split_kernel(const float* __restrict__ x, __half* __restrict__ out, long long plane, int N,
                                          int C, int H, int W, int Cpad, int n_mod) {
  const long long total = static_cast<long long>(N) * H * W * Cpad;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % Cpad);
    const long long pix = i / Cpad;
    const int w = static_cast<int>(pix % W);
    const int h = static_cast<int>((pix / W) % H);
    int n = static_cast<int>(pix / (static_cast<long long>(W) * H));
    if (n_mod > 0) n %= n_mod;   // classifier-free guidance as one 2B batch: both halves read the same B samples
    const float v = c < C ? x[((static_cast<long long>(n) * C + c) * H + h) * W + w] : 0.f;  // zero-padded channels
    split16(v, out[i], out[plane + i]);
  }
}

// Cpad % 8 == 0: one thread per (pixel, channel octet) — the eight lanes of a pixel write its 128-byte line with one
// 16-byte store per plane each (the element-per-thread kernel above wrote 2 bytes per thread: 32 us for the 16.8 MB
// stem input of the canonical UNet at B = 64, profiles/r02_ncu_unet_step.md); only octets below C read anything.
__global__ void pack_nchw_to_split8_kernel(const float* __restrict__ x, __half* __restrict__ out, long long plane, int N,
                                           int C, int H, int W, int Cpad, int n_mod) {
  const int octs = Cpad / 8;
  const long long HW = static_cast<long long>(H) * W;
  const long long total = static_cast<long long>(N) * HW * octs;
  bool clamped = false;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c0 = static_cast<int>(i % octs) * 8;
    const long long pix = i / octs;
    const long long hw = pix % HW;
    int n = static_cast<int>(pix / HW);
    if (n_mod > 0) n %= n_mod;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (c0 + j < C) ? __ldg(x + (static_cast<long long>(n) * C + c0 + j) * HW + hw) : 0.f;
    uint4 oh, ol;
    uint32_t* ph = reinterpret_cast<uint32_t*>(&oh);
    uint32_t* pl = reinterpret_cast<uint32_t*>(&ol);
#pragma unroll
    for (int j = 0; j < 4; ++j) split16x2_flag(v[2 * j], v[2 * j + 1], ph[j], pl[j], clamped);
    *reinterpret_cast<uint4*>(out + pix * Cpad + c0) = oh;
    *reinterpret_cast<uint4*>(out + plane + pix * Cpad + c0) = ol;
  }
  sat16_report(clamped);
}

int pack_nchw_to_split(const float* x, __half* out, long long plane, int N, int C, int H, int W, cudaStream_t s,
                       int Cpad, int n_mod) {
  if (Cpad < C) Cpad = C;
  const long long total = static_cast<long long>(N) * H * W * Cpad;
  if (total == 0) return 0;
  if (Cpad % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && plane % 8 == 0) {
    const int blocks = static_cast<int>(std::min<long long>((total / 8 + 255) / 256, 148 * 16));
    pack_nchw_to_split8_kernel<<<blocks, 256, 0, s>>>(x, out, plane, N, C, H, W, Cpad, n_mod);
  } else {
    const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 16));
    pack_nchw_to_split_kernel<<<blocks, 256, 0, s>>>(x, out, plane, N, C, H, W, Cpad, n_mod);
  }
  MF_CUDA_OK(cudaGetLastError());
  return 0;
}

__global__ void unpack_to_nchw_kernel(const void* __restrict__ in_v, long long plane, int layout, float* __restrict__ out,
                                      int N, int C, int H, int W) {
  const float* in = reinterpret_cast<const float*>(in_v);
  const __half* inh = reinterpret_cast<const __half*>(in_v);
  const long long total = static_cast<long long>(N) * H * W * C;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const long long pix = i / C;
    const int w = static_cast<int>(pix % W);
    const int h = static_cast<int>((pix / W) % H);
    const int n = static_cast<int>(pix / (static_cast<long long>(W) * H));
    const float v = (layout == kNHWCSplit) ? join16(inh[i], inh[plane + i]) : in[i];
    out[((static_cast<long long>(n) * C + c) * H + h) * W + w] = v;
  }
}

int unpack_to_nchw(const void* in, long long plane, int in_layout, float* out, int N, int C, int H, int W,
                   cudaStream_t s) {
  MF_REQUIRE(in_layout == kNHWCRaw || in_layout == kNHWCSplit, "unpack_to_nchw expects an NHWC source");
  const long long total = static_cast<long long>(N) * H * W * C;
  if (total == 0) return 0;
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 16));
  unpack_to_nchw_kernel<<<blocks, 256, 0, s>>>(in, plane, in_layout, out, N, C, H, W);
  MF_CUDA_OK(cudaGetLastError());
  return 0;
}

// =================================================================================================
// Weight re-layout (once, at load time)
// =================================================================================================
// max |w| of a tensor -> *out (device float), used to choose the power-of-two weight pre-scale
__global__ void absmax_kernel(const float* __restrict__ w, long long n, float* __restrict__ out) {
  float m = 0.f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    m = fmaxf(m, fabsf(w[i]));
  for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));  // m >= 0: int order == float order
}
// scale = 2^S with max|w| * 2^S in [2^12, 2^13): the lo parts (~2^-11 of the hi parts) stay in fp16's normal range and
// the largest scaled weight is far below 65504.  scales[0] = 2^S, scales[1] = 2^-S (read by conv_tc's drain FMA).
__global__ void weight_scale_kernel(const float* __restrict__ absmax, float* __restrict__ scales) {
  const float m = *absmax;
  int e = 0;
  if (m > 0.f && isfinite(m)) {
    frexpf(m, &e);  // m = f * 2^e, f in [0.5, 1)
    e = 13 - e;
  }
  e = max(-24, min(24, e));
  scales[0] = ldexpf(1.0f, e);
  scales[1] = ldexpf(1.0f, -e);
}

__global__ void prep_weight_tc_kernel(const float* __restrict__ w, __half* __restrict__ out,
                                      const float* __restrict__ scales, int Cout, int Cin, int kh, int kw, int CinPad) {
  const long long K = static_cast<long long>(kh) * kw * CinPad;
  const long long total = K * Cout;
  const float sc = scales[0];
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long k = i % K;
    const int o = static_cast<int>(i / K);
    // K = ((c / 64) * taps + tap) * 64 + c % 64 : 64-channel slab major, tap minor (matches conv_tc's K loop)
    const int taps = kh * kw;
    const int cb = static_cast<int>(k / (taps * 64));
    const int rem = static_cast<int>(k % (taps * 64));
    const int tap = rem / 64;
    const int c = cb * 64 + rem % 64;
    const float v = c < Cin ? w[(static_cast<long long>(o) * Cin + c) * kh * kw + tap] * sc : 0.f;
    split16(v, out[i], out[total + i]);
  }
}

static int weight_scales(const float* w, long long n, float* scales, cudaStream_t s) {
  // scales: device float[4]: [0] = 2^S, [1] = 2^-S, [2] = scratch for the max
  MF_CUDA_OK(cudaMemsetAsync(scales + 2, 0, sizeof(float), s));
  const int blocks = static_cast<int>(std::min<long long>((n + 255) / 256, 148 * 8));
  absmax_kernel<<<blocks, 256, 0, s>>>(w, n, scales + 2);
  weight_scale_kernel<<<1, 1, 0, s>>>(scales + 2, scales);
  MF_CUDA_OK(cudaGetLastError());
  return 0;
}

int prep_weight_tc(const float* w_oihw, __half* out, float* scales, int Cout, int Cin, int kh, int kw, cudaStream_t s,
                   int cin_pad) {
  if (cin_pad <= 0) cin_pad = Cin;
  MF_REQUIRE(cin_pad % 64 == 0 && cin_pad >= Cin, "prep_weight_tc needs the (padded) channel count to be a multiple of 64");
  int rc = weight_scales(w_oihw, static_cast<long long>(Cout) * Cin * kh * kw, scales, s);
  if (rc) return rc;
  const long long total = static_cast<long long>(Cout) * cin_pad * kh * kw;
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 16));
  prep_weight_tc_kernel<<<blocks, 256, 0, s>>>(w_oihw, out, scales, Cout, Cin, kh, kw, cin_pad);
  MF_CUDA_OK(cudaGetLastError());
  return 0;
}

// BasicUp fold: conv3x3(nearest_x2(x)) == four 2x2 convolutions on x, one per output parity (a, b):
//   rows: a=0 -> {h-1: w[r=0], h: w[1]+w[2]}   a=1 -> {h: w[0]+w[1], h+1: w[2]}   (same for columns)
// out[2][4*Cout][4*Cin] in conv_tc's K order (32-channel slab major, tap (u,v) minor).
__global__ void prep_weight_up_tc_kernel(const float* __restrict__ w, __half* __restrict__ out,
                                         const float* __restrict__ scales, int Cout, int Cin) {
  const long long K = 4LL * Cin;
  const long long total = K * Cout * 4;
  const float sc = scales[0];
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long k = i % K;
    const long long row = i / K;
    const int o = static_cast<int>(row % Cout);
    const int phase = static_cast<int>(row / Cout);
    const int a = phase >> 1, b = phase & 1;
    const int cb = static_cast<int>(k / 256);
    const int rem = static_cast<int>(k % 256);
    const int tap = rem / 64, c = cb * 64 + rem % 64;
    const int u = tap >> 1, v = tap & 1;
    const float* wk = w + (static_cast<long long>(o) * Cin + c) * 9;
    // rows/cols of the 3x3 kernel that land on low-res offset u (resp. v) for output parity a (resp. b)
    const int r_lo = (a == 0) ? (u == 0 ? 0 : 1) : (u == 0 ? 0 : 2);
    const int r_hi = (a == 0) ? (u == 0 ? 0 : 2) : (u == 0 ? 1 : 2);
    const int s_lo = (b == 0) ? (v == 0 ? 0 : 1) : (v == 0 ? 0 : 2);
    const int s_hi = (b == 0) ? (v == 0 ? 0 : 2) : (v == 0 ? 1 : 2);
    float acc = 0.f;
    for (int r = r_lo; r <= r_hi; ++r)
      for (int sx = s_lo; sx <= s_hi; ++sx) acc += wk[r * 3 + sx];
    split16(acc * sc, out[i], out[total + i]);
  }
}

int prep_weight_up_tc(const float* w_oihw, __half* out, float* scales, int Cout, int Cin, cudaStream_t s) {
  MF_REQUIRE(Cin % 64 == 0, "prep_weight_up_tc needs Cin % 64 == 0");
  const long long total = 16LL * Cout * Cin;
  // the pre-summed phase taps are at most 4x the largest 3x3 weight: the scale chosen from max|w| leaves 2^13 * 4 < 65504
  int rc = weight_scales(w_oihw, 9LL * Cout * Cin, scales, s);
  if (rc) return rc;
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 16));
  prep_weight_up_tc_kernel<<<blocks, 256, 0, s>>>(w_oihw, out, scales, Cout, Cin);
  MF_CUDA_OK(cudaGetLastError());
  return 0;
}

__global__ void prep_weight_simt_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin, int kh,
                                        int kw) {
  const long long K = static_cast<long long>(kh) * kw * Cin;
  const long long total = K * Cout;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int o = static_cast<int>(i % Cout);
    const long long k = i / Cout;
    const int c = static_cast<int>(k % Cin);
    const int tap = static_cast<int>(k / Cin);
    out[i] = w[(static_cast<long long>(o) * Cin + c) * kh * kw + tap];
  }
}

int prep_weight_simt(const float* w_oihw, float* out, int Cout, int Cin, int kh, int kw, cudaStream_t s) {
  const long long total = static_cast<long long>(Cout) * Cin * kh * kw;
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 16));
  prep_weight_simt_kernel<<<blocks, 256, 0, s>>>(w_oihw, out, Cout, Cin, kh, kw);
  MF_CUDA_OK(cudaGetLastError());
  return 0;
}

// =================================================================================================
// Exact fp32 convolution on CUDA cores (implicit GEMM, smem tiled, 4x4 register micro-tile)
//   reference call sites: unet2.py:67 (in_conv, Cin=8), conv_blocks.py:43-52 (BasicDown, stride 2),
//   unet2.py:213 (UnetOutBlock 1x1 -> 8), latent_embedders.py:719 (inc_dec stem, Cin=8), :743 (outc 64->3)
// =================================================================================================
constexpr int kSimtBK = 16;

template <int TM, int TN>
__global__ void __launch_bounds__(256)
conv_simt_kernel(const ConvSimtDesc d, int Hout, int Wout) {
  static_assert(TM * TN == 4096, "256 threads x 16 outputs");
  __shared__ float As[kSimtBK][TM + 4];
  __shared__ float Bs[kSimtBK][TN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % (TN / 4);
  const int ty = tid / (TN / 4);
  const long long M = static_cast<long long>(d.N) * Hout * Wout;
  const long long m0 = static_cast<long long>(blockIdx.x) * TM;
  const int n0 = blockIdx.y * TN;
  const int K = d.ksize * d.ksize * d.Cin;
  const int pad = d.ksize / 2;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += kSimtBK) {
    // ---- A tile: TM pixels x BK (tap, channel) entries, gathered with zero padding
    for (int e = tid; e < TM * kSimtBK; e += 256) {
      int kk, mm;
      if (d.in_layout == kNCHW) { mm = e % TM; kk = e / TM; } else { kk = e % kSimtBK; mm = e / kSimtBK; }
      const int k = k0 + kk;
      const long long m = m0 + mm;
      float v = 0.f;
      if (k < K && m < M) {
        const int c = k % d.Cin;
        const int tap = k / d.Cin;
        const int r = tap / d.ksize, sx = tap % d.ksize;
        const int wo = static_cast<int>(m % Wout);
        const int ho = static_cast<int>((m / Wout) % Hout);
        const int n = static_cast<int>(m / (static_cast<long long>(Wout) * Hout));
        const int hi = ho * d.stride + r - pad;
        const int wi = wo * d.stride + sx - pad;
        if (hi >= 0 && hi < d.Hin && wi >= 0 && wi < d.Win) {
          if (d.in_layout == kNCHW) {
            v = reinterpret_cast<const float*>(d.in)[((static_cast<long long>(n) * d.Cin + c) * d.Hin + hi) * d.Win + wi];
          } else {
            // channel-concatenated sources (torch.cat of unet2.py:259): c < C0 reads source 0, else source 1
            const int C0 = d.Cin - d.C1;
            const bool second = c >= C0;
            const void* src = second ? d.in1 : d.in;
            const long long plane = second ? d.in1_plane : d.in_plane;
            const int cs = second ? d.C1 : C0, cc = second ? c - C0 : c;
            const long long off = ((static_cast<long long>(n) * d.Hin + hi) * d.Win + wi) * cs + cc;
            if (d.in_layout == kNHWCSplit) {
              const __half* ih = reinterpret_cast<const __half*>(src);
              v = join16(ih[off], ih[plane + off]);
            } else {
              v = reinterpret_cast<const float*>(src)[off];
            }
          }
        }
      }
      As[kk][mm] = v;
    }
    // ---- B tile: BK x TN weights ([K][Cout], Cout contiguous)
    for (int e = tid; e < TN * kSimtBK; e += 256) {
      const int nn = e % TN, kk = e / TN;
      const int k = k0 + kk, o = n0 + nn;
      Bs[kk][nn] = (k < K && o < d.Cout) ? d.w_kc[static_cast<long long>(k) * d.Cout + o] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kSimtBK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const int wo = static_cast<int>(m % Wout);
    const int ho = static_cast<int>((m / Wout) % Hout);
    const int n = static_cast<int>(m / (static_cast<long long>(Wout) * Hout));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int o = n0 + tx * 4 + j;
      if (o >= d.Cout) continue;
      float v = acc[i][j] + (d.bias ? d.bias[o] : 0.f);
      if (d.out_layout == kNCHW) {
        reinterpret_cast<float*>(d.out)[((static_cast<long long>(n) * d.Cout + o) * Hout + ho) * Wout + wo] = v;
      } else {
        const long long off = m * d.Cout + o;
        if (d.out_layout == kNHWCSplit) {
          __half* oh = reinterpret_cast<__half*>(d.out);
          split16(v, oh[off], oh[d.out_plane + off]);
        } else {
          reinterpret_cast<float*>(d.out)[off] = v;
        }
      }
    }
  }
}

int conv_simt(const ConvSimtDesc& d, cudaStream_t s) {
  MF_REQUIRE(d.ksize == 1 || d.ksize == 3, "conv_simt supports 1x1 and 3x3");
  MF_REQUIRE(d.stride == 1 || d.stride == 2, "conv_simt supports stride 1 and 2");
  MF_REQUIRE(d.C1 == 0 || (d.in1 != nullptr && d.in_layout != kNCHW && d.C1 < d.Cin),
             "conv_simt: the second source must be an NHWC tensor");
  const int pad = d.ksize / 2;
  const int Hout = (d.Hin + 2 * pad - d.ksize) / d.stride + 1;
  const int Wout = (d.Win + 2 * pad - d.ksize) / d.stride + 1;
  const long long M = static_cast<long long>(d.N) * Hout * Wout;
  if (M == 0) return 0;
  if (d.Cout <= 8) {
    dim3 grid(ceil_div_ll(M, 512), ceil_div_ll(d.Cout, 8));
    conv_simt_kernel<512, 8><<<grid, 256, 0, s>>>(d, Hout, Wout);
  } else if (d.Cout <= 16) {
    dim3 grid(ceil_div_ll(M, 256), ceil_div_ll(d.Cout, 16));
    conv_simt_kernel<256, 16><<<grid, 256, 0, s>>>(d, Hout, Wout);
  } else {
    dim3 grid(ceil_div_ll(M, 64), ceil_div_ll(d.Cout, 64));
    conv_simt_kernel<64, 64><<<grid, 256, 0, s>>>(d, Hout, Wout);
  }
  MF_CUDA_OK(cudaGetLastError());
  return 0;
}

// =================================================================================================
// GroupNorm (reference: conv_blocks.py:177,187 -> nn.GroupNorm(eps=1e-5, affine), biased variance)
// =================================================================================================
// One block per (n, 8-channel slab): sum / sumsq over all pixels.  Only used after SIMT convs.
__global__ void gn_partial_from_raw_kernel(const void* __restrict__ raw_v, long long plane, float* __restrict__ partial,
                                           int HW, int C) {
  const int n = blockIdx.y, g8 = blockIdx.x;
  const long long boff = static_cast<long long>(n) * HW * C + g8 * 8;
  float s = 0.f, ss = 0.f;
  for (int p = threadIdx.x; p < HW; p += blockDim.x) {
    float4 a, b;
    if (plane != 0) {  // split tensor: x = hi + lo (fp16 planes)
      const __half* h = reinterpret_cast<const __half*>(raw_v) + boff + static_cast<long long>(p) * C;
      a = ld_join4(h, h + plane);
      b = ld_join4(h + 4, h + plane + 4);
    } else {
      const float* base = reinterpret_cast<const float*>(raw_v) + boff + static_cast<long long>(p) * C;
      a = *reinterpret_cast<const float4*>(base);
      b = *reinterpret_cast<const float4*>(base + 4);
    }
    s += ((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w));
    ss += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w + b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
  }
  __shared__ float sh[2][32];
  for (int off = 16; off > 0; off >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, off);
    ss += __shfl_xor_sync(0xffffffffu, ss, off);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sh[0][warp] = s; sh[1][warp] = ss; }
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    s = lane < nw ? sh[0][lane] : 0.f;
    ss = lane < nw ? sh[1][lane] : 0.f;
    for (int off = 16; off > 0; off >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, off);
      ss += __shfl_xor_sync(0xffffffffu, ss, off);
    }
    if (lane == 0) {
      float* dst = partial + (static_cast<long long>(n) * (C / 8) + g8) * 2;
      dst[0] = s;
      dst[1] = ss;
    }
  }
}

int gn_partial_from_raw(const void* raw, float* partial, int N, int HW, int C, cudaStream_t s, long long plane) {
  MF_REQUIRE(C % 8 == 0, "GroupNorm partial sums need C % 8 == 0");
  dim3 grid(C / 8, N);
  gn_partial_from_raw_kernel<<<grid, 256, 0, s>>>(raw, plane, partial, HW, C);
  MF_CUDA_OK(cudaGetLastError());
  return 0;
}

// One warp per (n, group): reduce chunks x (cpg/8) partial pairs in double for a stable variance.
__global__ void gn_finalize_kernel(const float* __restrict__ partial, float* __restrict__ mean_rstd, int N, int chunks,
                                   int C, int G, int HW, float eps) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (gw >= N * G) return;
  const int n = gw / G, g = gw % G;
  const int cpg = C / G, sub = cpg / 8, c8 = C / 8;
  double s = 0.0, ss = 0.0;
  const int items = chunks * sub;
  for (int it = lane; it < items; it += 32) {
    const int ch = it / sub, j = it % sub;
    const float* src = partial + ((static_cast<long long>(n) * chunks + ch) * c8 + g * sub + j) * 2;
    s += static_cast<double>(src[0]);
    ss += static_cast<double>(src[1]);
  }
  for (int off = 16; off > 0; off >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, off);
    ss += __shfl_xor_sync(0xffffffffu, ss, off);
  }
  if (lane == 0) {
    const double cnt = static_cast<double>(cpg) * HW;
    const double mean = s / cnt;
    double var = ss / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    mean_rstd[(static_cast<long long>(n) * G + g) * 2 + 0] = static_cast<float>(mean);
    mean_rstd[(static_cast<long long>(n) * G + g) * 2 + 1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  }
}

int gn_finalize(const float* partial, float* mean_rstd, int N, int chunks, int C, int G, int HW, float eps,
                cudaStream_t s) {
  MF_REQUIRE(C % G == 0 && (C / G) % 8 == 0, "GroupNorm groups must span a multiple of 8 channels");
  const int warps = N * G;
  gn_finalize_kernel<<<ceil_div_ll(warps, 8), 256, 0, s>>>(partial, mean_rstd, N, chunks, C, G, HW, eps);
  MF_CUDA_OK(cudaGetLastError());
  return 0;
}

// y = swish(gn(raw)) + residual (+ emb)   -> split planes
//   reference: conv_blocks.py:184-192 (conv -> norm -> drop(p=0) -> act), :236-240 (+ residual), :362-363 (+= emb)
// Thread layout: a thread owns ONE channel quad for its whole life (gamma / beta / group index stay in registers) and
// walks the pixels; 256 / (C/4) pixels per block iteration.  Pure streaming: 4 B raw + 4 B residual in, 4 B out per element.
__global__ void __launch_bounds__(256) gn_apply_kernel(const GnApplyDesc d) {
  const int c4n = d.C / 4;
  const int tpp = c4n < 256 ? c4n : 256;          // threads per pixel
  const int ppb = 256 / tpp;                      // pixels per block iteration
  const int cq = threadIdx.x % tpp;
  const int psub = threadIdx.x / tpp;
  if (psub >= ppb) return;                        // C/4 does not divide 256: the remainder threads idle
  const int cpg = d.C / d.G;
  const long long npix = static_cast<long long>(d.N) * d.HW;
  for (int c = cq * 4; c < d.C; c += tpp * 4) {   // one pass unless C > 1024
    const float4 ga = __ldg(reinterpret_cast<const float4*>(d.gamma + c));
    const float4 be = __ldg(reinterpret_cast<const float4*>(d.beta + c));
    const int g = c / cpg;
    for (long long pix = static_cast<long long>(blockIdx.x) * ppb + psub; pix < npix;
         pix += static_cast<long long>(gridDim.x) * ppb) {
      const int n = static_cast<int>(pix / d.HW);
      const long long off = pix * d.C + c;
      float4 x;
      if (d.raw_plane != 0) {
        const __half* xh = reinterpret_cast<const __half*>(d.raw) + off;
        x = ld_join4(xh, xh + d.raw_plane);
      } else {
        x = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(d.raw) + off);
      }
      const float2 mr = __ldg(reinterpret_cast<const float2*>(d.mean_rstd + (static_cast<long long>(n) * d.G + g) * 2));
      float y[4] = {(x.x - mr.x) * mr.y * ga.x + be.x, (x.y - mr.x) * mr.y * ga.y + be.y,
                    (x.z - mr.x) * mr.y * ga.z + be.z, (x.w - mr.x) * mr.y * ga.w + be.w};
      if (d.act != 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) y[j] = swish(y[j]);
      }
      if (d.res_kind == kResSplit) {
        const __half* rh = reinterpret_cast<const __half*>(d.res) + off;
        const float4 r = ld_join4(rh, rh + d.res_plane);
        y[0] += r.x; y[1] += r.y; y[2] += r.z; y[3] += r.w;
      } else if (d.res_kind == kResRaw) {
        const float4 r = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(d.res) + off);
        y[0] += r.x; y[1] += r.y; y[2] += r.z; y[3] += r.w;
      }
      if (d.emb != nullptr) {
        const long long er = d.emb_index ? d.emb_index[n] : static_cast<long long>(n);
        const float4 e = __ldg(reinterpret_cast<const float4*>(d.emb + er * d.emb_stride + c));
        y[0] += e.x; y[1] += e.y; y[2] += e.z; y[3] += e.w;
      }
      st_split4(d.out + off, d.out + d.out_plane + off, make_float4(y[0], y[1], y[2], y[3]));
    }
  }
}

// flat variant: every thread-iteration handles one channel quad of one pixel (index math per element, maximal parallelism)
__global__ void __launch_bounds__(256) gn_apply_flat_kernel(const GnApplyDesc d) {
  const int c4n = d.C / 4;
  const unsigned total = static_cast<unsigned>(static_cast<long long>(d.N) * d.HW * c4n);
  const int cpg = d.C / d.G;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned pix = i / c4n;
    const int c = static_cast<int>(i - pix * c4n) * 4;
    const int n = static_cast<int>(pix / d.HW);
    const long long off = static_cast<long long>(pix) * d.C + c;
    float4 x;
    if (d.raw_plane != 0) {
      const __half* xh = reinterpret_cast<const __half*>(d.raw) + off;
      x = ld_join4(xh, xh + d.raw_plane);
    } else {
      x = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(d.raw) + off);
    }
    const float2 mr = __ldg(reinterpret_cast<const float2*>(d.mean_rstd + (static_cast<long long>(n) * d.G + c / cpg) * 2));
    const float4 ga = __ldg(reinterpret_cast<const float4*>(d.gamma + c));
    const float4 be = __ldg(reinterpret_cast<const float4*>(d.beta + c));
    float y[4] = {(x.x - mr.x) * mr.y * ga.x + be.x, (x.y - mr.x) * mr.y * ga.y + be.y,
                  (x.z - mr.x) * mr.y * ga.z + be.z, (x.w - mr.x) * mr.y * ga.w + be.w};
    if (d.act != 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) y[j] = swish(y[j]);
    }
    if (d.res_kind == kResSplit) {
      const __half* rh = reinterpret_cast<const __half*>(d.res) + off;
      const float4 r = ld_join4(rh, rh + d.res_plane);
      y[0] += r.x; y[1] += r.y; y[2] += r.z; y[3] += r.w;
    } else if (d.res_kind == kResRaw) {
      const float4 r = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(d.res) + off);
      y[0] += r.x; y[1] += r.y; y[2] += r.z; y[3] += r.w;
    }
    if (d.emb != nullptr) {
      const long long er = d.emb_index ? d.emb_index[n] : static_cast<long long>(n);
      const float4 e = __ldg(reinterpret_cast<const float4*>(d.emb + er * d.emb_stride + c));
      y[0] += e.x; y[1] += e.y; y[2] += e.z; y[3] += e.w;
    }
    st_split4(d.out + off, d.out + d.out_plane + off, make_float4(y[0], y[1], y[2], y[3]));
  }
}

// fused variant (default): one block = a pixel range of ONE sample.
//   prologue: the block reduces that sample's partial statistics to (mean, rstd) per group itself (fp64 sums, L lanes per
//             group and 32 / L groups per warp pass; rstd by fp32 rsqrt + Newton: equal to gn_finalize_kernel's to ~1 ulp),
//             so the separate finalize launch disappears;
//   body:     a thread handles 8 consecutive channels of a pixel (two 16-byte loads, one 16-byte store per plane),
//             2 (with a residual) or 4 pixels in flight per iteration.
__device__ __forceinline__ void ld_raw8(const GnApplyDesc& d, long long off, float (&x)[8]) {
  if (d.raw_plane != 0) {
    const __half* xh = reinterpret_cast<const __half*>(d.raw) + off;
    const uint4 uh = *reinterpret_cast<const uint4*>(xh);
    const uint4 ul = *reinterpret_cast<const uint4*>(xh + d.raw_plane);
    const __half2* h = reinterpret_cast<const __half2*>(&uh);
    const __half2* l = reinterpret_cast<const __half2*>(&ul);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 a = __half22float2(h[j]), b = __half22float2(l[j]);
      x[2 * j] = a.x + b.x; x[2 * j + 1] = a.y + b.y;
    }
  } else {
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(d.raw) + off);
    const float4 a = p[0], b = p[1];
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
  }
}

// 8 values of the residual at `off` (split planes or raw fp32)
__device__ __forceinline__ void ld_res8(const GnApplyDesc& d, long long off, float (&r)[8]) {
  if (d.res_kind == kResSplit) {
    const __half* rh = reinterpret_cast<const __half*>(d.res) + off;
    const uint4 uh = *reinterpret_cast<const uint4*>(rh);
    const uint4 ul = *reinterpret_cast<const uint4*>(rh + d.res_plane);
    const __half2* h = reinterpret_cast<const __half2*>(&uh);
    const __half2* l = reinterpret_cast<const __half2*>(&ul);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 a = __half22float2(h[j]), b = __half22float2(l[j]);
      r[2 * j] = a.x + b.x; r[2 * j + 1] = a.y + b.y;
    }
  } else {
    const float4* rp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(d.res) + off);
    const float4 a = rp[0], b = rp[1];
    r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
  }
}

// HC: compile-time bound on the folded head's outputs (0 = no head, 4, 8); RES: a residual is added.
// U pixels are in flight per thread and iteration (all their loads are issued before the first use): 4 without a residual,
// 2 with one — 128 bytes of loads per thread either way.  Round 2 ncu (profiles/r02_ncu_unet_step.md): the one-pixel loop ran at
// 64 registers = 4 blocks per SM with 32-64 bytes in flight per thread and a grid sized for 8 blocks per SM (2.05 waves):
// 38-57 % of the HBM peak; the folded-head instantiation (110 registers, 2 blocks per SM) at 26 %.
template <int HC, bool RES>
__global__ void __launch_bounds__(256, 3) gn_apply_fused_kernel(const GnApplyDesc d) {
  constexpr int U = RES ? 2 : 4;
  __shared__ float s_mean[128], s_rstd[128];
  __shared__ __align__(16) float s_hw[HC > 0 ? HC * 256 : 4];   // folded head weights [o][c] (C <= 256)
  pdl_launch_dependents();
  // parameters (gamma, beta, head weights) do not depend on the previous kernel: requested before griddepcontrol.wait, so
  // their latency overlaps the tail of the convolution that produces this kernel's input
  const int c8n = d.C / 8;
  const int c = (threadIdx.x % c8n) * 8;
  const float4 ga0 = __ldg(reinterpret_cast<const float4*>(d.gamma + c)), ga1 = __ldg(reinterpret_cast<const float4*>(d.gamma + c + 4));
  const float4 be0 = __ldg(reinterpret_cast<const float4*>(d.beta + c)), be1 = __ldg(reinterpret_cast<const float4*>(d.beta + c + 4));
  if (HC > 0) {
    for (int i = threadIdx.x; i < d.head_cout * d.C; i += 256) s_hw[(i / d.C) * 256 + i % d.C] = __ldg(d.head_w + i);
  }
  pdl_wait();
  const int n = blockIdx.y;
  const int cpg = d.C / d.G;
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sub = cpg / 8, c8 = d.C / 8, items = d.chunks * sub;
    const double inv_cnt = 1.0 / (static_cast<double>(cpg) * d.HW);
    // L lanes per group (the partial sums of a group are few: 8 at the UNet's 32x32 level), 32 / L groups per warp pass
    int L = 1;
    while (L < items && L < 32) L <<= 1;
    const int gpp = 32 / L;
    for (int g0 = warp * gpp; g0 < d.G; g0 += 8 * gpp) {
      const int g = g0 + lane / L, it0 = lane % L;
      double s = 0.0, ss = 0.0;
      if (g < d.G) {
        for (int it = it0; it < items; it += L) {
          const int ch = it / sub, j = it % sub;
          const float2 v = __ldg(reinterpret_cast<const float2*>(
              d.partial + ((static_cast<long long>(n) * d.chunks + ch) * c8 + g * sub + j) * 2));
          s += static_cast<double>(v.x);
          ss += static_cast<double>(v.y);
        }
      }
      for (int off = L >> 1; off > 0; off >>= 1) {        // stays inside the group's L-lane segment
        s += __shfl_xor_sync(0xffffffffu, s, off);
        ss += __shfl_xor_sync(0xffffffffu, ss, off);
      }
      if (it0 == 0 && g < d.G) {
        // fp64 only where the cancellation is (E[x^2] - mean^2); the reciprocal square root is fp32 rsqrt + one Newton
        // step (<= 1 ulp): three fp64 divisions and an fp64 sqrt per group, in every block, were microseconds of
        // dependent fp64 instructions in front of a 20-35 us kernel
        const double mean = s * inv_cnt;
        double var = fma(ss, inv_cnt, -mean * mean);
        if (var < 0.0) var = 0.0;
        const float v32 = static_cast<float>(var + static_cast<double>(d.eps));
        float r = rsqrtf(v32);
        r = r * fmaf(-0.5f * v32, r * r, 1.5f);
        s_mean[g] = static_cast<float>(mean);
        s_rstd[g] = r;
      }
    }
  }
  __syncthreads();
  // A thread keeps ONE channel octet for its whole life (256 % (C/8) == 0 is checked by the host): the affine
  // coefficients a = rstd*gamma, b = beta (+ embedding) stay in registers and the loop body is pure streaming.
  const int ppb = 256 / c8n;                                // pixels per block and sub-iteration
  const int psub = threadIdx.x / c8n;
  const int g = c / cpg;
  const float mean = s_mean[g], rstd = s_rstd[g];
  float ca[8], cb[8], ce[8];
  {
    const float ga[8] = {ga0.x, ga0.y, ga0.z, ga0.w, ga1.x, ga1.y, ga1.z, ga1.w};
    const float be[8] = {be0.x, be0.y, be0.z, be0.w, be1.x, be1.y, be1.z, be1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) { ca[j] = rstd * ga[j]; cb[j] = be[j]; ce[j] = 0.f; }
    if (d.emb != nullptr) {
      const long long erow = (d.emb_index ? d.emb_index[n] : static_cast<long long>(n)) * d.emb_stride;
      const float4 e0 = __ldg(reinterpret_cast<const float4*>(d.emb + erow + c));
      const float4 e1 = __ldg(reinterpret_cast<const float4*>(d.emb + erow + c + 4));
      ce[0] = e0.x; ce[1] = e0.y; ce[2] = e0.z; ce[3] = e0.w; ce[4] = e1.x; ce[5] = e1.y; ce[6] = e1.z; ce[7] = e1.w;
    }
  }
  const long long base = static_cast<long long>(n) * d.HW * d.C + c;
  const int step = static_cast<int>(gridDim.x) * ppb;
  bool clamped = false;
  // block-uniform trip count (the head reduction shuffles across the C/8 lanes of a pixel)
  for (int pix0 = blockIdx.x * ppb; pix0 < d.HW; pix0 += U * step) {
    float x[U][8], r[RES ? U : 1][8];
    bool live[U];
    long long off[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pix = pix0 + u * step + psub;
      live[u] = pix < d.HW;
      off[u] = base + static_cast<long long>(live[u] ? pix : 0) * d.C;
#pragma unroll
      for (int j = 0; j < 8; ++j) x[u][j] = 0.f;
      if (live[u]) ld_raw8(d, off[u], x[u]);
    }
    if (RES) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[u][j] = 0.f;
        if (live[u]) ld_res8(d, off[u], r[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float y[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        y[j] = fmaf(x[u][j] - mean, ca[j], cb[j]);
        if (d.act != 0) y[j] = swish_stream(y[j]);
        if (RES) y[j] += r[u][j];
        y[j] += ce[j];
      }
      if (HC > 0) {
#pragma unroll
        for (int o = 0; o < HC; ++o) {
          if (o < d.head_cout) {                      // block-uniform
            const float4 w0 = *reinterpret_cast<const float4*>(&s_hw[o * 256 + c]);
            const float4 w1 = *reinterpret_cast<const float4*>(&s_hw[o * 256 + c + 4]);
            float acc = 0.f;
            acc = fmaf(y[0], w0.x, acc); acc = fmaf(y[1], w0.y, acc); acc = fmaf(y[2], w0.z, acc); acc = fmaf(y[3], w0.w, acc);
            acc = fmaf(y[4], w1.x, acc); acc = fmaf(y[5], w1.y, acc); acc = fmaf(y[6], w1.z, acc); acc = fmaf(y[7], w1.w, acc);
#pragma unroll
            for (int sh = 16; sh > 0; sh >>= 1)
              if (sh < c8n) acc += __shfl_xor_sync(0xffffffffu, acc, sh);     // block-uniform predicate, no runtime loop
            if (live[u] && c == 0) {                                          // c == 0: first of the pixel's C/8 lanes
              const int pix = pix0 + u * step + psub;
              const float v = acc + __ldg(d.head_b + o);
              if (d.head_out != nullptr) d.head_out[(static_cast<long long>(n) * d.head_cout + o) * d.HW + pix] = v;
              if (d.head_out_u8 != nullptr) {
                // scripts/helpers/sample_dataset.py:47-50: clip(-1,1) -> (x+1)/2*255 -> HWC -> astype(uint8) (truncation)
                const float c01 = fminf(fmaxf(v, -1.f), 1.f);
                const float v255 = __fmul_rn(__fmul_rn(__fadd_rn(c01, 1.f), 0.5f), 255.f);
                d.head_out_u8[(static_cast<long long>(n) * d.HW + pix) * d.head_cout + o] = static_cast<unsigned char>(v255);
              }
            }
          }
        }
      }
      if (d.out != nullptr && live[u]) {
        uint4 oh, ol;
        uint32_t* ph = reinterpret_cast<uint32_t*>(&oh);
        uint32_t* pl = reinterpret_cast<uint32_t*>(&ol);
#pragma unroll
        for (int j = 0; j < 4; ++j) split16x2_flag(y[2 * j], y[2 * j + 1], ph[j], pl[j], clamped);
        *reinterpret_cast<uint4*>(d.out + off[u]) = oh;
        *reinterpret_cast<uint4*>(d.out + d.out_plane + off[u]) = ol;
      }
    }
  }
  sat16_report(clamped);
}

// resident blocks per SM of an instantiation (asked once), and the grid: bps blocks per sample such that N * bps fills
// whole waves of the 148 x resident block slots as evenly as possible
template <int HC, bool RES>
static int gn_fused_launch(const GnApplyDesc& d, cudaLaunchConfig_t& cfg, int max_bps) {
  static int occ = 0, sms = 0;
  if (occ == 0) {
    MF_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gn_apply_fused_kernel<HC, RES>, 256, 0));
    int dev = 0;
    MF_CUDA_OK(cudaGetDevice(&dev));
    MF_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (occ < 1) occ = 1;
  }
  const long long slots = static_cast<long long>(sms) * occ;
  // cost of a grid in units of one loop iteration: waves x (prologue + iterations per block).  The prologue (statistics
  // finalisation, coefficient loads) is worth ~3 iterations, so a small sample wants FEW, fat blocks in ONE wave (8x8 level of
  // the UNet: 6 blocks per sample instead of 13: 18 -> 12 us per launch), a large one as many as fill whole waves.
  const int upi = (RES ? 2 : 4) * (256 / (d.C / 8));
  int best = 1;
  double best_cost = 1e30;
  for (int bps = 1; bps <= max_bps; ++bps) {
    const long long blocks = static_cast<long long>(d.N) * bps;
    const long long waves = (blocks + slots - 1) / slots;
    if (waves > 4 && bps > 1) break;
    const long long iters = (d.HW + static_cast<long long>(upi) * bps - 1) / (static_cast<long long>(upi) * bps);
    const double cost = static_cast<double>(waves) * (3.0 + static_cast<double>(iters));
    if (cost < best_cost * 0.98) { best_cost = cost; best = bps; }
  }
  cfg.gridDim = dim3(best, d.N, 1);
  MF_CUDA_OK(cudaLaunchKernelEx(&cfg, gn_apply_fused_kernel<HC, RES>, d));
  return 0;
}

int g_gn_variant = 3;
int g_pdl = 1;
int gn_apply(const GnApplyDesc& d, cudaStream_t s) {
  MF_REQUIRE(d.C % 4 == 0 && d.C % d.G == 0 && (d.C / d.G) % 4 == 0, "gn_apply channel constraints");
  MF_REQUIRE(d.emb == nullptr || d.emb_stride % 4 == 0, "emb rows must be float4 aligned");
  const long long npix = static_cast<long long>(d.N) * d.HW;
  if (npix == 0) return 0;
  if (d.partial != nullptr) {
    MF_REQUIRE(d.C % 8 == 0 && (d.C / d.G) % 8 == 0 && d.G <= 128 && d.chunks >= 1 && 256 % (d.C / 8) == 0,
               "fused gn_apply: C/G % 8 == 0, G <= 128, C/8 divides 256");
    MF_REQUIRE(d.N <= 65535, "fused gn_apply: batch too large");
    MF_REQUIRE(d.head_cout == 0 || (d.head_cout <= 8 && d.C / 8 <= 32 && ((d.C / 8) & (d.C / 8 - 1)) == 0 && d.head_w),
               "folded head: <= 8 outputs, C/8 a power of two <= 32");
    MF_REQUIRE(d.out != nullptr || d.head_cout > 0, "gn_apply without an output");
    const int ppb = 256 / (d.C / 8);
    const bool res = d.res != nullptr && d.res_kind != kResNone;
    const int upi = (res ? 2 : 4) * ppb;                     // pixels per block iteration
    const int max_bps = std::max(1, (d.HW + upi - 1) / upi);
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(256, 1, 1);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_pdl ? 1 : 0;
    if (d.head_cout == 0) return res ? gn_fused_launch<0, true>(d, cfg, max_bps) : gn_fused_launch<0, false>(d, cfg, max_bps);
    MF_REQUIRE(d.C <= 256, "folded head: at most 256 channels");
    if (d.head_cout <= 4) return res ? gn_fused_launch<4, true>(d, cfg, max_bps) : gn_fused_launch<4, false>(d, cfg, max_bps);
    return res ? gn_fused_launch<8, true>(d, cfg, max_bps) : gn_fused_launch<8, false>(d, cfg, max_bps);
  }
  MF_REQUIRE(d.head_cout == 0, "the folded head exists in the fused gn_apply variant only");
  const int c4n = d.C / 4;
  const int ppb = c4n < 256 ? 256 / c4n : 1;
  const long long total = npix * c4n;
  if (g_gn_variant == 1 || total >= (1LL << 32)) {
    const int blocks = static_cast<int>(std::min<long long>((npix + ppb - 1) / ppb, 148 * 16));
    gn_apply_kernel<<<blocks, 256, 0, s>>>(d);
  } else if (g_gn_variant == 2) {
    gn_apply_flat_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(d);   // one quad per thread
  } else {
    const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 32));
    gn_apply_flat_kernel<<<blocks, 256, 0, s>>>(d);
  }
  MF_CUDA_OK(cudaGetLastError());
  return 0;
}

// =================================================================================================
// nearest x2 upsample (split -> split)
// =================================================================================================
__global__ void upsample2x_split_kernel(const __half* __restrict__ in, long long in_plane, __half* __restrict__ out,
                                        long long out_plane, int N, int H, int W, int C) {
  const int c4n = C / 4;
  const int Ho = 2 * H, Wo = 2 * W;
  const long long total = static_cast<long long>(N) * Ho * Wo * c4n;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % c4n) * 4;
    const long long pix = i / c4n;
    const int wo = static_cast<int>(pix % Wo);
    const int ho = static_cast<int>((pix / Wo) % Ho);
    const int n = static_cast<int>(pix / (static_cast<long long>(Wo) * Ho));
    const long long src = ((static_cast<long long>(n) * H + (ho >> 1)) * W + (wo >> 1)) * C + c;
    const long long dst = pix * C + c;
    *reinterpret_cast<uint2*>(out + dst) = *reinterpret_cast<const uint2*>(in + src);
    *reinterpret_cast<uint2*>(out + out_plane + dst) = *reinterpret_cast<const uint2*>(in + in_plane + src);
  }
}

int upsample2x_split(const __half* in, long long in_plane, __half* out, long long out_plane, int N, int H, int W, int C,
                     cudaStream_t s) {
  MF_REQUIRE(C % 4 == 0, "upsample2x_split needs C % 4 == 0");
  const long long total = static_cast<long long>(N) * 4 * H * W * (C / 4);
  if (total == 0) return 0;
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 32));
  upsample2x_split_kernel<<<blocks, 256, 0, s>>>(in, in_plane, out, out_plane, N, H, W, C);
  MF_CUDA_OK(cudaGetLastError());
  return 0;
}

// =================================================================================================
// Embedding MLP (reference: time_embedder.py:15-28,67-75; cond_embedders.py:18-23; conv_blocks.py:340-350)
//   One warp per output feature j; the weight row stays in registers while the warp walks the batch.
// =================================================================================================
template <int KPL>  // K per lane (K = 32*KPL)
__global__ void linear_small_kernel(const LinearDesc d) {
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (j >= d.J) return;
  float w[KPL];
#pragma unroll
  for (int i = 0; i < KPL; ++i) w[i] = d.W[static_cast<long long>(j) * d.K + i * 32 + lane];
  const float bj = d.bias ? d.bias[j] : 0.f;
  for (int b = 0; b < d.B; ++b) {
    float acc = 0.f;
    if (d.in_mode == 0) {
      const float* x = d.in + static_cast<long long>(b) * d.K;
#pragma unroll
      for (int i = 0; i < KPL; ++i) acc = fmaf(w[i], x[i * 32 + lane], acc);
    } else {
      // sinusoidal position embedding: cat(sin(t*f), cos(t*f)), f given by the host table
      const float tf = d.t_float ? d.t_float[static_cast<long long>(b) * d.t_stride]
                                 : static_cast<float>(d.t[static_cast<long long>(b) * d.t_stride]);
      const int half = d.K / 2;
#pragma unroll
      for (int i = 0; i < KPL; ++i) {
        const int k = i * 32 + lane;
        const float a = tf * d.freqs[k < half ? k : k - half];
        acc = fmaf(w[i], k < half ? sinf(a) : cosf(a), acc);
      }
    }
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) {
      float v = acc + bj;
      if (d.add_table != nullptr) v += d.add_table[(d.add_idx ? d.add_idx[b] : static_cast<long long>(b)) * d.J + j];
      if (d.post == 1) v = swish(v);
      d.out[static_cast<long long>(b) * d.J + j] = v;
      if (d.out2 != nullptr) d.out2[static_cast<long long>(b) * d.J + j] = swish(v);
    }
  }
}

// any K (e.g. the 16-wide sinusoid of a 64-wide time embedding, time_embedder.py:62): one thread per output, sequential k
__global__ void linear_generic_kernel(const LinearDesc d) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<long long>(d.B) * d.J) return;
  const int b = static_cast<int>(i / d.J), j = static_cast<int>(i % d.J);
  float acc = 0.f;
  if (d.in_mode == 0) {
    for (int k = 0; k < d.K; ++k) acc = fmaf(d.W[static_cast<long long>(j) * d.K + k], d.in[static_cast<long long>(b) * d.K + k], acc);
  } else {
    const float tf = d.t_float ? d.t_float[static_cast<long long>(b) * d.t_stride]
                               : static_cast<float>(d.t[static_cast<long long>(b) * d.t_stride]);
    const int half = d.K / 2;
    for (int k = 0; k < d.K; ++k) {
      const float a = tf * d.freqs[k < half ? k : k - half];
      acc = fmaf(d.W[static_cast<long long>(j) * d.K + k], k < half ? sinf(a) : cosf(a), acc);
    }
  }
  float v = acc + (d.bias ? d.bias[j] : 0.f);
  if (d.add_table != nullptr) v += d.add_table[(d.add_idx ? d.add_idx[b] : static_cast<long long>(b)) * d.J + j];
  if (d.post == 1) v = swish(v);
  d.out[i] = v;
  if (d.out2 != nullptr) d.out2[i] = swish(v);
}

int linear_small(const LinearDesc& d, cudaStream_t s) {
  if (d.B == 0 || d.J == 0) return 0;
  const int kpl = d.K / 32;
  const bool fast = d.K % 32 == 0 && d.K <= 2048 && (kpl & (kpl - 1)) == 0;
  if (!fast) {
    const long long total = static_cast<long long>(d.B) * d.J;
    linear_generic_kernel<<<static_cast<unsigned>((total + 127) / 128), 128, 0, s>>>(d);
    MF_CUDA_OK(cudaGetLastError());
    return 0;
  }
  const int blocks = ceil_div_ll(static_cast<long long>(d.J) * 32, 256);
  switch (d.K / 32) {
#define MF_LIN_CASE(n) case n: linear_small_kernel<n><<<blocks, 256, 0, s>>>(d); break;
    MF_LIN_CASE(1) MF_LIN_CASE(2) MF_LIN_CASE(4) MF_LIN_CASE(8) MF_LIN_CASE(16) MF_LIN_CASE(32) MF_LIN_CASE(64)
#undef MF_LIN_CASE
    default:
      set_error("linear_small: unsupported K (need K/32 in {1,2,4,8,16,32,64})");
      return 2;
  }
  MF_CUDA_OK(cudaGetLastError());
  return 0;
}

// =================================================================================================
// Scheduler step (reference: gaussian_scheduler.py:80-124, diffusion_pipeline.py:240-244, :297-304)
// =================================================================================================
// one element of the reverse step; `pred` is the estimator output for this element (before guidance)
// estimate_x_t (gaussian_scheduler.py:61-77): t < 0 -> x_0, t >= T -> x_T, else sqrt(ac_t) x_0 + sqrt(1-ac_t) x_T
__device__ __forceinline__ float sched_x_t(const SchedStepDesc& d, float x0, float xT, long long t) {
  if (t < 0) return x0;
  if (t >= d.T) return xT;
  return __fadd_rn(__fmul_rn(d.sqrt_ac[t], x0), __fmul_rn(d.sqrt_1mac[t], xT));
}

// `pred` is element i of the estimator output in the [B, CHW] numbering of x_t; ip is the same element's position
// inside the (possibly channel-stacked) estimator tensors.
__device__ __forceinline__ void sched_element(const SchedStepDesc& d, long long i, int b, float pred) {
  const long long t = d.t[b];
  const long long ip = d.pred_bstride ? static_cast<long long>(b) * d.pred_bstride + (i - static_cast<long long>(b) * d.CHW) : i;
  if (d.pred_uncond != nullptr) {
    const float pu = d.pred_uncond[ip];
    pred = pu + d.guidance * (pred - pu);  // classifier-free guidance combine
  }
  const float xt = d.x_t[i];
  const float A = d.tab.sqrt_recip_ac[t], Bm = d.tab.sqrt_recipm1_ac[t];
  float x0, xT;
  if (d.objective_x0) {
    x0 = pred;
    if (d.clip_x0) x0 = fminf(fmaxf(x0, -1.f), 1.f);
    xT = __fdiv_rn(__fsub_rn(__fmul_rn(A, xt), x0), Bm);
  } else {
    xT = pred;
    x0 = __fsub_rn(__fmul_rn(A, xt), __fmul_rn(Bm, pred));
    if (d.clip_x0) x0 = fminf(fmaxf(x0, -1.f), 1.f);
  }
  float prior;
  if (d.cold) {
    // gaussian_scheduler.py:88-93: x_T re-estimated from the (always clamped, estimate_x_T default) x_0, then the
    // deterministic difference of two forward diffusions is removed from x_t; no random draw
    const float x0c = fminf(fmaxf(x0, -1.f), 1.f);
    const float xTe = __fdiv_rn(__fsub_rn(__fmul_rn(A, xt), x0c), Bm);
    const float est_t = sched_x_t(d, x0, xTe, t);
    const float est_p = sched_x_t(d, x0, xTe, t - 1);
    prior = __fsub_rn(xt, __fsub_rn(est_t, est_p));
  } else {
    const float mean = __fadd_rn(__fmul_rn(d.tab.coef1[t], x0), __fmul_rn(d.tab.coef2[t], xt));
    float stdv = 0.f;
    if (t != 0) {
      const float lo = logf(fmaxf(d.tab.post_var[t], 1e-20f));
      float logvar = lo;
      if (d.pred_var != nullptr) {
        // learned variance (diffusion_pipeline.py:246-256, gaussian_scheduler.py:110-116): v in [-1,1] -> scale in [0,1]
        float pv = d.pred_var[ip];
        if (d.pred_var_uncond != nullptr) {
          const float pvu = d.pred_var_uncond[ip];
          pv = pvu + d.guidance * (pv - pvu);
        }
        const float vs = __fadd_rn(__fmul_rn(pv, 0.5f), 0.5f);
        const float hi = logf(fmaxf(d.tab.betas[t], 1e-20f));
        logvar = __fadd_rn(__fmul_rn(vs, hi), __fmul_rn(__fsub_rn(1.f, vs), lo));
      }
      stdv = expf(0.5f * logvar);
    }
    const float nz = d.noise ? d.noise[i] : 0.f;
    prior = __fadd_rn(mean, __fmul_rn(stdv, nz));
  }
  if (d.x_prior) d.x_prior[i] = prior;
  if (d.x_0) d.x_0[i] = x0;
  if (d.x_T) d.x_T[i] = xT;
  if (d.x_next) {
    float xn = prior;
    if (d.t_next != nullptr) {
      // DDIM-form re-noise with eta == 1 (diffusion_pipeline.py:297-304)
      const float a = d.tab.alphas_cumprod[t];
      const float an = d.tab.alphas_cumprod[*d.t_next];
      const float sig2arg =
          __fdiv_rn(__fmul_rn(__fsub_rn(1.f, __fdiv_rn(a, an)), __fsub_rn(1.f, an)), __fsub_rn(1.f, a));
      const float sigma = __fsqrt_rn(sig2arg);
      const float cc = __fsqrt_rn(__fsub_rn(__fsub_rn(1.f, an), __fmul_rn(sigma, sigma)));
      const float n2 = d.noise2 ? d.noise2[i] : 0.f;
      xn = __fadd_rn(__fadd_rn(__fmul_rn(x0, __fsqrt_rn(an)), __fmul_rn(cc, xT)), __fmul_rn(sigma, n2));
    }
    d.x_next[i] = xn;
  }
}

__global__ void sched_step_kernel(const SchedStepDesc d) {
  const long long total = static_cast<long long>(d.B) * d.CHW;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
  {
    const int b = static_cast<int>(i / d.CHW);
    const long long ip = d.pred_bstride ? static_cast<long long>(b) * d.pred_bstride + (i - static_cast<long long>(b) * d.CHW) : i;
    sched_element(d, i, b, d.pred[ip]);
  }
}

// =================================================================================================
// Narrow 1x1 head (Cout <= 8) on a split NHWC tensor -> NCHW fp32, optionally fused with the scheduler update:
//   reference: unet2.py:213,267 (UnetOutBlock 256 -> 8) + gaussian_scheduler.py:80-124;  latent_embedders.py:743 (64 -> 3)
// One warp per pixel: lanes stride the channels, 8 butterfly reductions, lane o owns output channel o.
// =================================================================================================
__global__ void __launch_bounds__(256) head1x1_kernel(const HeadDesc h, const SchedStepDesc sd) {
  extern __shared__ float wsm[];  // [Cout][C] weights + [Cout] bias
  for (int e = threadIdx.x; e < h.Cout * h.C; e += blockDim.x) wsm[e] = h.w[e];
  for (int e = threadIdx.x; e < h.Cout; e += blockDim.x) wsm[h.Cout * h.C + e] = h.bias ? h.bias[e] : 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const long long npix = static_cast<long long>(h.N) * h.HW;
  const int passes = h.cfg_pair ? 2 : 1;
  for (long long pix = warp0; pix < npix; pix += nwarps) {
    float ysel[2] = {0.f, 0.f};
    for (int ps = 0; ps < passes; ++ps) {
      // cfg_pair: the estimator ran on a 2N batch — samples [0, N) without the label, [N, 2N) with it
      const __half* xh = h.in + (pix + static_cast<long long>(ps) * npix) * h.C;
      const __half* xl = xh + h.in_plane;
      float acc[8];
#pragma unroll
      for (int o = 0; o < 8; ++o) acc[o] = 0.f;
      for (int c = lane * 2; c < h.C; c += 64) {  // two channels per lane per pass (4-byte loads on both planes)
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(xh + c));
        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(xl + c));
        const float x0 = a.x + b.x, x1 = a.y + b.y;
#pragma unroll
        for (int o = 0; o < 8; ++o)
          if (o < h.Cout) acc[o] = fmaf(x1, wsm[o * h.C + c + 1], fmaf(x0, wsm[o * h.C + c], acc[o]));
      }
#pragma unroll
      for (int o = 0; o < 8; ++o)
        if (o < h.Cout)   // warp-uniform
          for (int off = 16; off > 0; off >>= 1) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], off);
      float yy = 0.f;
#pragma unroll
      for (int o = 0; o < 8; ++o)
        if (o == lane) yy = acc[o];
      ysel[ps] = yy + wsm[h.Cout * h.C + (lane < h.Cout ? lane : 0)];
    }
    if (lane < h.Cout) {
      float y = ysel[0];
      // classifier-free guidance combine (diffusion_pipeline.py:244): pred_uncond + g * (pred_cond - pred_uncond)
      if (h.cfg_pair) y = ysel[0] + h.cfg_guidance * (ysel[1] - ysel[0]);
      const int n = static_cast<int>(pix / h.HW);
      const long long i = (static_cast<long long>(n) * h.Cout + lane) * h.HW + (pix - static_cast<long long>(n) * h.HW);
      if (h.out != nullptr) h.out[i] = y;
      if (h.out_u8 != nullptr) {
        // scripts/helpers/sample_dataset.py:47-50: clip(-1,1) -> (x+1)/2*255 -> HWC -> astype(uint8) (truncation)
        const float c01 = fminf(fmaxf(y, -1.f), 1.f);
        const float v255 = __fmul_rn(__fmul_rn(__fadd_rn(c01, 1.f), 0.5f), 255.f);
        h.out_u8[pix * h.Cout + lane] = static_cast<unsigned char>(v255);
      }
      if (h.fuse_step) sched_element(sd, i, n, y);
    }
  }
}

// C == 256 (the canonical UNet head, unet2.py:213): a lane holds 8 consecutive channels of the pixel (one 16-byte load per
// plane) and its 8 x 8 weights in registers; the 8 dot products are reduced over the warp by recursive halving (9 shuffles
// instead of 40), after which lane 4*o (o = bit-reversed lane bits 4..2) holds output o.  The 4-byte-load kernel above
// needed 81 us for the 67 MB activation of B = 64 (0.84 TB/s, profiles/r02_ncu_unet_step.md).
__global__ void __launch_bounds__(256) head1x1_c256_kernel(const HeadDesc h, const SchedStepDesc sd) {
  const int lane = threadIdx.x & 31;
  float wr[8][8];
#pragma unroll
  for (int o = 0; o < 8; ++o) {
#pragma unroll
    for (int j = 0; j < 8; ++j) wr[o][j] = (o < h.Cout) ? __ldg(h.w + o * 256 + lane * 8 + j) : 0.f;
  }
  const int my_o = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
  const bool writer = (lane & 3) == 0 && my_o < h.Cout;
  const float my_bias = (writer && h.bias != nullptr) ? __ldg(h.bias + my_o) : 0.f;
  const long long warp0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const long long npix = static_cast<long long>(h.N) * h.HW;
  // two slots per iteration, both requested before the first use: the two halves of a CFG pair (samples n and n + N), or
  // two different pixels
  const long long pstep = h.cfg_pair ? nwarps : 2 * nwarps;
  for (long long pix = warp0; pix < npix; pix += pstep) {
    float ysel[2] = {0.f, 0.f};
    uint4 uh[2], ul[2];
    long long spix[2];
    bool live[2];
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
      spix[sl] = h.cfg_pair ? pix : pix + sl * nwarps;
      live[sl] = spix[sl] < npix;                  // warp-uniform
      const long long src = h.cfg_pair ? pix + static_cast<long long>(sl) * npix : (live[sl] ? spix[sl] : pix);
      const __half* xh = h.in + src * 256 + lane * 8;
      uh[sl] = *reinterpret_cast<const uint4*>(xh);
      ul[sl] = *reinterpret_cast<const uint4*>(xh + h.in_plane);
    }
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
      const __half2* hh = reinterpret_cast<const __half2*>(&uh[sl]);
      const __half2* ll = reinterpret_cast<const __half2*>(&ul[sl]);
      float x[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 a = __half22float2(hh[j]), b = __half22float2(ll[j]);
        x[2 * j] = a.x + b.x; x[2 * j + 1] = a.y + b.y;
      }
      float acc[8];
#pragma unroll
      for (int o = 0; o < 8; ++o) {
        acc[o] = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[o] = fmaf(x[j], wr[o][j], acc[o]);
      }
      float a4[4], a2[2];
      const bool b16 = lane & 16, b8 = lane & 8, b4 = lane & 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float keep = b16 ? acc[4 + j] : acc[j], send = b16 ? acc[j] : acc[4 + j];
        a4[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float keep = b8 ? a4[2 + j] : a4[j], send = b8 ? a4[j] : a4[2 + j];
        a2[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
      float a1 = (b4 ? a2[1] : a2[0]) + __shfl_xor_sync(0xffffffffu, b4 ? a2[0] : a2[1], 4);
      a1 += __shfl_xor_sync(0xffffffffu, a1, 2);
      a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
      ysel[sl] = a1 + my_bias;
    }
    if (writer) {
      // classifier-free guidance combine (diffusion_pipeline.py:244): pred_uncond + g * (pred_cond - pred_uncond)
      if (h.cfg_pair) { ysel[0] = ysel[0] + h.cfg_guidance * (ysel[1] - ysel[0]); live[1] = false; }
#pragma unroll
      for (int sl = 0; sl < 2; ++sl) {
        if (!live[sl]) continue;
        const float y = ysel[sl];
        const long long px = spix[sl];
        const int n = static_cast<int>(px / h.HW);
        const long long i = (static_cast<long long>(n) * h.Cout + my_o) * h.HW + (px - static_cast<long long>(n) * h.HW);
        if (h.out != nullptr) h.out[i] = y;
        if (h.out_u8 != nullptr) {
          const float c01 = fminf(fmaxf(y, -1.f), 1.f);
          const float v255 = __fmul_rn(__fmul_rn(__fadd_rn(c01, 1.f), 0.5f), 255.f);
          h.out_u8[px * h.Cout + my_o] = static_cast<unsigned char>(v255);
        }
        if (h.fuse_step) sched_element(sd, i, n, y);
      }
    }
  }
}

int head1x1(const HeadDesc& h, const SchedStepDesc* step, cudaStream_t s) {
  MF_REQUIRE(h.Cout >= 1 && h.Cout <= 8 && h.C % 64 == 0, "head1x1: Cout <= 8 and C % 64 == 0");
  MF_REQUIRE(!h.cfg_pair || (step != nullptr && step->pred_uncond == nullptr && h.out_u8 == nullptr),
             "head1x1 pair mode is the fused CFG step (no separate uncond prediction, no uint8 output)");
  const long long npix = static_cast<long long>(h.N) * h.HW;
  if (npix == 0) return 0;
  HeadDesc hh = h;
  SchedStepDesc sd{};
  hh.fuse_step = 0;
  if (step != nullptr) {
    sd = *step;
    hh.fuse_step = 1;
    MF_REQUIRE(sd.B == h.N && sd.CHW == h.Cout * h.HW, "fused scheduler step geometry must match the head output");
  }
  if (h.C == 256 && (reinterpret_cast<uintptr_t>(h.in) & 15) == 0 && h.in_plane % 8 == 0) {
    const int blocks = static_cast<int>(std::min<long long>((npix + 7) / 8, 148 * 2));   // 2 x 8 warps per SM, persistent
    head1x1_c256_kernel<<<blocks, 256, 0, s>>>(hh, sd);
    MF_CUDA_OK(cudaGetLastError());
    return 0;
  }
  const size_t smem = (static_cast<size_t>(h.Cout) * h.C + h.Cout) * sizeof(float);
  const int blocks = static_cast<int>(std::min<long long>((npix + 7) / 8, 148 * 8));
  head1x1_kernel<<<blocks, 256, smem, s>>>(hh, sd);
  MF_CUDA_OK(cudaGetLastError());
  return 0;
}

// =================================================================================================
// VAE reparameterisation (latent_embedders.py:20-33 DiagonalGaussianDistribution):
//   moments [B, 2E, HW] = (mean | logvar);  z = mean + exp(0.5 * clamp(logvar, -30, 20)) * noise
// =================================================================================================
__global__ void vae_reparam_kernel(const float* __restrict__ moments, const float* __restrict__ noise,
                                   float* __restrict__ z, float* __restrict__ moments_out, int B, int EHW) {
  const long long total = static_cast<long long>(B) * EHW;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long b = i / EHW, r = i - b * EHW;
    const float mean = moments[b * 2 * EHW + r];
    const float lv = moments[b * 2 * EHW + EHW + r];
    if (moments_out != nullptr) {
      moments_out[b * 2 * EHW + r] = mean;
      moments_out[b * 2 * EHW + EHW + r] = lv;
    }
    const float stdv = expf(__fmul_rn(0.5f, fminf(fmaxf(lv, -30.f), 20.f)));
    z[i] = noise != nullptr ? __fadd_rn(mean, __fmul_rn(stdv, noise[i])) : mean;
  }
}

int vae_reparam(const float* moments, const float* noise, float* z, float* moments_out, int B, int EHW, cudaStream_t s) {
  const long long total = static_cast<long long>(B) * EHW;
  if (total == 0) return 0;
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 8));
  vae_reparam_kernel<<<blocks, 256, 0, s>>>(moments, noise, z, moments_out, B, EHW);
  MF_CUDA_OK(cudaGetLastError());
  return 0;
}

int sched_step(const SchedStepDesc& d, cudaStream_t s) {
  const long long total = static_cast<long long>(d.B) * d.CHW;
  if (total == 0) return 0;
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 8));
  sched_step_kernel<<<blocks, 256, 0, s>>>(d);
  MF_CUDA_OK(cudaGetLastError());
  return 0;
}

// =================================================================================================
// GroupNorm for any channels-per-group (reference: nn.GroupNorm(eps=1e-5), biased variance)
// =================================================================================================
__global__ void __launch_bounds__(256) gn_stats_generic_kernel(const float* __restrict__ raw, float* __restrict__ mean_rstd,
                                                               int HW, int C, int G, float eps) {
  const int g = blockIdx.x, n = blockIdx.y;
  const int cpg = C / G;
  const float* base = raw + static_cast<long long>(n) * HW * C + g * cpg;
  double s = 0.0, ss = 0.0;
  const long long cnt = static_cast<long long>(HW) * cpg;
  for (long long i = threadIdx.x; i < cnt; i += blockDim.x) {
    const long long p = i / cpg;
    const int c = static_cast<int>(i - p * cpg);
    const double x = static_cast<double>(base[p * C + c]);
    s += x;
    ss += x * x;
  }
  __shared__ double sh[2][8];
  for (int off = 16; off > 0; off >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, off);
    ss += __shfl_xor_sync(0xffffffffu, ss, off);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sh[0][warp] = s; sh[1][warp] = ss; }
  __syncthreads();
  if (threadIdx.x == 0) {
    s = 0.0; ss = 0.0;
    for (int w = 0; w < 8; ++w) { s += sh[0][w]; ss += sh[1][w]; }
    const double mean = s / static_cast<double>(cnt);
    double var = ss / static_cast<double>(cnt) - mean * mean;
    if (var < 0.0) var = 0.0;
    float* dst = mean_rstd + (static_cast<long long>(n) * G + g) * 2;
    dst[0] = static_cast<float>(mean);
    dst[1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  }
}

int gn_stats_generic(const float* raw, float* mean_rstd, int N, int HW, int C, int G, float eps, cudaStream_t s) {
  MF_REQUIRE(G > 0 && C % G == 0 && N <= 65535, "gn_stats_generic: C must be divisible by the group count");
  if (N == 0) return 0;
  gn_stats_generic_kernel<<<dim3(G, N), 256, 0, s>>>(raw, mean_rstd, HW, C, G, eps);
  MF_CUDA_OK(cudaGetLastError());
  return 0;
}

__global__ void __launch_bounds__(256) gn_apply_generic_kernel(const GnApplyDesc d) {
  const long long total = static_cast<long long>(d.N) * d.HW * d.C;
  const int cpg = d.C / d.G;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % d.C);
    const int n = static_cast<int>(i / (static_cast<long long>(d.HW) * d.C));
    const float x = reinterpret_cast<const float*>(d.raw)[i];
    const float2 mr = *reinterpret_cast<const float2*>(d.mean_rstd + (static_cast<long long>(n) * d.G + c / cpg) * 2);
    float y = (x - mr.x) * mr.y * d.gamma[c] + d.beta[c];
    if (d.act != 0) y = swish(y);
    if (d.res_kind == kResSplit) {
      const __half* rh = reinterpret_cast<const __half*>(d.res);
      y += join16(rh[i], rh[d.res_plane + i]);
    } else if (d.res_kind == kResRaw) {
      y += reinterpret_cast<const float*>(d.res)[i];
    }
    if (d.emb != nullptr) {
      const long long er = d.emb_index ? d.emb_index[n] : static_cast<long long>(n);
      y += d.emb[er * d.emb_stride + c];
    }
    split16(y, d.out[i], d.out[d.out_plane + i]);
  }
}

int gn_apply_generic(const GnApplyDesc& d, cudaStream_t s) {
  MF_REQUIRE(d.raw_plane == 0 && d.mean_rstd != nullptr && d.C % d.G == 0, "gn_apply_generic: raw fp32 input + mean/rstd");
  const long long total = static_cast<long long>(d.N) * d.HW * d.C;
  if (total == 0) return 0;
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 32));
  gn_apply_generic_kernel<<<blocks, 256, 0, s>>>(d);
  MF_CUDA_OK(cudaGetLastError());
  return 0;
}

// =================================================================================================
// Vector quantiser lookup (reference: latent_embedders.py:40-72 VectorQuantizer.forward; used by VQVAE.decode :314-316)
//   dist[b, k] = sum_c z^2 + sum_c e_k^2 - 2 * sum_c z_c e_kc   (the reference's expansion, not ||z - e||^2, so near-ties
//   resolve the same way); argmin keeps the FIRST minimum; z_q = z + (e - z) in fp32 (the straight-through form :69).
// One thread per latent vector; the codebook streams through shared memory in tiles of 1024 rows.
// =================================================================================================
constexpr int kVqMaxC = 16;
constexpr int kVqTile = 1024;
__global__ void __launch_bounds__(256) vq_quantize_kernel(const float* __restrict__ z, const float* __restrict__ cb,
                                                          float* __restrict__ zq, int* __restrict__ idx_out, int B, int C,
                                                          int HW, int K) {
  extern __shared__ float tile[];          // [kVqTile][C] codebook rows + [kVqTile] squared norms
  float* enorm = tile + kVqTile * C;
  const long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long nvec = static_cast<long long>(B) * HW;
  const bool live = v < nvec;
  const int b = live ? static_cast<int>(v / HW) : 0;
  const int p = live ? static_cast<int>(v - static_cast<long long>(b) * HW) : 0;
  float zc[kVqMaxC];
  float zz = 0.f;
#pragma unroll
  for (int c = 0; c < kVqMaxC; ++c) {
    zc[c] = (live && c < C) ? z[(static_cast<long long>(b) * C + c) * HW + p] : 0.f;
    if (c < C) zz = __fadd_rn(zz, __fmul_rn(zc[c], zc[c]));
  }
  float best = INFINITY;
  int best_k = 0;
  for (int k0 = 0; k0 < K; k0 += kVqTile) {
    const int kn = min(kVqTile, K - k0);
    __syncthreads();
    for (int e = threadIdx.x; e < kn * C; e += blockDim.x) tile[e] = cb[static_cast<long long>(k0) * C + e];
    __syncthreads();
    for (int r = threadIdx.x; r < kn; r += blockDim.x) {
      float ee = 0.f;
      for (int c = 0; c < C; ++c) ee = __fadd_rn(ee, __fmul_rn(tile[r * C + c], tile[r * C + c]));
      enorm[r] = ee;
    }
    __syncthreads();
    if (live) {
      for (int r = 0; r < kn; ++r) {
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < kVqMaxC; ++c)
          if (c < C) dot = fmaf(zc[c], tile[r * C + c], dot);
        const float dist = __fsub_rn(__fadd_rn(zz, enorm[r]), __fmul_rn(2.f, dot));
        if (dist < best) { best = dist; best_k = k0 + r; }
      }
    }
  }
  if (live) {
    if (idx_out != nullptr) idx_out[v] = best_k;
    for (int c = 0; c < C; ++c) {
      const float e = cb[static_cast<long long>(best_k) * C + c];
      zq[(static_cast<long long>(b) * C + c) * HW + p] = __fadd_rn(zc[c], __fsub_rn(e, zc[c]));
    }
  }
}

int vq_quantize(const float* z, const float* codebook, float* z_q, int* idx_out, int B, int C, int HW, int K,
                cudaStream_t s) {
  MF_REQUIRE(C >= 1 && C <= kVqMaxC && K >= 1, "vq_quantize: emb_channels must be in [1, 16]");
  const long long nvec = static_cast<long long>(B) * HW;
  if (nvec == 0) return 0;
  const size_t smem = static_cast<size_t>(kVqTile) * (C + 1) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    MF_CUDA_OK(cudaFuncSetAttribute(vq_quantize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    kVqTile * (kVqMaxC + 1) * static_cast<int>(sizeof(float))));
    attr_set = true;
  }
  vq_quantize_kernel<<<static_cast<unsigned>((nvec + 255) / 256), 256, smem, s>>>(z, codebook, z_q, idx_out, B, C, HW, K);
  MF_CUDA_OK(cudaGetLastError());
  return 0;
}

MF_DEFINE_SATURATION_READER(sat_read_kernels)

}  // namespace mf
