// medfusion_b200 — shared device/host helpers for the sm_100a kernels.
//
// Everything here is hand-written for Blackwell (sm_100a): mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (alloc / mma / commit / ld) inline-PTX wrappers, plus the TF32 hi/lo split used by the
// error-compensated 3xTF32 convolution path (SURVEY.md finding 5: fp32 parity needs 3xTF32).
#pragma once

#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

namespace mf {

// ------------------------------------------------------------------------------------------------
// Host-side error plumbing (thread-local last error string; C-ABI returns int status)
// ------------------------------------------------------------------------------------------------
void set_error(const std::string& msg);
const char* get_error();

#define MF_CUDA_OK(expr)                                                                       \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      ::mf::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " @" +       \
                      __FILE__ + ":" + std::to_string(__LINE__));                              \
      return 1;                                                                                \
    }                                                                                          \
  } while (0)

#define MF_REQUIRE(cond, msg)                                                                  \
  do {                                                                                         \
    if (!(cond)) {                                                                             \
      ::mf::set_error(std::string("requirement failed: ") + #cond + " — " + (msg) + " @" +     \
                      __FILE__ + ":" + std::to_string(__LINE__));                              \
      return 2;                                                                                \
    }                                                                                          \
  } while (0)

// ------------------------------------------------------------------------------------------------
// fp16 hi/lo split ("fp16x3").  hi = fp16(x) (11 significant bits, the same as TF32), lo = fp16(x - hi).
// x*y is recovered as lo*hi' + hi*lo' + hi*hi' with fp32 accumulation: |x - hi - lo| <= 2^-22 |x| (or 3e-8 absolute
// once lo is subnormal), the fp16 x fp16 products are exact in fp32.  Compared with the TF32 split the operands are
// half as wide and kind::f16 MMAs cover K = 16 per instruction instead of 8: twice the math per tensor-core cycle and
// per shared-memory byte, at identical accuracy.  Range: |x| <= 65504 (larger values saturate); weights are
// pre-scaled by a power of two so that their lo parts stay in fp16's normal range.
// ------------------------------------------------------------------------------------------------
typedef __half sp_t;  // element type of the split (hi / lo) planes
// Sticky saturation counter.  The reference computes in fp32 range; the split planes are fp16, so a value beyond
// +-65504 (or a non-finite one) is clamped and parity is lost from there on.  Every clamp bumps this counter (one
// instance per translation unit: the library is built without relocatable device code; mf_saturation_count sums
// them), so the caller can tell that a result left the representable range instead of silently getting a wrong one.
static __device__ unsigned int g_mf_saturated = 0;
__device__ __forceinline__ float sat16(float x) {
  const float c = fminf(fmaxf(x, -65504.f), 65504.f);
  if (c != x) atomicAdd(&g_mf_saturated, 1u);   // also true for NaN (fmaxf(NaN, a) = a)
  return c;
}
__device__ __forceinline__ void split16(float x, __half& hi, __half& lo) {
  const float xc = sat16(x);
  hi = __float2half_rn(xc);
  lo = __float2half_rn(xc - __half2float(hi));   // |lo| <= 2^-11 |hi|: always in range
}
// two values at once: packed conversions (cvt.rn.f16x2.f32), same results as split16 on each
__device__ __forceinline__ void split16x2(float a, float b, uint32_t& hi2, uint32_t& lo2) {
  const float ac = sat16(a), bc = sat16(b);
  const __half2 h = __floats2half2_rn(ac, bc);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(ac - hf.x, bc - hf.y);
  hi2 = *reinterpret_cast<const uint32_t*>(&h);
  lo2 = *reinterpret_cast<const uint32_t*>(&l);
}
// Streaming variant for instruction-bound loops (GroupNorm-apply: 36 thread instructions per element, the per-value
// compare + branch + warp-aggregated atomic of sat16 was a fifth of them): same clamp, the event is OR-ed into a
// per-thread flag that the caller reports ONCE with sat16_report (the counter then counts threads that clamped, which is
// all its users need: zero / non-zero, and a lower bound on the number of values).
__device__ __forceinline__ void split16x2_flag(float a, float b, uint32_t& hi2, uint32_t& lo2, bool& clamped) {
  const float ac = fminf(fmaxf(a, -65504.f), 65504.f), bc = fminf(fmaxf(b, -65504.f), 65504.f);
  clamped = clamped || (ac != a) || (bc != b);      // true for NaN as well
  const __half2 h = __floats2half2_rn(ac, bc);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(ac - hf.x, bc - hf.y);
  hi2 = *reinterpret_cast<const uint32_t*>(&h);
  lo2 = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void sat16_report(bool clamped) {
  if (clamped) atomicAdd(&g_mf_saturated, 1u);
}
// host side of the counter: defined once per translation unit that launches kernels using split16
#define MF_DEFINE_SATURATION_READER(fn)                                                                     \
  int fn(unsigned long long* total, int reset, cudaStream_t s) {                                            \
    unsigned int v = 0;                                                                                     \
    MF_CUDA_OK(cudaMemcpyFromSymbolAsync(&v, g_mf_saturated, sizeof(v), 0, cudaMemcpyDeviceToHost, s));     \
    MF_CUDA_OK(cudaStreamSynchronize(s));                                                                   \
    if (reset && v != 0) {                                                                                  \
      const unsigned int z = 0;                                                                             \
      MF_CUDA_OK(cudaMemcpyToSymbolAsync(g_mf_saturated, &z, sizeof(z), 0, cudaMemcpyHostToDevice, s));     \
      MF_CUDA_OK(cudaStreamSynchronize(s));                                                                 \
    }                                                                                                       \
    *total += v;                                                                                            \
    return 0;                                                                                               \
  }
__device__ __forceinline__ float join16(__half hi, __half lo) { return __half2float(hi) + __half2float(lo); }
// four consecutive channels: 8-byte vector accesses on both planes
__device__ __forceinline__ float4 ld_join4(const __half* hi, const __half* lo) {
  const uint2 uh = *reinterpret_cast<const uint2*>(hi);
  const uint2 ul = *reinterpret_cast<const uint2*>(lo);
  const __half2 h0 = *reinterpret_cast<const __half2*>(&uh.x), h1 = *reinterpret_cast<const __half2*>(&uh.y);
  const __half2 l0 = *reinterpret_cast<const __half2*>(&ul.x), l1 = *reinterpret_cast<const __half2*>(&ul.y);
  const float2 a = __half22float2(h0), b = __half22float2(h1), c = __half22float2(l0), d = __half22float2(l1);
  return make_float4(a.x + c.x, a.y + c.y, b.x + d.x, b.y + d.y);
}
__device__ __forceinline__ void st_split4(__half* hi, __half* lo, float4 v) {
  __half h[4], l[4];
  split16(v.x, h[0], l[0]);
  split16(v.y, h[1], l[1]);
  split16(v.z, h[2], l[2]);
  split16(v.w, h[3], l[3]);
  uint2 uh, ul;
  uh.x = *reinterpret_cast<uint32_t*>(&h[0]); uh.y = *reinterpret_cast<uint32_t*>(&h[2]);
  ul.x = *reinterpret_cast<uint32_t*>(&l[0]); ul.y = *reinterpret_cast<uint32_t*>(&l[2]);
  *reinterpret_cast<uint2*>(hi) = uh;
  *reinterpret_cast<uint2*>(lo) = ul;
}

// programmatic dependent launch (no-ops when the kernel was launched without the attribute)
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ------------------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a lost arrival traps (kernel aborts with an error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) {
      printf("mf: mbarrier timeout block=(%d,%d) thread=%d\n", blockIdx.x, blockIdx.y, threadIdx.x);
      __trap();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor), tile mode, completion on an mbarrier
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3), "r"(c4)
      : "memory");
}

// TMA prefetch of a 3-D box into L2 (no shared-memory destination, no completion to wait for)
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* m, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
               :
               : "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

// TMA store (shared::cta -> global, tile mode), bulk-group completion.  The generic-proxy writes that filled the tile
// must be made visible to the async proxy (fence_proxy_async by every writer, then a barrier) before the issue.
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3,
                                             int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
      :
      : "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all committed groups of this thread have finished READING shared memory (the tile may be overwritten)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// ------------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, MMA, commit, load
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]^T, fp16 inputs (K = 16 per instruction), FP32 accumulate, single-CTA group.
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// All previously issued MMAs of this thread arrive on `bar` when they complete.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives row (lane base + i).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ------------------------------------------------------------------------------------------------
// Thread-block clusters / CTA pairs (cta_group::2): two CTAs on one TPC share the B operand of a 256-row MMA
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address of `p` in THIS CTA -> shared::cluster address of the same offset in CTA `rank`
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
// Remote arrive on a barrier of the pair's leader CTA.  RELAXED: what the arrival publishes is "this warp's tcgen05.ld of
// the TMEM buffer has completed", which tcgen05.wait::ld + tcgen05.fence::before_thread_sync order — no generic-proxy data
// is handed over, so the release.cluster form (a MEMBAR.ALL.GPU + ERRBAR per drained chunk: 17 % of all stall samples in
// profiles/r02_conv_tc_ncu_32x32_before.md) is not needed.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads of a CTA pair: data lands in the issuing CTA's smem, completion bytes are credited to the barrier at
// `bar_cluster_addr` (the leader CTA's full barrier).
__device__ __forceinline__ void tma_load_3d_2sm(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d_2sm(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// 256-row MMA across the CTA pair (issued by the leader CTA only)
__device__ __forceinline__ void umma_f16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// completion of all prior MMAs arrives on the barrier at this smem offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}

// ------------------------------------------------------------------------------------------------
// UMMA descriptors (bit layouts: cute/arch/mma_sm100_desc.hpp in the CUTLASS tree; PTX ISA
// "tcgen05 shared memory descriptor" / "instruction descriptor").
// ------------------------------------------------------------------------------------------------
// K-major operand tile, 128-byte swizzle: rows are 128 B apart, 8-row groups are 1024 B apart.
// base_offset: descriptor bits [49,52).  Left 0 everywhere: measured on B200 (tests/test_gpu_kernels.py row-patch shapes), a tile
// addressed `r` rows into a TMA-written 128B-swizzled patch (start + r * 128 B) is read correctly with base offset 0 — the
// swizzle XOR is taken from the absolute shared-memory address bits [7,10) — and wrongly with base offset r.
__device__ __forceinline__ uint64_t umma_smem_desc_sw128(uint32_t smem_addr, uint32_t base_offset = 0) {
  uint64_t d = static_cast<uint64_t>(base_offset & 7u) << 49;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);  // start address  [0,14)
  d |= static_cast<uint64_t>(1) << 16;                      // leading byte offset (unused for SW128 K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;              // stride byte offset: 8 rows * 128 B
  d |= static_cast<uint64_t>(1) << 46;                      // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;                      // layout type: SWIZZLE_128B
  return d;
}
// kind::f16 with fp16 operands, A/B K-major, FP32 accumulator, M x N tile.
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
  return (1u << 4)                              // c_format = F32
         | (0u << 7)                            // a_format = F16
         | (0u << 10)                           // b_format = F16
         | (static_cast<uint32_t>(N >> 3) << 17)  // n_dim
         | (static_cast<uint32_t>(M >> 4) << 24); // m_dim
}

}  // namespace mf
