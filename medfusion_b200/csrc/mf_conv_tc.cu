// See mf_conv_tc.cuh for the design.  sm_100a only: TMA + mbarrier pipeline + tcgen05.mma (TMEM accumulators).
#include "mf_conv_tc.cuh"

#include <algorithm>
#include <cmath>
#include <mutex>

namespace mf {

int g_default_drain_interval = 3;
int g_default_cta_group = 0;
int g_default_block_n = 0;    // 0 = auto
// Fused GroupNorm epilogue: correct and tested, but OFF by default — at the levels where it applies a CTA pair owns about
// one tile, so the normalisation work runs exposed at the kernel's tail instead of in a bandwidth-bound kernel of its own:
// canonical step 8.85 ms fused vs 8.72 ms separate, 66.7 vs 67.9 images/s (same box, profiles/r02_gn_fusion.md).
int g_fuse_gn = 0;
int g_stream_k = 1;           // persistent stream-K schedule (0: one tile per CTA group)
int g_row_patch = 1;          // 1: row-patch mode (TcCfg ROW3) for the eligible 3x3 layers (tiles of one image row)
int g_split_fill = 8;         // small-batch fill: layers with fewer tiles than half the SM pairs split K (>= this many K blocks per group); 0 = off
float g_debias_eps_per_kblock = -1.0f;  // < 0: calibrated table (default); 0: off; > 0: explicit relative correction per K block  // 0 = auto (pairs whenever a conv has >= 2 M tiles)

// =================================================================================================
// Device side
// =================================================================================================
// ROW3 ("row patch" mode, 3x3 stride-1 convolutions whose tile is 128 pixels of ONE image row): a pipeline stage holds the
// (128 + 2)-pixel patch of one input row and one 64-channel slab (hi and lo planes) plus the weight tiles of the THREE taps
// that read this row; the MMAs of tap dx address the same patch shifted by dx rows of 128 bytes (descriptor start +dx*128 B;
// the swizzle follows the absolute address, no base offset).  The activation slab is fetched 3 times per tile instead of 9: profiles/r02_row_patch.md (the 128x128 /
// 256x256 levels of the VAE were bound by the L2 -> SM fabric, ~8-9 TB/s, not by shared memory or the tensor pipe).
constexpr int kTcPatchRows = kTcBlockM + 2;
template <int BLOCK_N, int CG, bool ROW3 = false>
struct TcCfg {
  static constexpr int kABytes = kTcBlockM * 128;        // one 128-row x 128-byte plane tile
  static constexpr int kAPatchBytes = ((kTcPatchRows * 128 + 1023) / 1024) * 1024;   // 130 rows, padded to the swizzle atom
  static constexpr int kBRows = BLOCK_N / CG;            // weight rows this CTA stages (a CTA pair splits B)
  static constexpr int kBBytes = kBRows * 128;
  static constexpr int kStageBytes = ROW3 ? (2 * kAPatchBytes + 6 * kBBytes) : (2 * kABytes + 2 * kBBytes);
  static constexpr int kStageTxBytes = ROW3 ? (2 * kTcPatchRows * 128 + 6 * kBBytes) : kStageBytes;   // bytes TMA delivers per stage
  // epilogue staging: per column half one [128 rows][128 B] tile (32 fp32 channels, or 32 fp16 channels hi | lo) that a
  // single thread hands to the TMA store engine — the global write is one bulk tensor store per 16 KB, not 32
  // row-strided STG.128 per thread (profiles/r02_conv_tc_ncu_32x32_before.md: the old epilogue cost 24 k cycles per tile)
  static constexpr int kStagingHalf = kTcBlockM * 128;
  static constexpr int kStagingBytes = 2 * kStagingHalf;
  static constexpr int kRedBytes = 4 * (BLOCK_N / 8) * 2 * 4;   // GroupNorm partial sums [4 quarters][BLOCK_N/8][2]
  static constexpr int kXchBytes = 2 * (BLOCK_N / 8) * 8;           // fused GroupNorm: the pair's other CTA writes its per-slab
                                                                    // (sum, sumsq) here (double-buffered by tile parity)
  static constexpr int kAuxBytes = 512 + kRedBytes + BLOCK_N * 4 + kXchBytes;   // barriers + TMEM slot | red | bias | xch
  static constexpr int kSmemMax = 227 * 1024;
  static constexpr int kStagesFit = (kSmemMax - kAuxBytes - kStagingBytes) / kStageBytes;
  static constexpr int kStages = kStagesFit > 6 ? 6 : kStagesFit;
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + kAuxBytes;
  static constexpr int kAccBufs = (512 / BLOCK_N) > 4 ? 4 : (512 / BLOCK_N);  // ring of partial-sum accumulators
  static constexpr int kTmemCols = kAccBufs * BLOCK_N;   // 512 columns for BLOCK_N >= 128
  static constexpr int kColsPerWarp = BLOCK_N / 2;       // 8 drain warps: 4 lane quarters x 2 column halves
  static_assert(kStages >= 2, "pipeline needs at least two stages");
  static_assert(kSmemBytes <= kSmemMax, "shared memory budget");
};

// Why the accumulation leaves the tensor core: tcgen05.mma adds into its fp32 TMEM accumulator with
// truncation (round-toward-zero), so a long chain of MMAs drifts by ~0.5 ulp(|acc|) per instruction
// (measured: 2e-4 abs at K=2304, i.e. far outside the 1e-5 parity budget).  We therefore let the
// tensor core produce SHORT partial sums (`drain_interval` 32-channel K blocks: the tiny cross-term MMAs first,
// then the hi*hi MMAs) into a ring of TMEM buffers, and the drain warps add each finished partial (de-biased)
// into round-to-nearest fp32 running sums held in registers while the next K blocks are being multiplied.
//
// CG == 2: two CTAs of a cluster (one TPC) form a pair.  Each stages its own 128 pixel rows of A and HALF of the
// weight tile; the leader CTA issues 256-row tcgen05.mma.cta_group::2 instructions that read both CTAs' shared
// memory and write both CTAs' TMEM.  Per CTA a stage shrinks from 96 KB to 64 KB (3 stages in flight instead of 2)
// and the weight traffic per SM halves.
//
// Persistent stream-K schedule: the launch has at most one CTA group per SM (pair).  All (tile, K block) units of
// the convolution are split EVENLY over the groups, so there is no partial last wave (512 tiles on 148 SMs used to
// cost 4 waves for 3.46 waves of work).  A group whose range starts inside a tile writes that tile's partial sums
// to a scratch buffer and raises a flag; the group that owns the tile's first K block waits for the flag(s), adds
// the partials and runs the epilogue.  Partial writers always process that segment FIRST, finalizers process theirs
// LAST, so nobody waits on work that has not been scheduled.
// GN: compiled with the fused GroupNorm epilogue (p.gn_mode != 0).  A separate instantiation because that epilogue is large:
// it runs as a rolled loop over 32-column chunks, while the plain epilogue stays fully unrolled (no register-window moves).
template <int BLOCK_N, int CG, bool GN, bool ROW3 = false>
__global__ void __launch_bounds__(kTcThreads, 1)
conv_tc_kernel(const __grid_constant__ TcMaps maps, const ConvTcParams p) {
  using Cfg = TcCfg<BLOCK_N, CG, ROW3>;
  constexpr int CPW = Cfg::kColsPerWarp;
  constexpr int NB = Cfg::kAccBufs;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // Programmatic dependent launch: let the next kernel of the stream be scheduled as this one's CTAs retire (it parks
  // at its own griddepcontrol.wait until this grid has completed), and run our own set-up — barrier init, TMEM
  // allocation, cluster handshake — while the previous kernel drains.  Both are no-ops without the launch attribute.
  pdl_launch_dependents();
  // swizzled TMA tiles need 1024-byte alignment; the dynamic segment starts at the CTA's shared window (no static
  // shared memory in this kernel), which is aligned — checked, not assumed
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) {
    if (threadIdx.x == 0) printf("mf: conv_tc dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* staging = smem + Cfg::kStages * Cfg::kStageBytes;   // [2 column halves][16 KB]
  uint8_t* aux = staging + Cfg::kStagingBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);  // [kStages] TMA -> MMA (leader CTA's are the live ones)
  uint64_t* empty_bar = full_bar + Cfg::kStages;           // [kStages] MMA -> TMA (every CTA)
  uint64_t* acc_full_bar = empty_bar + Cfg::kStages;       // [NB] MMA -> drain (every CTA)
  uint64_t* acc_empty_bar = acc_full_bar + NB;             // [NB] drain -> MMA (leader CTA's)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty_bar + NB);
  float* red = reinterpret_cast<float*>(aux + 512);        // [4 quarters][BLOCK_N/8][2]
  float* bias_s = reinterpret_cast<float*>(aux + 512 + Cfg::kRedBytes);   // [BLOCK_N] bias of the current tile's columns
  float2* xch = reinterpret_cast<float2*>(aux + 512 + Cfg::kRedBytes + BLOCK_N * 4);   // [2][BLOCK_N/8] written by the peer CTA
  uint64_t* xch_bar = reinterpret_cast<uint64_t*>(aux + 256);              // [2] peer -> this CTA (fused GroupNorm exchange)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;

  // ---- work decomposition ---------------------------------------------------------------------------
  const int cin = p.C0 + p.C1;
  const int cblks = cin / kTcBlockK;
  const int nkb = p.ntaps * cblks;
  const int drain = p.drain_interval < 1 ? 1 : p.drain_interval;  // K blocks per TMEM partial sum
  const int group = static_cast<int>(blockIdx.x) / CG;            // CTA group (pair) index
  const int ngroups = static_cast<int>(gridDim.x) / CG;
  const long long total_units = static_cast<long long>(p.num_tiles) * nkb;
  const long long u_begin = total_units * group / ngroups;
  const long long u_end = total_units * (group + 1) / ngroups;

  // tile id -> coordinates.  tile = ((phase * n_tiles) + nt) * m_groups + mg   (M fastest: neighbours share weights)
  struct TileCoord { int nt, phase, n0, h0, w0, th, tw; };
  auto decode_tile = [&](int tile) {
    TileCoord c;
    const int mg = tile % p.m_groups;
    const int rest = tile / p.m_groups;
    c.nt = rest % p.n_tiles;
    c.phase = rest / p.n_tiles;
    const int mt = mg * CG + static_cast<int>(cta_rank);   // may run past the real tile count: padding CTA of an odd pair
    c.tw = mt % p.tiles_w;
    c.th = (mt / p.tiles_w) % p.tiles_h;
    const int tn = mt / (p.tiles_w * p.tiles_h);
    c.n0 = tn * p.bn; c.h0 = c.th * p.bh; c.w0 = c.tw * p.bw;
    return c;
  };

  // ---- one-time setup ---------------------------------------------------------------------------
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.a[0]);
    tma_prefetch_desc(&maps.a[1]);
    tma_prefetch_desc(&maps.w);
    tma_prefetch_desc(&maps.o[0]);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < NB; ++b) {
      mbar_init(&acc_full_bar[b], 1);
      mbar_init(&acc_empty_bar[b], kTcDrainWarps * CG);
    }
    mbar_init(&xch_bar[0], BLOCK_N / 8);
    mbar_init(&xch_bar[1], BLOCK_N / 8);
    fence_mbar_init();
  }
  if (warp == 1) {
    if (CG == 2) { tmem_alloc_2sm(tmem_slot, Cfg::kTmemCols); tmem_relinquish_2sm(); }
    else { tmem_alloc(tmem_slot, Cfg::kTmemCols); tmem_relinquish(); }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Weights do not depend on the previous kernel: the weight tiles of this CTA's first pipeline stages are prefetched into L2
  // before griddepcontrol.wait, so the pipeline fill after the wait finds them there (0.8 GB of weights per step do not stay
  // in the 126 MB L2 from one timestep to the next).
  if (warp == 0 && lane == 0 && u_begin < u_end) {
    const int tile = static_cast<int>(u_begin / nkb);
    const int kb0 = static_cast<int>(u_begin - static_cast<long long>(tile) * nkb);
    const TileCoord tc = decode_tile(tile);
    const int brow = tc.phase * p.Cout + tc.nt * BLOCK_N + static_cast<int>(cta_rank) * Cfg::kBRows;
    const int taps_per_kb = ROW3 ? (p.up2 ? 2 : 3) : 1;
    const int nk = min(Cfg::kStages, nkb - kb0) * taps_per_kb;
    for (int i = 0; i < nk; ++i) {
      const int kw = (kb0 * taps_per_kb + i) * kTcBlockK;
      tma_prefetch_3d(&maps.w, kw, brow, 0);
      tma_prefetch_3d(&maps.w, kw, brow, 1);
    }
  }
  pdl_wait();   // everything above touched only shared memory / TMEM (and prefetched constants); dependent reads start below

  // Register re-partitioning (setmaxnreg): the TMA / MMA warpgroup needs a handful of registers, the two drain
  // warpgroups hold 128 fp32 running sums per thread.  384 threads x 168 = 128 x 72 + 256 x 216.
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 72;" ::: "memory");
  if (warp == 0) {
    // ===================== TMA producer (every CTA stages its own A rows and its share of B) =====================
    if (lane == 0) {
      int it = 0;  // running K-block count -> smem stage / phase
      for (long long u = u_begin; u < u_end;) {
        const int tile = static_cast<int>(u / nkb);
        const int kb0 = static_cast<int>(u - static_cast<long long>(tile) * nkb);
        const int kb1 = static_cast<int>(min(static_cast<long long>(nkb), kb0 + (u_end - u)));
        const TileCoord tc = decode_tile(tile);
        const int oa = tc.phase >> 1, ob = tc.phase & 1;
        const int brow = tc.phase * p.Cout + tc.nt * BLOCK_N + static_cast<int>(cta_rank) * Cfg::kBRows;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % Cfg::kStages;
          const uint32_t ph = (it / Cfg::kStages) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* st = smem + s * Cfg::kStageBytes;
          // K order is channel-block major, tap minor: the 9 taps of one 32-channel slab are consecutive, so the
          // shifted re-reads of the same activation lines hit in L2 (tap-major order thrashed it: 1.9 GB of DRAM
          // reads for a 134 MB input at Cin=1024, profiles/r01_conv_tc_ncu_raw.md).
          const int cb = kb / p.ntaps;
          const int tap = kb - cb * p.ntaps;
          const int c = cb * kTcBlockK;
          // up2: nearest-x2 + conv3x3 folded into four 2x2 phase convolutions on the low-resolution input:
          // output pixel (2h+oa, 2w+ob) reads input rows {h+oa-1, h+oa} and cols {w+ob-1, w+ob} (pre-summed weights)
          const int x = p.up2 ? (tc.w0 + ob - 1 + (tap & 1)) : (tc.w0 + p.dx[tap]);
          const int y = p.up2 ? (tc.h0 + oa - 1 + (tap >> 1)) : (tc.h0 + p.dy[tap]);
          const CUtensorMap* ma;
          int cc = c;
          if (p.per_tap_map) {
            ma = &maps.a[p.tap_map[tap]];
          } else if (c < p.C0) {
            ma = &maps.a[0];
          } else {
            ma = &maps.a[1];
            cc = c - p.C0;
          }
          if (ROW3) {
            // tap = input row (3x3: dy = tap - 1; folded upsample: row ty of the phase's 2x2 taps): one 130-pixel patch of that
            // row, the weight tiles of the rt taps that read it (3, or 2 for a phase of the folded upsample)
            const int rt = p.up2 ? 2 : 3;
            const int yr = p.up2 ? (tc.h0 + oa - 1 + tap) : (tc.h0 + tap - 1);
            const int xr = p.up2 ? (tc.w0 + ob - 1) : (tc.w0 - 1);
            const int kw0 = (cb * (p.up2 ? 4 : 9) + tap * rt) * kTcBlockK;
            const uint32_t tx_bytes = 2 * kTcPatchRows * 128 + 2 * rt * Cfg::kBBytes;
            uint8_t* bt = st + 2 * Cfg::kAPatchBytes;
            if (CG == 2) {
              if (leader) mbar_expect_tx(&full_bar[s], 2 * tx_bytes);
              const uint32_t fb = mapa_u32(smem_u32(&full_bar[s]), 0);
              tma_load_5d_2sm(st, ma, fb, cc, xr, yr, tc.n0, 0);
              tma_load_5d_2sm(st + Cfg::kAPatchBytes, ma, fb, cc, xr, yr, tc.n0, 1);
#pragma unroll
              for (int dxi = 0; dxi < 3; ++dxi) {
                if (dxi >= rt) break;
                tma_load_3d_2sm(bt + (2 * dxi) * Cfg::kBBytes, &maps.w, fb, kw0 + dxi * kTcBlockK, brow, 0);
                tma_load_3d_2sm(bt + (2 * dxi + 1) * Cfg::kBBytes, &maps.w, fb, kw0 + dxi * kTcBlockK, brow, 1);
              }
            } else {
              mbar_expect_tx(&full_bar[s], tx_bytes);
              tma_load_5d(st, ma, &full_bar[s], cc, xr, yr, tc.n0, 0);
              tma_load_5d(st + Cfg::kAPatchBytes, ma, &full_bar[s], cc, xr, yr, tc.n0, 1);
#pragma unroll
              for (int dxi = 0; dxi < 3; ++dxi) {
                if (dxi >= rt) break;
                tma_load_3d(bt + (2 * dxi) * Cfg::kBBytes, &maps.w, &full_bar[s], kw0 + dxi * kTcBlockK, brow, 0);
                tma_load_3d(bt + (2 * dxi + 1) * Cfg::kBBytes, &maps.w, &full_bar[s], kw0 + dxi * kTcBlockK, brow, 1);
              }
            }
          } else if (CG == 2) {
            // all bytes of the pair are credited to the LEADER's full barrier
            if (leader) mbar_expect_tx(&full_bar[s], 2 * Cfg::kStageBytes);
            const uint32_t fb = mapa_u32(smem_u32(&full_bar[s]), 0);
            tma_load_5d_2sm(st, ma, fb, cc, x, y, tc.n0, 0);
            tma_load_5d_2sm(st + Cfg::kABytes, ma, fb, cc, x, y, tc.n0, 1);
            tma_load_3d_2sm(st + 2 * Cfg::kABytes, &maps.w, fb, kb * kTcBlockK, brow, 0);
            tma_load_3d_2sm(st + 2 * Cfg::kABytes + Cfg::kBBytes, &maps.w, fb, kb * kTcBlockK, brow, 1);
          } else {
            mbar_expect_tx(&full_bar[s], Cfg::kStageBytes);
            tma_load_5d(st, ma, &full_bar[s], cc, x, y, tc.n0, 0);
            tma_load_5d(st + Cfg::kABytes, ma, &full_bar[s], cc, x, y, tc.n0, 1);
            tma_load_3d(st + 2 * Cfg::kABytes, &maps.w, &full_bar[s], kb * kTcBlockK, brow, 0);
            tma_load_3d(st + 2 * Cfg::kABytes + Cfg::kBBytes, &maps.w, &full_bar[s], kb * kTcBlockK, brow, 1);
          }
        }
        u += kb1 - kb0;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread of the leader CTA) =====================
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = umma_idesc_f16(kTcBlockM * CG, BLOCK_N);
      int it = 0;  // running K-block count (smem ring)
      int jc = 0;  // running chunk count (TMEM ring)
      for (long long u = u_begin; u < u_end;) {
        const int tile = static_cast<int>(u / nkb);
        const int kb0 = static_cast<int>(u - static_cast<long long>(tile) * nkb);
        const int kb1 = static_cast<int>(min(static_cast<long long>(nkb), kb0 + (u_end - u)));
        for (int kb = kb0; kb < kb1; ++jc) {
          const int buf = jc % NB;
          mbar_wait(&acc_empty_bar[buf], ((jc / NB) & 1) ^ 1);  // drained NB chunks ago (first use: free)
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * BLOCK_N;
          const int kb_end = min(kb1, kb + drain);
          bool first = true;
          for (; kb < kb_end; ++kb, ++it) {
            const int s = it % Cfg::kStages;
            const uint32_t ph = (it / Cfg::kStages) & 1;
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            const uint32_t st = smem_u32(smem + s * Cfg::kStageBytes);
            if (ROW3) {
#pragma unroll
              for (int dxi = 0; dxi < 3; ++dxi) {
                if (dxi >= (p.up2 ? 2 : 3)) break;
                // measured on B200: the 128-byte swizzle is a function of the ABSOLUTE shared-memory address bits, so a tile that
                // starts dx rows into the TMA-written patch needs no descriptor base offset (with base offset = dx the results are wrong)
                const uint64_t ra_hi = umma_smem_desc_sw128(st + dxi * 128);
                const uint64_t ra_lo = umma_smem_desc_sw128(st + Cfg::kAPatchBytes + dxi * 128);
                const uint64_t rb_hi = umma_smem_desc_sw128(st + 2 * Cfg::kAPatchBytes + (2 * dxi) * Cfg::kBBytes);
                const uint64_t rb_lo = umma_smem_desc_sw128(st + 2 * Cfg::kAPatchBytes + (2 * dxi + 1) * Cfg::kBBytes);
#pragma unroll
                for (int k = 0; k < kTcBlockK / 16; ++k) {
                  const uint64_t koff = static_cast<uint64_t>(k * 2);
                  if (CG == 2) {
                    umma_f16_2sm(d_tmem, ra_lo + koff, rb_hi + koff, idesc, first ? 0u : 1u);
                    umma_f16_2sm(d_tmem, ra_hi + koff, rb_lo + koff, idesc, 1u);
                  } else {
                    umma_f16(d_tmem, ra_lo + koff, rb_hi + koff, idesc, first ? 0u : 1u);
                    umma_f16(d_tmem, ra_hi + koff, rb_lo + koff, idesc, 1u);
                  }
                  first = false;
                }
#pragma unroll
                for (int k = 0; k < kTcBlockK / 16; ++k) {
                  const uint64_t koff = static_cast<uint64_t>(k * 2);
                  if (CG == 2) umma_f16_2sm(d_tmem, ra_hi + koff, rb_hi + koff, idesc, 1u);
                  else umma_f16(d_tmem, ra_hi + koff, rb_hi + koff, idesc, 1u);
                }
              }
              if (CG == 2) umma_commit_2sm(&empty_bar[s]); else umma_commit(&empty_bar[s]);
              continue;
            }
            const uint64_t a_hi = umma_smem_desc_sw128(st);
            const uint64_t a_lo = umma_smem_desc_sw128(st + Cfg::kABytes);
            const uint64_t b_hi = umma_smem_desc_sw128(st + 2 * Cfg::kABytes);
            const uint64_t b_lo = umma_smem_desc_sw128(st + 2 * Cfg::kABytes + Cfg::kBBytes);
            // K advance inside the 128-byte swizzled row: 16 fp16 = 32 bytes = +2 in 16-byte units.
            // Cross terms first (tiny magnitudes, truncation negligible), dominant hi*hi products last.
#pragma unroll
            for (int k = 0; k < kTcBlockK / 16; ++k) {
              const uint64_t koff = static_cast<uint64_t>(k * 2);
              if (CG == 2) {
                umma_f16_2sm(d_tmem, a_lo + koff, b_hi + koff, idesc, first ? 0u : 1u);
                umma_f16_2sm(d_tmem, a_hi + koff, b_lo + koff, idesc, 1u);
              } else {
                umma_f16(d_tmem, a_lo + koff, b_hi + koff, idesc, first ? 0u : 1u);
                umma_f16(d_tmem, a_hi + koff, b_lo + koff, idesc, 1u);
              }
              first = false;
            }
#pragma unroll
            for (int k = 0; k < kTcBlockK / 16; ++k) {
              const uint64_t koff = static_cast<uint64_t>(k * 2);
              if (CG == 2) umma_f16_2sm(d_tmem, a_hi + koff, b_hi + koff, idesc, 1u);
              else umma_f16(d_tmem, a_hi + koff, b_hi + koff, idesc, 1u);
            }
            // frees this smem stage (in both CTAs of a pair) once the MMAs have read it
            if (CG == 2) umma_commit_2sm(&empty_bar[s]); else umma_commit(&empty_bar[s]);
          }
          // partial sum complete -> drain warps (of both CTAs)
          if (CG == 2) umma_commit_2sm(&acc_full_bar[buf]); else umma_commit(&acc_full_bar[buf]);
        }
        u += kb1 - kb0;
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;" ::: "memory");
    // ===================== drain + epilogue (warps 4..11: 4 TMEM lane quarters x 2 column halves) =========
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;       // which BLOCK_N/2 column range
    const int col0 = half * CPW;
    const int row = q * 32 + lane;          // accumulator row == pixel index inside the tile box
    const int te = threadIdx.x - 128;       // 0..255 among the drain threads
    const uint32_t acc_empty_base =
        (CG == 2) ? mapa_u32(smem_u32(&acc_empty_bar[0]), 0) : smem_u32(&acc_empty_bar[0]);
    const int osf = p.up2 ? 2 : 1;
    // stream-K scratch of THIS CTA: [128 rows][BLOCK_N] fp32 + one flag
    // drained partials are de-biased and un-scaled (weights were pre-scaled by a power of two) in one FMA
    const float pscale = p.partial_scale * __ldg(p.w_inv_scale);
    float* my_partial = p.sk_partials + static_cast<long long>(blockIdx.x) * (kTcBlockM * 256);
    int jc = 0;
    bool clamped = false;                   // a split-plane output had to be clamped to +-65504 (reported once at the end)
    int gn_tiles = 0;                       // fused-GroupNorm epilogues done (exchange buffer / barrier phase)
    float2 own_slab = make_float2(0.f, 0.f);

    for (long long u = u_begin; u < u_end;) {
      const int tile = static_cast<int>(u / nkb);
      const int kb0 = static_cast<int>(u - static_cast<long long>(tile) * nkb);
      const int kb1 = static_cast<int>(min(static_cast<long long>(nkb), kb0 + (u_end - u)));
      const int nchunks = (kb1 - kb0 + drain - 1) / drain;
      // this tile's bias value for column `te`: requested now, consumed by the epilogue a whole K loop later
      float bias_reg = 0.f;
      if (kb0 == 0 && p.bias != nullptr && te < BLOCK_N)
        bias_reg = __ldg(p.bias + static_cast<int>((tile / p.m_groups) % p.n_tiles) * BLOCK_N + te);
      float acc[CPW];
#pragma unroll
      for (int i = 0; i < CPW; ++i) acc[i] = 0.f;

      for (int j = 0; j < nchunks; ++j, ++jc) {
        const int buf = jc % NB;
        mbar_wait(&acc_full_bar[buf], (jc / NB) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BLOCK_N + col0;
#pragma unroll
        for (int ch = 0; ch < CPW / 32; ++ch) {
          float v[32];
          tmem_ld_32x32(taddr + ch * 32, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[ch * 32 + i] = fmaf(v[i], pscale, acc[ch * 32 + i]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG == 2) mbar_arrive_cluster(acc_empty_base + buf * 8);
          else mbar_arrive(&acc_empty_bar[buf]);
        }
      }

      if (kb0 > 0) {
        // ---- this group's range started inside the tile: publish the partial sums, the tile's owner finishes it
        // scratch layout [CPW/4][256 drain threads] float4: writer and reader are the SAME thread index (same row / column
        // range in every CTA), so consecutive lanes touch consecutive 16 bytes (it was [row][column]: one line per lane)
        float4* pv = reinterpret_cast<float4*>(my_partial) + te;
#pragma unroll
        for (int i = 0; i < CPW; i += 4)
          __stcg(pv + (i / 4) * 256, make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]));
        __threadfence();
        named_bar_sync(1, 256);
        if (te == 0) {
          int* flag = p.sk_flags + blockIdx.x;
          asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(1) : "memory");
        }
      } else {
        if (kb1 < nkb) {
          // ---- owner of a tile that other groups helped with: add their partial sums (they processed them first)
          // helpers are the groups right after this one; all their flags are polled in parallel (one thread each), then
          // the partial tiles are added in group order (deterministic) with nothing but loads between them
          const long long tile_end = static_cast<long long>(tile + 1) * nkb;
          int helpers = 0;
          for (long long covered = static_cast<long long>(tile) * nkb + kb1; covered < tile_end; ++helpers)
            covered = min(tile_end, total_units * (group + helpers + 2) / ngroups);
          for (int h0 = 0; h0 < helpers; h0 += 256) {
            const int hn = min(256, helpers - h0);
            if (te < hn) {
              const int other_cta = (group + 1 + h0 + te) * CG + static_cast<int>(cta_rank);
              const int* flag = p.sk_flags + other_cta;
              int v = 0, spins = 0;
              do {
                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
                if (++spins > (1 << 28)) { printf("mf: stream-K flag timeout cta=%d waits for %d\n", blockIdx.x, other_cta); __trap(); }
              } while (v == 0);
            }
            named_bar_sync(1, 256);
            for (int h = 0; h < hn; ++h) {
              const int other_cta = (group + 1 + h0 + h) * CG + static_cast<int>(cta_rank);
              const float4* ov = reinterpret_cast<const float4*>(p.sk_partials + static_cast<long long>(other_cta) * (kTcBlockM * 256)) + te;
#pragma unroll
              for (int i = 0; i < CPW; i += 4) {
                const float4 o = __ldcg(ov + (i / 4) * 256);
                acc[i] += o.x; acc[i + 1] += o.y; acc[i + 2] += o.z; acc[i + 3] += o.w;
              }
            }
            named_bar_sync(1, 256);
            if (te < hn) p.sk_flags[(group + 1 + h0 + te) * CG + static_cast<int>(cta_rank)] = 0;   // re-arm (CUDA-graph replay safe)
          }
        }
        // ---- epilogue: registers -> (+bias, statistics, optional residual / vector) -> swizzled staging tile -> TMA store
        const TileCoord tc = decode_tile(tile);
        const int oa = tc.phase >> 1, ob = tc.phase & 1;
        const int nt = tc.nt;
        const int iw = row % p.bw;
        const int ih = (row / p.bw) % p.bh;
        const int in = row / (p.bw * p.bh);
        const int n = tc.n0 + in, h = tc.h0 + ih, w = tc.w0 + iw;
        const bool valid = n < p.N;
        const long long pix = (static_cast<long long>(n) * (p.H * osf) + h * osf + oa) * (p.W * osf) + w * osf + ob;
        const long long ooff = pix * p.Cout + nt * BLOCK_N + col0;
        const CUtensorMap* omap = &maps.o[tc.phase];
        uint8_t* stg = staging + half * Cfg::kStagingHalf;
        const uint32_t stg_u = smem_u32(stg);
        const bool issuer = (q == 0) && (lane == 0);        // one thread per column half owns the bulk store groups

        if (te < BLOCK_N) bias_s[te] = bias_reg;
        named_bar_sync(1, 256);
        const bool want_stats = (p.stats != nullptr) || GN;
        constexpr int G8 = BLOCK_N / 8;

        // ---- phase A: + bias, per-slab (sum, sumsq) over this warp's 32 rows
#pragma unroll
        for (int ch = 0; ch < CPW / 32; ++ch) {
          float* v = &acc[ch * 32];
          {
            const float4* b4 = reinterpret_cast<const float4*>(bias_s + col0 + ch * 32);   // broadcast LDS
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
              const float4 b = b4[jj];
              v[4 * jj + 0] += b.x; v[4 * jj + 1] += b.y; v[4 * jj + 2] += b.z; v[4 * jj + 3] += b.w;
            }
          }
        }
        if (want_stats) {
          // V = 2 values (sum, sumsq) per 8-channel slab, SV slabs at a time; reduced over the warp's 32 rows by recursive
          // halving (lane L ends up with value L >> (5 - log2 V) of the group): V + ... shuffles instead of 5 per value (the
          // butterfly all-reduce cost 160 SHFL + 160 FADD per thread and tile at N = 256), same pairwise sums, same bits
          constexpr int SV = CPW >= 64 ? 8 : 4;        // slabs per group
          constexpr int V = 2 * SV;
#pragma unroll
          for (int gp = 0; gp < CPW / (8 * SV); ++gp) {
            float x[V];
#pragma unroll
            for (int g = 0; g < SV; ++g) {
              float s = 0.f, ss = 0.f;
#pragma unroll
              for (int jj = 0; jj < 8; ++jj) {
                const float xv = acc[(gp * SV + g) * 8 + jj];
                s += xv;
                ss = fmaf(xv, xv, ss);
              }
              x[2 * g] = valid ? s : 0.f;
              x[2 * g + 1] = valid ? ss : 0.f;
            }
            int nv = V;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
              if (nv > 1) {
                nv >>= 1;
                const bool up = (lane & off) != 0;
#pragma unroll
                for (int j = 0; j < V / 2; ++j) {
                  if (j < nv) {
                    const float keep = up ? x[nv + j] : x[j], send = up ? x[j] : x[nv + j];
                    x[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                  }
                }
              } else {
                x[0] += __shfl_xor_sync(0xffffffffu, x[0], off);
              }
            }
            constexpr int kShift = (V == 16) ? 1 : 2;       // lanes per value after the halving steps
            if ((lane & ((1 << kShift) - 1)) == 0)
              red[((q * G8) + (col0 + gp * SV * 8) / 8) * 2 + (lane >> kShift)] = x[0];
          }
        }

        // ---- phase B: statistics out (separate GroupNorm kernel follows) or finalised here (fused GroupNorm)
        const int spt = kTcBlockM / p.rows_per_sample;   // samples per tile (1, 2 or 4)
        const int wps = 4 / spt;                         // lane quarters per sample
        if (p.stats != nullptr) {
          named_bar_sync(1, 256);
          if (te < G8 * spt) {
            const int j = te / G8, i = te - j * G8;
            float s = 0.f, ss = 0.f;
            for (int wq = j * wps; wq < (j + 1) * wps; ++wq) {
              s += red[(wq * G8 + i) * 2 + 0];
              ss += red[(wq * G8 + i) * 2 + 1];
            }
            const int ns = tc.n0 + (spt > 1 ? j : 0);
            const int chunk = (p.chunks_per_sample > 1) ? (tc.th * p.tiles_w + tc.tw) : 0;
            if (ns < p.N) {
              float* dst = p.stats + ((static_cast<long long>(ns) * p.chunks_per_sample + chunk) * (p.Cout / 8) +
                                      nt * G8 + i) * 2;
              dst[0] = s;
              dst[1] = ss;
            }
          }
          named_bar_sync(1, 256);   // red[] is reused by the next tile
        }
        if (GN) {
          // per (sample of the tile, group): mean / rstd.  Same partial sums as the separate kernel: fp32 per (sample,
          // 8-channel slab, 128-pixel chunk), combined in fp64 across the group's slabs (and the pair's two chunks).
          named_bar_sync(1, 256);
          const int slabs = p.gn_cpg / 8;
          const int ngrp = BLOCK_N / p.gn_cpg;
          float g_mean = 0.f, g_rstd = 0.f;
          if (p.gn_mode == 2) {
            // this CTA's per-slab sums -> own copy (red quarter 0) and the other CTA's exchange buffer, then one remote
            // arrival per slab; the tile parity double-buffers the exchange
            const int par = gn_tiles & 1;
            if (te < G8) {
              float s = 0.f, ss = 0.f;
              for (int wq = 0; wq < 4; ++wq) { s += red[(wq * G8 + te) * 2]; ss += red[(wq * G8 + te) * 2 + 1]; }
              const uint32_t peer = cta_rank ^ 1u;
              const uint32_t dst = mapa_u32(smem_u32(&xch[par * G8 + te]), peer);
              asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(dst), "f"(s), "f"(ss) : "memory");
              asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(
                               mapa_u32(smem_u32(&xch_bar[par]), peer)) : "memory");
              own_slab = make_float2(s, ss);
            }
            if (te < G8) {
              // wait for the peer's values (acquire at cluster scope), bounded
              uint32_t ok = 0, spins = 0;
              const uint32_t bar = smem_u32(&xch_bar[par]);
              const uint32_t parity = (gn_tiles >> 1) & 1;
              do {
                asm volatile("{\n\t.reg .pred p;\n\t"
                             "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
                             "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
                if (++spins > (1u << 26)) { printf("mf: GroupNorm exchange timeout cta=%d\n", blockIdx.x); __trap(); }
              } while (!ok);
              const float2 other = xch[par * G8 + te];
              // chunk order of the separate kernel: rank 0's rows first
              const float2 c0v = cta_rank == 0 ? own_slab : other, c1v = cta_rank == 0 ? other : own_slab;
              red[te * 2] = c0v.x; red[te * 2 + 1] = c0v.y;               // quarter 0 <- chunk 0
              red[(G8 + te) * 2] = c1v.x; red[(G8 + te) * 2 + 1] = c1v.y;   // quarter 1 <- chunk 1
            }
            named_bar_sync(1, 256);
            if (te < ngrp) {
              double ds = 0.0, dss = 0.0;
              for (int chk = 0; chk < 2; ++chk)
                for (int sl = 0; sl < slabs; ++sl) {
                  ds += static_cast<double>(red[(chk * G8 + te * slabs + sl) * 2]);
                  dss += static_cast<double>(red[(chk * G8 + te * slabs + sl) * 2 + 1]);
                }
              const double cnt = static_cast<double>(p.gn_cpg) * (2 * kTcBlockM);
              const double mean = ds / cnt;
              double var = dss / cnt - mean * mean;
              if (var < 0.0) var = 0.0;
              g_mean = static_cast<float>(mean);
              g_rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(p.gn_eps)));
            }
          } else if (te < spt * ngrp) {
            const int j = te / ngrp, gi = te - j * ngrp;
            double ds = 0.0, dss = 0.0;
            for (int sl = 0; sl < slabs; ++sl) {
              float s = 0.f, ss = 0.f;
              for (int wq = j * wps; wq < (j + 1) * wps; ++wq) {
                s += red[(wq * G8 + gi * slabs + sl) * 2];
                ss += red[(wq * G8 + gi * slabs + sl) * 2 + 1];
              }
              ds += static_cast<double>(s);
              dss += static_cast<double>(ss);
            }
            const double cnt = static_cast<double>(p.gn_cpg) * p.rows_per_sample;
            const double mean = ds / cnt;
            double var = dss / cnt - mean * mean;
            if (var < 0.0) var = 0.0;
            g_mean = static_cast<float>(mean);
            g_rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(p.gn_eps)));
          }
          named_bar_sync(1, 256);                // everybody has read red[]: it now carries (mean, rstd) per (sample, group)
          if (te < (p.gn_mode == 2 ? 1 : spt) * ngrp) { red[te * 2] = g_mean; red[te * 2 + 1] = g_rstd; }
          named_bar_sync(1, 256);
          ++gn_tiles;
        }

        // ---- phase C: (GroupNorm + Swish + residual + embedding | residual / vector of the attention convs) -> staging
        //      tile -> TMA store
        const int srow_sample = (p.gn_mode == 1) ? (row / p.rows_per_sample) : 0;
        // GN: NOT unrolled — the body is ~1.5 k instructions and runs once per tile; unrolled 4x it was 108 KB of
        // straight-line code whose instruction fetch (stall_no_inst) cost more than its arithmetic
        // (profiles/r02_gn_fusion.md).  The chunk's 32 accumulators are selected into a fixed register window instead.
#pragma unroll(GN ? 1 : 4)
        for (int ch = 0; ch < CPW / 32; ++ch) {
          float v[32];
#pragma unroll
          for (int c2 = 0; c2 < CPW / 32; ++c2)
            if (c2 == ch) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = acc[c2 * 32 + i];
            }
          const int cbase = nt * BLOCK_N + col0 + ch * 32;       // first output channel of this 32-column chunk
          if (GN) {
            const int ngrp = BLOCK_N / p.gn_cpg;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int gi = (col0 + ch * 32 + g * 8) / p.gn_cpg;
              const float mean = red[(srow_sample * ngrp + gi) * 2], rstd = red[(srow_sample * ngrp + gi) * 2 + 1];
              const float4 ga0 = __ldg(reinterpret_cast<const float4*>(p.gn_gamma + cbase + g * 8));
              const float4 ga1 = __ldg(reinterpret_cast<const float4*>(p.gn_gamma + cbase + g * 8 + 4));
              const float4 be0 = __ldg(reinterpret_cast<const float4*>(p.gn_beta + cbase + g * 8));
              const float4 be1 = __ldg(reinterpret_cast<const float4*>(p.gn_beta + cbase + g * 8 + 4));
              const float ga[8] = {ga0.x, ga0.y, ga0.z, ga0.w, ga1.x, ga1.y, ga1.z, ga1.w};
              const float be[8] = {be0.x, be0.y, be0.z, be0.w, be1.x, be1.y, be1.z, be1.w};
#pragma unroll
              for (int jj = 0; jj < 8; ++jj) {
                // same operation order as gn_apply_fused_kernel: fma(x - mean, rstd * gamma, beta), Swish with MUFU division
                const float y = fmaf(v[8 * g + jj] - mean, rstd * ga[jj], be[jj]);
                v[8 * g + jj] = __fdividef(y, 1.0f + expf(-y));
              }
            }
            if (valid && p.gn_res_kind != 0) {
              const long long roff = ooff + ch * 32;
#pragma unroll
              for (int jj = 0; jj < 8; ++jj) {
                float4 r;
                if (p.gn_res_kind == 1) {
                  const __half* rh = reinterpret_cast<const __half*>(p.gn_res) + roff + 4 * jj;
                  r = ld_join4(rh, rh + p.gn_res_plane);
                } else {
                  r = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.gn_res) + roff + 4 * jj);
                }
                v[4 * jj + 0] += r.x; v[4 * jj + 1] += r.y; v[4 * jj + 2] += r.z; v[4 * jj + 3] += r.w;
              }
            }
            if (valid && p.gn_emb != nullptr) {
              const long long er = (p.gn_emb_index ? p.gn_emb_index[n] : static_cast<long long>(n)) * p.gn_emb_stride;
              const float4* e4 = reinterpret_cast<const float4*>(p.gn_emb + er + cbase);
#pragma unroll
              for (int jj = 0; jj < 8; ++jj) {
                const float4 e = __ldg(e4 + jj);
                v[4 * jj + 0] += e.x; v[4 * jj + 1] += e.y; v[4 * jj + 2] += e.z; v[4 * jj + 3] += e.w;
              }
            }
          }
          if (valid && p.res_kind != 0) {
            // fused residual add: out = conv + bias + residual(same pixel, same channels)   (attention blocks)
            const long long roff = ooff + ch * 32;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
              float4 r;
              if (p.res_kind == 1) {
                const __half* rh = reinterpret_cast<const __half*>(p.res) + roff + 4 * jj;
                r = ld_join4(rh, rh + p.res_plane);
              } else {
                r = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.res) + roff + 4 * jj);
              }
              v[4 * jj + 0] += r.x; v[4 * jj + 1] += r.y; v[4 * jj + 2] += r.z; v[4 * jj + 3] += r.w;
            }
          }
          if (valid && p.emb != nullptr) {
            // per-sample channel vector (one-token cross-attention collapses to this, attention_blocks.py:160-195)
            const float4* e4 = reinterpret_cast<const float4*>(p.emb + static_cast<long long>(n) * p.emb_stride + cbase);
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
              const float4 e = __ldg(e4 + jj);
              v[4 * jj + 0] += e.x; v[4 * jj + 1] += e.y; v[4 * jj + 2] += e.z; v[4 * jj + 3] += e.w;
            }
          }
          // the previous bulk store of this half must have finished reading the staging tile
          if (issuer) tma_store_wait_read();
          named_bar_sync(2 + half, 128);
          if (p.out_mode == kOutRaw) {
            // [128 rows][32 fp32] with the 128-byte swizzle of the store's tensor map: conflict-free 16-byte stores
            const uint32_t srow = stg_u + row * 128;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj)
              st_shared_v4(srow + ((jj ^ (row & 7)) << 4), __float_as_uint(v[4 * jj]), __float_as_uint(v[4 * jj + 1]),
                           __float_as_uint(v[4 * jj + 2]), __float_as_uint(v[4 * jj + 3]));
          } else {
            // hi tile [128 rows][32 fp16] then lo tile, 64-byte swizzle
            const uint32_t srow = stg_u + row * 64;
            const int sw = (row >> 1) & 3;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              uint32_t hi[4], lo[4];
#pragma unroll
              for (int k2 = 0; k2 < 4; ++k2) split16x2_flag(v[8 * jj + 2 * k2], v[8 * jj + 2 * k2 + 1], hi[k2], lo[k2], clamped);
              st_shared_v4(srow + ((jj ^ sw) << 4), hi[0], hi[1], hi[2], hi[3]);
              st_shared_v4(srow + Cfg::kStagingHalf / 2 + ((jj ^ sw) << 4), lo[0], lo[1], lo[2], lo[3]);
            }
          }
          fence_proxy_async_smem();
          named_bar_sync(2 + half, 128);
          if (issuer) {
            tma_store_5d(omap, stg, cbase, tc.w0, tc.h0, tc.n0, 0);
            if (p.out_mode != kOutRaw) tma_store_5d(omap, stg + Cfg::kStagingHalf / 2, cbase, tc.w0, tc.h0, tc.n0, 1);
            tma_store_commit();
          }
        }
        if (GN) named_bar_sync(1, 256);   // red[] (mean / rstd) is overwritten by the next tile's statistics
      }
      u += kb1 - kb0;
    }
    sat16_report(clamped);                                // values beyond the fp16 range of the split planes (sticky counter)
    if ((q == 0) && (lane == 0)) tma_store_wait_all();   // bulk stores issued by this thread have been written
  }

  // ---- teardown ---------------------------------------------------------------------------------
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    __syncwarp();
    if (CG == 2) tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
    else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// =================================================================================================
// Host side
// =================================================================================================
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

static int encode_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_b,
                      const cuuint32_t* box, const cuuint32_t* estr,
                      CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                      CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return 3;
  }
  CUresult r = fn(m, dtype, rank, const_cast<void*>(base), dims, strides_b, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)));
    return 3;
  }
  return 0;
}

// stream-K scratch: one [128][256] fp32 partial tile and one flag per CTA of the persistent grid
int streamk_scratch_alloc(StreamKScratch* sc) {
  int dev = 0, sms = 0;
  MF_CUDA_OK(cudaGetDevice(&dev));
  MF_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  sc->max_ctas = sms;
  sc->device = dev;
  MF_CUDA_OK(cudaMalloc(&sc->partials, static_cast<size_t>(sms) * kTcBlockM * 256 * sizeof(float)));
  MF_CUDA_OK(cudaMalloc(&sc->flags, static_cast<size_t>(sms) * sizeof(int)));
  MF_CUDA_OK(cudaMemset(sc->flags, 0, static_cast<size_t>(sms) * sizeof(int)));
  return 0;
}
void streamk_scratch_free(StreamKScratch* sc) {
  if (sc->partials) cudaFree(sc->partials);
  if (sc->flags) cudaFree(sc->flags);
  *sc = StreamKScratch();
}
// library-owned scratch for stand-alone mf_op_* launches: one per device (never shared across devices; launches that
// use it must be stream-ordered with each other, which the single-stream test surface is)
static int get_streamk_scratch(const StreamKScratch** out) {
  static std::mutex mu;
  static StreamKScratch per_device[64];
  std::lock_guard<std::mutex> lock(mu);
  int dev = 0;
  MF_CUDA_OK(cudaGetDevice(&dev));
  MF_REQUIRE(dev >= 0 && dev < 64, "device index out of range");
  if (per_device[dev].partials == nullptr) {
    int rc = streamk_scratch_alloc(&per_device[dev]);
    if (rc) return rc;
  }
  *out = &per_device[dev];
  return 0;
}

static bool pick_box(int H, int W, int* bw, int* bh, int* bn) {
  if (W >= kTcBlockM) {
    if (W % kTcBlockM) return false;
    *bw = kTcBlockM; *bh = 1; *bn = 1;
    return true;
  }
  if (W <= 0 || kTcBlockM % W) return false;
  const int rows = kTcBlockM / W;
  *bw = W;
  if (H >= rows) {
    if (H % rows) return false;
    *bh = rows; *bn = 1;
  } else {
    if (rows % H) return false;
    *bh = H; *bn = rows / H;
  }
  return true;
}

int conv_tc_stats_chunks(int H, int W) {
  const int hw = H * W;
  return hw >= kTcBlockM ? hw / kTcBlockM : 1;
}

// N, H, W: OUTPUT geometry
int conv_tc_supported(int N, int H, int W, int C0, int C1, int Cout, int ksize, int stride) {
  int bw, bh, bn;
  if (stride != 1 && stride != 2) return 0;
  if (stride == 2 && (ksize != 3 || C1 != 0)) return 0;
  if (ksize != 1 && ksize != 3) return 0;
  if (C0 <= 0 || C0 % kTcBlockK || C1 % kTcBlockK || Cout % 64) return 0;
  if (!pick_box(H, W, &bw, &bh, &bn)) return 0;
  const int hw = H * W;
  if (hw < 32 || (hw < kTcBlockM && kTcBlockM % hw) || (hw >= kTcBlockM && hw % kTcBlockM)) return 0;
  (void)N;
  return 1;
}

// Fused GroupNorm epilogue: the finalising CTA (H*W <= 128) or CTA pair (H*W == 256) must hold whole samples, the
// widest tile the channel count allows must hold whole groups of a multiple of 8 channels.
int conv_tc_gn_fusable(int H, int W, int Cout, int groups) {
  if (!g_fuse_gn || groups <= 0 || Cout % groups) return 0;
  const int cpg = Cout / groups;
  if (cpg % 8) return 0;
  int bn = 256;
  while (bn > 64 && Cout % bn) bn /= 2;
  if (Cout % bn || bn % cpg) return 0;
  const int hw = H * W;
  if (hw == 2 * kTcBlockM) return 2;
  if (hw >= 32 && hw <= kTcBlockM && kTcBlockM % hw == 0) return 1;
  return 0;
}

// View of an NHWC split tensor sub-sampled by `sub` in H and W starting at (ph, pw): sub = 1 is the tensor itself,
// sub = 2 selects one of the four input-parity grids a stride-2 convolution reads.
static int encode_act_map(CUtensorMap* m, const __half* base, long long plane, int N, int H, int W, int C, int bw,
                          int bh, int bn, int sub = 1, int ph = 0, int pw = 0) {
  base += (static_cast<long long>(ph) * W + pw) * C;
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)(W / sub), (cuuint64_t)(H / sub), (cuuint64_t)N, 2};
  cuuint64_t strides[4] = {(cuuint64_t)sub * C * 2, (cuuint64_t)sub * W * C * 2, (cuuint64_t)H * W * C * 2,
                           (cuuint64_t)plane * 2};
  cuuint32_t box[5] = {(cuuint32_t)kTcBlockK, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  return encode_map(m, base, 5, dims, strides, box, estr);
}

// Output view for the epilogue's TMA stores: NHWC [N, Ho, Wo, C] sub-sampled by `sub` starting at pixel (ph, pw)
// (sub = 2: the output-parity grid one phase of the folded upsample writes).  32-channel boxes of 128 pixels.
static int encode_out_map(CUtensorMap* m, void* base, int out_mode, long long plane, int N, int Ho, int Wo, int C, int bw,
                          int bh, int bn, int sub = 1, int ph = 0, int pw = 0) {
  const bool raw = out_mode == kOutRaw;
  const size_t es = raw ? 4 : 2;
  char* b = static_cast<char*>(base) + (static_cast<size_t>(ph) * Wo + pw) * C * es;
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)(Wo / sub), (cuuint64_t)(Ho / sub), (cuuint64_t)N, raw ? 1u : 2u};
  const cuuint64_t plane_b = raw ? (cuuint64_t)N * Ho * Wo * C * es : (cuuint64_t)plane * es;
  cuuint64_t strides[4] = {(cuuint64_t)sub * C * es, (cuuint64_t)sub * Wo * C * es, (cuuint64_t)Ho * Wo * C * es, plane_b};
  cuuint32_t box[5] = {32, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  return encode_map(m, b, 5, dims, strides, box, estr, raw ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                    raw ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
}

// Everything conv_tc_build decides about the SCHEDULE of a layer, as a pure function of the shape and the knobs (no device
// access: exported as mf_op_conv_tc_plan so that the heuristics are testable without a GPU).
//   m_tiles: 128-pixel tiles of the output; (bw, bh, bn_box): the pixel box of a tile; max_ctas: SMs of the device.
int conv_tc_plan_shape(int m_tiles, int bw, int bh, int bn_box, int C0, int C1, int Cout, int ksize, int stride, int up2,
                       int gn_mode, int forced_cg, int forced_bn, int drain_interval, int max_ctas, ConvTcShapePlan* out) {
  int cg = forced_cg > 0 ? forced_cg : g_default_cta_group;
  if (cg != 1 && cg != 2) cg = (m_tiles >= 2) ? 2 : 1;  // auto
  if (gn_mode == 2) cg = 2;                             // a sample = the two tiles of one CTA pair
  const int m_groups = (m_tiles + cg - 1) / cg;
  const int phases = up2 ? 4 : 1;
  const int ntaps = up2 ? 4 : ksize * ksize;
  const int max_groups = std::max(1, max_ctas / cg);
  const long long nkb = static_cast<long long>(ntaps) * ((C0 + C1) / kTcBlockK);
  // persistent grid: one CTA group per SM (pair); stream-K splits the (tile, K block) units evenly over the groups.
  // With stream-K off every tile gets its own group (classic one-tile-per-CTA launch).
  // Small-batch fill: a layer with fewer tiles than HALF the SM pairs (scripts/sample.py runs B = 4..16: the 8x8 level has
  // 4-16 tiles of 256 x 256 for 74 pairs) would leave most of the GPU idle while a few SMs stream the whole weight tensor.
  // Two levers, chosen by a small cost model (microseconds; constants from profiles/r02_small_batch.md): narrower tiles
  // (more of them, cheaper partial tiles, but a lower MMA rate: shared-memory-bound below N = 256) and sharing the K
  // range of a tile between several groups (each keeps >= min_kb K blocks; the owner adds the helpers' partial tiles,
  // which costs time per helper).  NOT done when the tiles fill at least half the pairs (B = 64): under the 1000 W cap the
  // idle SMs' power budget holds the busy ones' clock (profiles/r01_conv_tc_feed_probe.md).
  auto groups_for = [&](int bn_c, int min_kb) {
    const long long tiles = static_cast<long long>(m_groups) * (Cout / bn_c) * phases;
    long long g = std::min<long long>(tiles, max_groups);
    if (min_kb > 0 && 2 * tiles <= max_groups)
      g = std::min<long long>(max_groups, std::max<long long>(tiles, tiles * nkb / min_kb));
    return static_cast<int>(g);
  };
  auto cost_us = [&](int bn_c, int groups_c) {
    const long long tiles = static_cast<long long>(m_groups) * (Cout / bn_c) * phases;
    const double t_kb = bn_c == 256 ? 1.32 : (bn_c == 128 ? 0.94 : 0.72);       // one 64-channel K block of a 256 x bn tile
    const double per_group = std::ceil(static_cast<double>(tiles * nkb) / groups_c);
    const double helpers = std::ceil(static_cast<double>(groups_c) / static_cast<double>(tiles)) - 1.0;
    return per_group * t_kb + helpers * (0.5 + 1.5 * bn_c / 256.0) + (helpers > 0 ? 2.0 : 0.0);
  };
  int bn = forced_bn > 0 ? forced_bn : g_default_block_n;
  const bool bn_auto = (bn != 64 && bn != 128 && bn != 256);
  if (bn_auto) bn = 256;                                // auto: widest tile the channel count allows ...
  if (gn_mode != 0) bn = 256;
  while (Cout % bn) bn /= 2;
  MF_REQUIRE(bn >= 8, "conv_tc: output channels must be a multiple of 64");
  int groups = static_cast<int>(static_cast<long long>(m_groups) * (Cout / bn) * phases);
  if (g_stream_k) {
    groups = groups_for(bn, 0);
    if (g_split_fill > 0 && gn_mode == 0) {
      double best = cost_us(bn, groups);
      const int bn_top = bn;
      for (int bn_c = bn_top; bn_c >= 64; bn_c /= 2) {  // ... unless the layer cannot fill the GPU with it
        if (Cout % bn_c || (!bn_auto && bn_c != bn_top)) continue;
        if (2LL * m_groups * (Cout / bn_top) * phases > max_groups) break;   // enough tiles: keep the default plan
        for (int min_kb : {g_split_fill * 4, g_split_fill * 2, g_split_fill}) {
          const int g_c = groups_for(bn_c, min_kb);
          const double c = cost_us(bn_c, g_c);
          if (c < best * 0.97) { best = c; bn = bn_c; groups = g_c; }
        }
      }
    }
  }
  out->cta_group = cg;
  out->block_n = bn;
  out->groups = groups;
  out->m_groups = m_groups;
  out->n_tiles = Cout / bn;
  out->num_tiles = m_groups * out->n_tiles * phases;
  out->nkb = static_cast<int>(nkb);
  // Row-patch mode: 3x3 stride-1 layers whose tile is 128 pixels of one image row (W % 128 == 0: the VAE's 128x128 and
  // 256x256 levels) and whose narrow tiles (N <= 128) make the activation re-reads the bottleneck.  A pipeline stage then
  // covers the taps of one input row (3, or 2 for a phase of a folded up-conv), one TMEM partial per stage.
  out->row3 = (g_row_patch && ksize == 3 && stride == 1 && bh == 1 && bw == kTcBlockM && bn_box == 1 && gn_mode == 0 && cg == 2 &&
               (bn == 64 || bn == 128) && drain_interval >= 3 && 2LL * out->num_tiles > max_groups && g_stream_k) ? 1 : 0;
  return 0;
}

// shape-only query behind mf_op_conv_tc_plan: out8 = {supported, block_n, cta_group, groups, row3, num_tiles, nkb, m_groups}
int conv_tc_plan_query(int N, int H, int W, int C0, int C1, int Cout, int ksize, int stride, int up2, int sm_count, int* out8) {
  for (int i = 0; i < 8; ++i) out8[i] = 0;
  stride = stride == 2 ? 2 : 1;
  if (H % stride || W % stride) return 0;
  const int Ho = H / stride, Wo = W / stride;
  if (!conv_tc_supported(N, Ho, Wo, C0, C1, Cout, ksize, stride)) return 0;
  int bw, bh, bnb;
  if (!pick_box(Ho, Wo, &bw, &bh, &bnb)) return 0;
  const int m_tiles = (Wo / bw) * (Ho / bh) * ((N + bnb - 1) / bnb);
  ConvTcShapePlan sp;
  const int rc = conv_tc_plan_shape(m_tiles, bw, bh, bnb, C0, C1, Cout, ksize, stride, up2, 0, 0, 0, g_default_drain_interval,
                                    sm_count > 0 ? sm_count : 148, &sp);
  if (rc) return rc;
  out8[0] = 1; out8[1] = sp.block_n; out8[2] = sp.cta_group; out8[3] = sp.groups; out8[4] = sp.row3; out8[5] = sp.num_tiles;
  out8[6] = sp.nkb; out8[7] = sp.m_groups;
  return 0;
}

int conv_tc_build(const ConvTcDesc& d, ConvTcPlan* plan) {
  const int stride = d.stride == 2 ? 2 : 1;
  MF_REQUIRE(!d.up2 || (stride == 1 && d.ksize == 3 && d.C1 == 0 && d.stats == nullptr),
             "folded upsample conv: 3x3, stride 1, single source, no GroupNorm statistics");
  MF_REQUIRE(d.H % stride == 0 && d.W % stride == 0, "stride-2 conv_tc needs even input height/width");
  const int Ho = d.H / stride, Wo = d.W / stride;
  MF_REQUIRE(conv_tc_supported(d.N, Ho, Wo, d.C0, d.C1, d.Cout, d.ksize, stride), "shape not supported by conv_tc");
  MF_REQUIRE((reinterpret_cast<uintptr_t>(d.src0) & 15) == 0 && (reinterpret_cast<uintptr_t>(d.w_planes) & 15) == 0,
             "TMA sources must be 16-byte aligned");
  ConvTcParams& p = plan->p;
  p.N = d.N; p.H = Ho; p.W = Wo;
  pick_box(Ho, Wo, &p.bw, &p.bh, &p.bn);
  p.tiles_w = Wo / p.bw;
  p.tiles_h = Ho / p.bh;
  p.tiles_n = (d.N + p.bn - 1) / p.bn;
  p.C0 = d.C0; p.C1 = d.C1; p.Cout = d.Cout;
  p.ntaps = d.up2 ? 4 : d.ksize * d.ksize;
  p.up2 = d.up2 ? 1 : 0;
  p.per_tap_map = stride == 2;
  for (int t = 0; t < p.ntaps; ++t) {
    const int r = t / d.ksize, sx = t % d.ksize;
    if (stride == 1) {
      p.dy[t] = r - d.ksize / 2;
      p.dx[t] = sx - d.ksize / 2;
      p.tap_map[t] = 0;
    } else {
      // input row 2*ho + r - 1:  r=0 -> odd grid row ho-1,  r=1 -> even grid row ho,  r=2 -> odd grid row ho
      const int phr = (r == 1) ? 0 : 1, pwc = (sx == 1) ? 0 : 1;
      p.dy[t] = (r == 0) ? -1 : 0;
      p.dx[t] = (sx == 0) ? -1 : 0;
      p.tap_map[t] = 2 * phr + pwc;
    }
  }
  p.drain_interval = d.drain_interval > 0 ? d.drain_interval : g_default_drain_interval;
  // Each TMEM partial sum went through ~4*drain (+8*(drain-1)) truncating (round-toward-zero) accumulations, i.e. it
  // is biased towards zero by about half an ulp per accumulation.  Multiplying it by (1 + eps) on the way into the
  // round-to-nearest register sum removes the mean of that bias for free (the add becomes an FMA).
  // Calibrated on B200 against fp64 (profiles/r01_debias_calibration.jsonl): the relative bias of a drained partial is
  // -0.87e-7 / -2.7e-7 / -6.3e-7 for drain = 1 / 2 / 4, the same for Gaussian, all-positive and Swish-like operands.
  auto set_partial_scale = [&](int kblocks_per_partial) {
    if (g_debias_eps_per_kblock < 0.f) {
      const long ulps = std::max(1L, std::lround(0.73 * std::pow(static_cast<double>(kblocks_per_partial), 1.43)));
      p.partial_scale = 1.0f + static_cast<float>(ulps) * 1.1920929e-7f;
    } else {
      p.partial_scale = 1.0f + g_debias_eps_per_kblock * static_cast<float>(kblocks_per_partial);
    }
  };
  set_partial_scale(p.drain_interval);
  p.bias = d.bias;
  p.out = d.out; p.out_plane = d.out_plane; p.out_mode = d.out_mode;
  p.stats = d.stats;
  p.res = d.res; p.res_plane = d.res_plane; p.res_kind = d.res ? d.res_kind : 0;
  MF_REQUIRE(d.w_inv_scale != nullptr, "conv_tc needs the weight scale produced by prep_weight_tc");
  p.w_inv_scale = d.w_inv_scale;
  p.emb = d.emb; p.emb_stride = d.emb_stride;
  MF_REQUIRE(!(d.up2 && (d.res || d.emb)), "folded upsample conv has no fused residual");
  p.chunks_per_sample = conv_tc_stats_chunks(Ho, Wo);
  p.rows_per_sample = Ho * Wo >= kTcBlockM ? kTcBlockM : Ho * Wo;
  p.gn_mode = 0;
  if (d.gn_groups > 0) {
    p.gn_mode = conv_tc_gn_fusable(Ho, Wo, d.Cout, d.gn_groups);
    MF_REQUIRE(p.gn_mode != 0, "fused GroupNorm epilogue requested for a geometry that does not allow it");
    MF_REQUIRE(!d.up2 && d.stats == nullptr && d.res == nullptr && d.emb == nullptr && d.out_mode == kOutSplit &&
                   d.gn_gamma != nullptr && d.gn_beta != nullptr,
               "fused GroupNorm epilogue: split output, no separate statistics / attention residual");
    p.gn_cpg = d.Cout / d.gn_groups;
    p.gn_eps = d.gn_eps;
    p.gn_gamma = d.gn_gamma; p.gn_beta = d.gn_beta;
    p.gn_res = d.gn_res; p.gn_res_plane = d.gn_res_plane; p.gn_res_kind = d.gn_res ? d.gn_res_kind : 0;
    p.gn_emb = d.gn_emb; p.gn_emb_stride = d.gn_emb_stride; p.gn_emb_index = nullptr;
  }

  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const StreamKScratch* scp = d.scratch;
  if (scp == nullptr) {
    int rcs = get_streamk_scratch(&scp);
    if (rcs) return rcs;
  }
  const StreamKScratch& sc = *scp;
  p.sk_partials = sc.partials;
  p.sk_flags = sc.flags;
  ConvTcShapePlan sp;
  {
    const int rcp = conv_tc_plan_shape(m_tiles, p.bw, p.bh, p.bn, d.C0, d.C1, d.Cout, d.ksize, stride, d.up2 ? 1 : 0, p.gn_mode,
                                       d.cta_group, d.block_n, p.drain_interval, sc.max_ctas, &sp);
    if (rcp) return rcp;
  }
  const int cg = sp.cta_group, bn = sp.block_n, groups = sp.groups;
  plan->cta_group = cg;
  MF_REQUIRE(p.gn_mode == 0 || bn % p.gn_cpg == 0, "fused GroupNorm: a tile must hold whole groups");
  MF_REQUIRE(p.gn_mode != 2 || m_tiles % 2 == 0, "fused GroupNorm (pair mode): even tile count");
  plan->block_n = bn;
  // a CTA pair owns two consecutive M tiles; an odd tile count gets a padding tile (all loads out of range -> zeros,
  // all stores masked)
  p.m_groups = sp.m_groups;
  p.n_tiles = sp.n_tiles;
  p.num_tiles = sp.num_tiles;
  plan->row3 = sp.row3;
  if (sp.row3) {
    const int rt = d.up2 ? 2 : 3;             // taps per input row (a phase of the folded upsample has 2x2 taps)
    p.ntaps = rt;                             // K blocks are now (channel slab, input row)
    set_partial_scale(rt * (p.drain_interval / rt));
    p.drain_interval /= rt;
  }
  MF_REQUIRE(groups * cg <= sc.max_ctas || !g_stream_k, "stream-K grid exceeds the scratch allocation");
  plan->grid = dim3(groups * cg, 1, 1);

  int rc = 0;
  if (stride == 2) {
    for (int ph = 0; ph < 2 && rc == 0; ++ph)
      for (int pw = 0; pw < 2 && rc == 0; ++pw)
        rc = encode_act_map(&plan->maps.a[2 * ph + pw], d.src0, d.src0_plane, d.N, d.H, d.W, d.C0, p.bw, p.bh, p.bn, 2,
                            ph, pw);
    if (rc) return rc;
  } else {
    const int abw = plan->row3 ? p.bw + 2 : p.bw;   // row-patch mode: the tile's row plus one halo pixel on each side
    rc = encode_act_map(&plan->maps.a[0], d.src0, d.src0_plane, d.N, d.H, d.W, d.C0, abw, p.bh, p.bn);
    if (rc) return rc;
    if (d.C1 > 0) {
      MF_REQUIRE((reinterpret_cast<uintptr_t>(d.src1) & 15) == 0, "TMA sources must be 16-byte aligned");
      rc = encode_act_map(&plan->maps.a[1], d.src1, d.src1_plane, d.N, d.H, d.W, d.C1, abw, p.bh, p.bn);
      if (rc) return rc;
    } else {
      plan->maps.a[1] = plan->maps.a[0];
    }
    plan->maps.a[2] = plan->maps.a[0];
    plan->maps.a[3] = plan->maps.a[0];
  }
  MF_REQUIRE((reinterpret_cast<uintptr_t>(d.out) & 15) == 0, "TMA store target must be 16-byte aligned");
  if (d.up2) {
    for (int ph = 0; ph < 2 && rc == 0; ++ph)
      for (int pw = 0; pw < 2 && rc == 0; ++pw)
        rc = encode_out_map(&plan->maps.o[2 * ph + pw], d.out, d.out_mode, d.out_plane, d.N, 2 * Ho, 2 * Wo, d.Cout, p.bw,
                            p.bh, p.bn, 2, ph, pw);
  } else {
    rc = encode_out_map(&plan->maps.o[0], d.out, d.out_mode, d.out_plane, d.N, Ho, Wo, d.Cout, p.bw, p.bh, p.bn);
    for (int i = 1; i < 4; ++i) plan->maps.o[i] = plan->maps.o[0];
  }
  if (rc) return rc;
  const long long K = static_cast<long long>(d.up2 ? 4 : d.ksize * d.ksize) * (d.C0 + d.C1);   // NOT p.ntaps: 3 row taps in row-patch mode
  const long long wrows = static_cast<long long>(d.Cout) * (d.up2 ? 4 : 1);  // up2: four phase matrices stacked
  cuuint64_t wd[3] = {(cuuint64_t)K, (cuuint64_t)wrows, 2};
  cuuint64_t ws[2] = {(cuuint64_t)K * 2, (cuuint64_t)K * wrows * 2};
  cuuint32_t wb[3] = {(cuuint32_t)kTcBlockK, (cuuint32_t)(plan->block_n / plan->cta_group), 1};
  cuuint32_t we[3] = {1, 1, 1};
  return encode_map(&plan->maps.w, d.w_planes, 3, wd, ws, wb, we);
}

template <int BLOCK_N, int CG, bool GN, bool ROW3 = false>
static int launch_t(const ConvTcPlan& plan, cudaStream_t stream, const ConvTcParams& params) {
  static bool attr_set = false;  // per template instantiation
  if (!attr_set) {
    MF_CUDA_OK(cudaFuncSetAttribute(conv_tc_kernel<BLOCK_N, CG, GN, ROW3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    TcCfg<BLOCK_N, CG, ROW3>::kSmemBytes));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = plan.grid;
  cfg.blockDim = dim3(kTcThreads, 1, 1);
  cfg.dynamicSmemBytes = TcCfg<BLOCK_N, CG, ROW3>::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl ? 2 : 1;
  MF_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_tc_kernel<BLOCK_N, CG, GN, ROW3>, plan.maps, params));
  return 0;
}

int conv_tc_launch(const ConvTcPlan& plan, cudaStream_t stream, int emb_dedup, const long long* emb_index) {
  ConvTcParams p = plan.p;
  if (emb_dedup && p.gn_emb != nullptr) {   // deduplicated embedding rows: one per class (index) or a single shared row
    p.gn_emb_index = emb_index;
    if (emb_index == nullptr) p.gn_emb_stride = 0;
  }
  if (p.gn_mode != 0) {
    switch (plan.block_n * 10 + plan.cta_group) {
      case 2561: return launch_t<256, 1, true>(plan, stream, p);
      case 2562: return launch_t<256, 2, true>(plan, stream, p);
      case 1281: return launch_t<128, 1, true>(plan, stream, p);
      case 1282: return launch_t<128, 2, true>(plan, stream, p);
      case 641: return launch_t<64, 1, true>(plan, stream, p);
      case 642: return launch_t<64, 2, true>(plan, stream, p);
    }
  }
  if (plan.row3) {
    if (plan.block_n == 64 && plan.cta_group == 2) return launch_t<64, 2, false, true>(plan, stream, p);
    if (plan.block_n == 128 && plan.cta_group == 2) return launch_t<128, 2, false, true>(plan, stream, p);
    set_error("conv_tc_launch: row-patch plan with an unsupported tile");
    return 2;
  }
  switch (plan.block_n * 10 + plan.cta_group) {
    case 2561: return launch_t<256, 1, false>(plan, stream, p);
    case 2562: return launch_t<256, 2, false>(plan, stream, p);
    case 1281: return launch_t<128, 1, false>(plan, stream, p);
    case 1282: return launch_t<128, 2, false>(plan, stream, p);
    case 641: return launch_t<64, 1, false>(plan, stream, p);
    case 642: return launch_t<64, 2, false>(plan, stream, p);
  }
  set_error("conv_tc_launch: bad block_n / cta_group");
  return 2;
}

MF_DEFINE_SATURATION_READER(sat_read_conv_tc)

}  // namespace mf
