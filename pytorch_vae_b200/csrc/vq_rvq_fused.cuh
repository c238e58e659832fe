// Persistent residual-VQ forward (eval mode): ONE kernel walks all L levels of a row tile without leaving the SM.
//
// Reference: models/vq_vae.py:226-263 -- for level in range(L): distances of the residual to the level's codes,
// argmin, gather, residual -= code; then indices = cat(levels), z_q = sum of the gathered codes in level order,
// z_q_st = z + (z_q - z).  At the stage-2 shape (4 x 1024 codes, D = 512, 8192 rows per step) the level-by-level
// pipeline (pre-pass -> tcgen05 search -> re-rank -> hand-back -> unpack per level, then one finalize pass) is 22
// kernels of 3-20 us each whose durations are fill / drain and cold-cache latency, not work: 0.255 ms for 34 GFLOP.
//
// Here a CTA owns BM rows (128, or 64 when the batch has fewer 128-row tiles than the chip has SMs: the stage-2 batch
// of 8192 rows then runs on 128 SMs instead of 64).  Per level:
//   14 worker warps  build the 16-bit operand tile of the residual in shared memory (A, BM x D, 128-byte swizzle) and
//                    the rows' admission margins                                        (level 0: from z)
//   warp 1           issues tcgen05.mma M=BM N=BN K=16 against the level's codebook tiles (BN = 256 codes wherever the
//                    shared memory allows: an M=128 N=128 MMA needs 128 B/clk of operands and runs at ~60 %), streamed by
//   warp 0           with TMA through an mbarrier ring; accumulators (pre-loaded with -|e|^2/2) double-buffered in TMEM
//   8 of the workers scan the scores out of TMEM into per-row candidate records in SHARED memory (same admission rule
//                    and error bound as vq_search_tc.cu), then ONE THREAD PER ROW prunes its records against the final
//                    threshold: a single survivor = certified (the great majority), anything else goes on a list
//   14 worker warps  take the listed rows (dynamic grab): several survivors = exact fp64 re-rank of the fp32 residual
//                    against the fp32 code rows (loads of several candidates in flight together); list overflow /
//                    non-finite = exhaustive exact search of the level by the warp; then form the next residual
//                    fl(r - e) (fp32, kept in an L2-resident scratch tile of the CTA), convert it into the operand
//                    tile of the next level and compute its margin.  After the last level the same warps emit z_q
//                    (level-order sum), z_q_st, the squared error and the histogram.
// The fp32 residual never round-trips through HBM as a full tensor, nothing is launched between levels, and the
// codebook operand tiles come from L2.
#pragma once
#include <cstdio>
#include <cstdlib>

#include "tc_ptx.cuh"

namespace vqb {

constexpr int RQ_NEPI = 8;                 // scanning warps: 4 TMEM lane quarters x 2 column slices
constexpr int RQ_WORKERS = 14;             // warps 2..15
constexpr int RQ_THREADS = 64 + RQ_WORKERS * 32;
constexpr int RQ_RS = 8;                   // candidate records per (row, column slice); more -> exhaustive search
constexpr int RQ_LIST = 32;                // surviving codes per row scored in the re-rank; more -> exhaustive search
constexpr int RQ_MAXL = 8;

struct RvqParams {
  int64_t n_rows;
  int D, K_per, L, mode;
  int row_tiles, code_tiles, stages;
  uint32_t idesc;
  const float* z;               // [n_rows, D]
  const float* E;               // [L * K_per, D] fp32
  const uint16_t* E_lp;         // [L * K_per, D] 16-bit operand plane of the mode
  const float* ee_half;         // [L * K_per] plane of the mode
  const float* level_meta;      // [L][8]
  float* scratch;               // [gridDim.x][BM][D] fp32 residual tile of each CTA
  int64_t* idx_out;             // [L][n_rows] level-major global ids
  float* zq_out;
  float* zq_st_out;
  double* sqerr_sum;
  int* hist;                    // [L * K_per]
  int* counters;                // [0] rows that took the exhaustive search, [1] rows the one-thread prune did not certify
  long long* trace;             // VQB200_DEBUG=5: CTA 0, first row tile: [role 0..2][level][event < 8] clock64 stamps
  // training mode (SCATTER): the EMA segment sums of models/vq_vae.py:80-83, per level, zero on entry
  float* seg_sum;               // [L * K_per, D]  sum of the residual rows that chose the code
  float* seg_cnt;               // [L * K_per]     how many did
  // statistics tail (optional): the LAST CTA to finish turns the histogram into perplexity / dead ratio / mean squared
  // error (the stats_finalize kernel of the separate path) -- one launch fewer on a 0.14 ms forward
  float* stats_out;             // [3] or NULL
  float* ep_usage;              // [L * K_per] or NULL
  float* ep_cnt;                // [1] or NULL
  float count_add;
  double inv_elems;
};
// roles: 0 = MMA warp, 1 = first scanning warp (warp 4), 2 = first helper warp (warp 2)
#define RQ_TR(role, lvl, ev)                                                                               \
  do {                                                                                                     \
    if (p.trace && blockIdx.x == 0 && it == 0 && lane == 0) p.trace[(((role) * RQ_MAXL) + (lvl)) * 8 + (ev)] = clock64(); \
  } while (0)

__device__ __forceinline__ void rq_bar_workers() { asm volatile("bar.sync 1, %0;" ::"n"(RQ_WORKERS * 32) : "memory"); }

// bytes of the per-CTA bookkeeping that follows the operand tile and the codebook ring in shared memory
__host__ __device__ constexpr int rq_misc_bytes(int BM, int BN) {
  return BM * 4                      // margin_s
         + BM * 2 * 4                // cnt_s
         + BM * 2 * 4                // best_s
         + BM * 2 * RQ_RS * 8        // rec_s
         + RQ_NEPI * (BN / 2) * 4    // ee_slots
         + RQ_MAXL * BM * 4          // res_s
         + RQ_WORKERS * RQ_LIST * 4  // list_s
         + BM * 4                    // order_s
         + 32;                       // ctl_s
}

// 32 accumulator columns of one row -> candidate records {code group << 8 | admit mask, group max} in shared memory.
__device__ __forceinline__ void rq_scan(const uint32_t (&v)[32], uint32_t code0, float margin, float& best, float& thr,
                                        int& cnt, uint2* rec) {
  float gm[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const float* s = reinterpret_cast<const float*>(&v[g * 8]);
    gm[g] = fmaxf(fmaxf(fmaxf(s[0], s[1]), fmaxf(s[2], s[3])), fmaxf(fmaxf(s[4], s[5]), fmaxf(s[6], s[7])));
  }
  const float cm = fmaxf(fmaxf(gm[0], gm[1]), fmaxf(gm[2], gm[3]));
  if (cm >= thr) {
    best = fmaxf(best, cm);
    thr = best - margin;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      if (gm[g] >= thr) {
        uint32_t mk = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) mk |= (__uint_as_float(v[g * 8 + i]) >= thr) ? (1u << i) : 0u;
        if (cnt == RQ_RS) {
          // list exactly full (nothing lost yet): records whose group maximum has fallen below the running
          // threshold can never matter again (the threshold only rises) -- compact them away before giving up
          int keep = 0;
          for (int i = 0; i < RQ_RS; ++i) {
            const uint2 e = rec[i];
            if (__uint_as_float(e.y) >= thr) rec[keep++] = e;
          }
          cnt = keep;                                     // == RQ_RS if nothing could go: the row overflows for good
        }
        if (cnt < RQ_RS) rec[cnt] = make_uint2((((code0 >> 3) + g) << 8) | mk, __float_as_uint(gm[g]));
        ++cnt;
      }
    }
  }
}

// ---- one row of the tile: fp32 values (this lane's SL float4 slices) -> 16-bit operand row + margin.  Free
// __forceinline__ functions, not lambdas: an out-of-line call would pass the row by address and park every row buffer
// of the caller in local memory (each load then waits for its own store: measured 30 k cycles per level).
template <int SL, bool BF16, int BM>
__device__ __forceinline__ void rq_emit_operand_row(uint8_t* a_tile, float* margin_s, int r, const float4 (&v)[SL],
                                                    const float* meta, int lane) {
  constexpr int D = SL * 128;
  float ss = 0.f, sse = 0.f;
#pragma unroll
  for (int s = 0; s < SL; ++s) {
    const float4 x = v[s];
    uint2 pk;
    if (BF16) {
      const __nv_bfloat162 a = __floats2bfloat162_rn(x.x, x.y), b2 = __floats2bfloat162_rn(x.z, x.w);
      pk.x = *reinterpret_cast<const uint32_t*>(&a); pk.y = *reinterpret_cast<const uint32_t*>(&b2);
      const float f0 = __uint_as_float(pk.x << 16), f1 = __uint_as_float(pk.x & 0xffff0000u);
      const float f2 = __uint_as_float(pk.y << 16), f3 = __uint_as_float(pk.y & 0xffff0000u);
      ss += f0 * f0 + f1 * f1 + f2 * f2 + f3 * f3;
    } else {
      pk.x = f16x2_bits_flush(x.x, x.y); pk.y = f16x2_bits_flush(x.z, x.w);
      const float2 g0 = f16x2_bits_to_float2(pk.x), g1 = f16x2_bits_to_float2(pk.y);
      ss += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
      const float d0 = x.x - g0.x, d1 = x.y - g0.y, d2 = x.z - g1.x, d3 = x.w - g1.y;      // exact differences
      sse += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    }
    // element column = s * 128 + lane * 4: k-block slab (64 columns), 16-byte chunk inside the 128-byte row, half
    const int col = s * 128 + lane * 4;
    const uint32_t slab = col >> 6, chunk = (col & 63) >> 3, half = (col & 7) >> 2;
    uint8_t* dst = a_tile + slab * (BM * 128) + r * 128 + ((chunk ^ (r & 7)) << 4) + half * 8;
    *reinterpret_cast<uint2*>(dst) = pk;
  }
  ss = warp_sum(ss);
  sse = warp_sum(sse);
  if (lane == 0) {
    float m;
    if (BF16) m = 2.f * static_cast<float>(D + 32) * 1.1920929e-7f * (sqrtf(ss) * 1.0001f) * meta[2] + 1e-30f;
    else m = admission_margin_fp32(ss, sse, meta[0], meta[4], meta[5], D);
    if (meta[1] != 0.f || !(ss < __int_as_float(0x7f800000)) || !(sse < __int_as_float(0x7f800000)))
      m = __int_as_float(0x7fc00000);   // NaN: exhaustive search
    margin_s[r] = m;
  }
}
// one code row of the exact re-rank (fp32 codes; bf16_input mode: the rounded plane)
template <int SL, bool BF16>
__device__ __forceinline__ void rq_load_code_row(float4 (&ev)[SL], const float* E, const uint16_t* E_lp, int64_t gid, int lane) {
  constexpr int D = SL * 128;
#pragma unroll
  for (int s = 0; s < SL; ++s) {
    if (BF16) {
      const uint2 b = *reinterpret_cast<const uint2*>(E_lp + gid * D + s * 128 + lane * 4);
      ev[s] = make_float4(__uint_as_float(b.x << 16), __uint_as_float(b.x & 0xffff0000u),
                          __uint_as_float(b.y << 16), __uint_as_float(b.y & 0xffff0000u));
    } else {
      ev[s] = __ldg(reinterpret_cast<const float4*>(E + gid * D + s * 128 + lane * 4));
    }
  }
}
// exact score of the row against one code (fp64 accumulation; bf16_input mode: of the rounded values)
template <int SL, bool BF16>
__device__ __forceinline__ double rq_exact_score(const float4 (&v)[SL], const float4 (&ev)[SL]) {
  double dot = 0.0, ee = 0.0;
#pragma unroll
  for (int s = 0; s < SL; ++s) {
    float zv[4] = {v[s].x, v[s].y, v[s].z, v[s].w};
    const float e4[4] = {ev[s].x, ev[s].y, ev[s].z, ev[s].w};
    if (BF16) {
#pragma unroll
      for (int q = 0; q < 4; ++q) zv[q] = bf16_round(zv[q]);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      dot = fma(static_cast<double>(zv[q]), static_cast<double>(e4[q]), dot);
      ee = fma(static_cast<double>(e4[q]), static_cast<double>(e4[q]), ee);
    }
  }
  return warp_sum(dot - 0.5 * ee);                         // one reduction per candidate (bit-identical codes still tie exactly)
}
// exhaustive exact search of one level by the warp: d' = |e|^2/2 - r.e (fp32), packed (key, index) minimum -- the rule
// of search_simt_kernel (lowest index on ties, a NaN distance wins).  Rare: kept out of line; it re-loads the row.
template <int SL, bool BF16>
__device__ __noinline__ uint32_t rq_exhaustive(const float* row, const float* E, const uint16_t* E_lp, const float* ee_half,
                                               int K_per, int l, int lane) {
  uint64_t bestk = ~0ull;
  float4 x[SL];
#pragma unroll
  for (int s = 0; s < SL; ++s) {
    x[s] = reinterpret_cast<const float4*>(row)[s * 32 + lane];
    if (BF16) { x[s].x = bf16_round(x[s].x); x[s].y = bf16_round(x[s].y); x[s].z = bf16_round(x[s].z); x[s].w = bf16_round(x[s].w); }
  }
  constexpr int XB = SL <= 2 ? 4 : 2;                     // codes per step whose loads are in flight together
  for (int k0 = 0; k0 < K_per; k0 += XB) {
    float4 e4[XB][SL];
    float eeh[XB];
#pragma unroll
    for (int u = 0; u < XB; ++u) {
      const int k = k0 + u < K_per ? k0 + u : K_per - 1;
      const int64_t gid = static_cast<int64_t>(l) * K_per + k;
      eeh[u] = ee_half[gid];
      rq_load_code_row<SL, BF16>(e4[u], E, E_lp, gid, lane);
    }
    float dot[XB];
#pragma unroll
    for (int u = 0; u < XB; ++u) {
      dot[u] = 0.f;
#pragma unroll
      for (int s = 0; s < SL; ++s) {
        dot[u] = fmaf(x[s].x, e4[u][s].x, dot[u]); dot[u] = fmaf(x[s].y, e4[u][s].y, dot[u]);
        dot[u] = fmaf(x[s].z, e4[u][s].z, dot[u]); dot[u] = fmaf(x[s].w, e4[u][s].w, dot[u]);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int u = 0; u < XB; ++u) dot[u] += __shfl_xor_sync(0xffffffffu, dot[u], o);
    }
#pragma unroll
    for (int u = 0; u < XB; ++u)
      if (k0 + u < K_per) bestk = umin64(bestk, pack_minloc(eeh[u] - dot[u], static_cast<uint32_t>(k0 + u)));
  }
  return static_cast<uint32_t>(bestk & 0xffffffffull);
}

// Training mode (SCATTER).  The reference updates the codebook after EVERY level (models/vq_vae.py:251 -> :77-89),
// and each update touches ALL K_total codes: for the codes of another level it is a decay-only step (their one-hot
// columns are empty: cs <- cs g, es <- es g, E <- es / (cs + eps)).  Level l is therefore searched against codes that
// have seen l decay-only steps and nothing that depends on this batch -- so the codebook every level will be searched
// against is known BEFORE the forward (refresh phase 1, vq_rowops.cu), the forward itself runs exactly as in eval mode
// while it reduces the residual rows into the segment sums of their codes, and the real update of every level
// followed by its L - 1 - l trailing decay-only steps is applied afterwards in one pass (refresh phase 2): bit for
// bit the reference's sequence, with no synchronisation between levels.
template <int SL, bool BF16, int BM, int BN, bool SCATTER>
__global__ void __launch_bounds__(RQ_THREADS, 1)
rvq_fused_kernel(const __grid_constant__ CUtensorMap tmap_e, const RvqParams p) {
  constexpr int D = SL * 128;
  constexpr int KBLK = D / TC_KB;
  constexpr uint32_t A_BYTES = BM * D * 2;
  constexpr uint32_t STAGE_BYTES = BN * TC_KB * 2;
  constexpr int WCOLS = BN / 2;                           // columns one scanning warp takes of each code tile
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t a_smem = base;
  const uint32_t e_smem = a_smem + A_BYTES;
  const uint32_t misc = e_smem + static_cast<uint32_t>(p.stages) * STAGE_BYTES;
  float* margin_s = reinterpret_cast<float*>(gen + (misc - base));                    // [BM]
  int* cnt_s = reinterpret_cast<int*>(margin_s + BM);                                  // [BM][2]
  float* best_s = reinterpret_cast<float*>(cnt_s + BM * 2);                            // [BM][2]
  uint2* rec_s = reinterpret_cast<uint2*>(best_s + BM * 2);                            // [BM][2][RQ_RS]
  float* ee_slots = reinterpret_cast<float*>(rec_s + BM * 2 * RQ_RS);                  // [8][WCOLS]
  uint32_t* res_s = reinterpret_cast<uint32_t*>(ee_slots + RQ_NEPI * WCOLS);           // [RQ_MAXL][BM] global ids
  uint32_t* list_s = res_s + RQ_MAXL * BM;                                             // [14][RQ_LIST]
  int* order_s = reinterpret_cast<int*>(list_s + RQ_WORKERS * RQ_LIST);                // [BM] rows the one-thread prune did not certify
  int* ctl_s = order_s + BM;                                                           // [0] hard rows, [1] grab counter
  const uint32_t bar0 = (misc + rq_misc_bytes(BM, BN) + 7u) & ~7u;
  const uint32_t bar_full = bar0, bar_empty = bar0 + 8 * 8, bar_tfull = bar0 + 16 * 8, bar_tempty = bar0 + 18 * 8;
  const uint32_t bar_afull = bar0 + 20 * 8, tmem_slot = bar0 + 22 * 8;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(bar_tfull + 8 * b, 1); mbar_init(bar_tempty + 8 * b, RQ_NEPI); }
    mbar_init(bar_afull, RQ_WORKERS);
    for (int i = 0; i < 8; ++i) ctl_s[i] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(2u * BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));
  const int n_tiles = p.row_tiles;
  const int my_tiles = (n_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (warp == 0) {
    // ============================== TMA producer: codebook tiles of every level, in scan order ==============================
    uint32_t stage = 0, phase = 0;
    for (int it = 0; it < my_tiles; ++it)
      for (int l = 0; l < p.L; ++l)
        for (int t = 0; t < p.code_tiles; ++t)
          for (int kb = 0; kb < KBLK; ++kb) {
            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            if (elect_one()) {
              mbar_expect_tx(bar_full + 8 * stage, STAGE_BYTES);
              tma_load_2d(e_smem + stage * STAGE_BYTES, &tmap_e, bar_full + 8 * stage, kb * TC_KB, l * p.K_per + t * BN);
            }
            __syncwarp();
            if (++stage == static_cast<uint32_t>(p.stages)) { stage = 0; phase ^= 1; }
          }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    uint32_t stage = 0, phase = 0, tg = 0, aseq = 0;
    const uint32_t a_lo = umma_desc_lo(a_smem), e_lo = umma_desc_lo(e_smem);
    for (int it = 0; it < my_tiles; ++it)
      for (int l = 0; l < p.L; ++l, ++aseq) {
        mbar_wait(bar_afull, aseq & 1);                    // operand tile of (tile, level) is in shared memory
        tc_fence_after();
        RQ_TR(0, l, 0);
        for (int t = 0; t < p.code_tiles; ++t, ++tg) {
          const uint32_t b = tg & 1;
          mbar_wait(bar_tempty + 8 * b, (tg >> 1) & 1);     // drained AND pre-loaded with -|e|^2/2
          tc_fence_after();
          for (int kb = 0; kb < KBLK; ++kb) {
            mbar_wait(bar_full + 8 * stage, phase);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t a0 = a_lo + ((kb * (BM * 128)) >> 4);
              const uint32_t b0 = e_lo + ((stage * STAGE_BYTES) >> 4);
#pragma unroll
              for (int k = 0; k < TC_KB / 16; ++k)
                tc_mma_bf16(tmem_base + b * BN, umma_desc(a0 + k * 2), umma_desc(b0 + k * 2), p.idesc, 1u);
              tc_commit(bar_empty + 8 * stage);
              if (kb == KBLK - 1) tc_commit(bar_tfull + 8 * b);
            }
            __syncwarp();
            if (++stage == static_cast<uint32_t>(p.stages)) { stage = 0; phase ^= 1; }
          }
          if (t == 0) RQ_TR(0, l, 1);
        }
        RQ_TR(0, l, 2);
      }
  } else {
    // ============================== workers ==============================
    const int w = warp - 2;                                 // 0..13
    const bool scanner = warp >= 4 && warp < 12;            // warps 4..11: TMEM lane quarter = warp % 4
    const int quarter = warp & 3, cs = (warp - 4) >> 2;     // column slice of the scanning warp
    // accumulator row held by this lane: M = 128 -> lane quarter * 32 + lane; M = 64 -> rows 16 q .. 16 q + 15 sit in
    // lanes 0..15 of quarter q (cute::UMMA 1-SM M=64 accumulator atom), the upper half-quarter is unused
    const int scan_row = BM == 128 ? quarter * 32 + lane : (lane < 16 ? quarter * 16 + lane : -1);
    const float kNegInf = __int_as_float(0xff800000);
    float* ee_slot = ee_slots + (scanner ? (warp - 4) : 0) * WCOLS;
    const uint32_t tcol = (static_cast<uint32_t>(quarter * 32) << 16) + cs * WCOLS;
    uint32_t* my_list = list_s + w * RQ_LIST;
    float* my_scratch = p.scratch + static_cast<int64_t>(blockIdx.x) * BM * D;
    float err_acc = 0.f;

    // ---- bias pre-load machinery of the scanning warps (as in search_tc_kernel)
    constexpr int BPL = WCOLS / 32;
    struct Bias { float v[BPL]; };
    auto load_bias = [&](int l, int t) -> Bias {
      Bias r;
#pragma unroll
      for (int j = 0; j < BPL; ++j) {
        const int c = t * BN + cs * WCOLS + lane * BPL + j;
        r.v[j] = (t >= 0 && c < p.K_per) ? -p.ee_half[l * p.K_per + c] : kNegInf;
      }
      return r;
    };
    auto preload = [&](const Bias& bias, uint32_t b) {
      __syncwarp();
#pragma unroll
      for (int j = 0; j < BPL; ++j) ee_slot[lane * BPL + j] = bias.v[j];
      __syncwarp();
#pragma unroll
      for (int hh = 0; hh < WCOLS / 16; ++hh) {
        uint32_t wv[16];
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 q4 = *reinterpret_cast<const float4*>(ee_slot + hh * 16 + j);
          wv[j + 0] = __float_as_uint(q4.x); wv[j + 1] = __float_as_uint(q4.y);
          wv[j + 2] = __float_as_uint(q4.z); wv[j + 3] = __float_as_uint(q4.w);
        }
        TC_ST16(tmem_base + tcol + b * BN + hh * 16, wv);
      }
      tc_wait_st();
    };
    // sequence of (level, code tile) over this CTA's row tiles, walked two tiles ahead of the scan
    const int64_t total_seq = static_cast<int64_t>(my_tiles) * p.L * p.code_tiles;
    int64_t la = 0;
    int la_t = 0, la_l = 0;
    auto la_next = [&](int& l_out) -> int {
      if (la++ >= total_seq) { l_out = 0; return -1; }
      const int t = la_t;
      l_out = la_l;
      if (++la_t == p.code_tiles) { la_t = 0; if (++la_l == p.L) la_l = 0; }
      return t;
    };
    int t_ahead = -1, l_ahead = 0;
    Bias bias_next{};
    if (scanner) {
      for (uint32_t b = 0; b < 2; ++b) {
        int ll;
        const int tt = la_next(ll);
        if (tt >= 0) preload(load_bias(ll, tt), b);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty + 8 * b);
      }
      t_ahead = la_next(l_ahead);
      bias_next = load_bias(l_ahead, t_ahead);
    }

    uint8_t* a_tile = gen + (a_smem - base);
    uint32_t tg = 0, aseq = 0;
    // SCATTER: runs of rows (in this warp's order) that chose the same code are summed in registers and cost ONE set of
    // reductions (a collapsed level sends every row to one code: thousands of reductions on 128 addresses otherwise)
    float4 run[SCATTER ? SL : 1];
    int run_gid = -1, run_cnt = 0;
    auto run_flush = [&]() {
      if (!SCATTER || run_gid < 0) return;
#pragma unroll
      for (int s = 0; s < (SCATTER ? SL : 1); ++s)
        red_add_v4(p.seg_sum + static_cast<int64_t>(run_gid) * D + s * 128 + lane * 4, run[s]);
      if (lane == 0) {
        atomicAdd(p.seg_cnt + run_gid, static_cast<float>(run_cnt));
        if (p.hist) atomicAdd(p.hist + run_gid, run_cnt);
      }
      run_gid = -1;
    };
    auto run_add = [&](int gid, const float4 (&v)[SL]) {
      if (!SCATTER) return;
      if (gid != run_gid) {
        run_flush();
        run_gid = gid; run_cnt = 0;
#pragma unroll
        for (int s = 0; s < (SCATTER ? SL : 1); ++s) run[s] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      ++run_cnt;
#pragma unroll
      for (int s = 0; s < (SCATTER ? SL : 1); ++s) { run[s].x += v[s].x; run[s].y += v[s].y; run[s].z += v[s].z; run[s].w += v[s].w; }
    };
    for (int it = 0; it < my_tiles; ++it) {
      const int tile = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
      const int64_t row0 = static_cast<int64_t>(tile) * BM;

      // ---------------- level 0 operand tile from z (several rows of loads in flight per warp) ----------------
      {
        constexpr int RB0 = SL == 4 ? 2 : (SL == 3 ? 2 : 4);
        for (int r0 = w; r0 < BM; r0 += RQ_WORKERS * RB0) {
          float4 v[RB0][SL];
#pragma unroll
          for (int u = 0; u < RB0; ++u) {
            const int r = r0 + u * RQ_WORKERS;
            const int64_t grow = row0 + r;
#pragma unroll
            for (int s = 0; s < SL; ++s)
              v[u][s] = (r < BM && grow < p.n_rows) ? ld_stream(reinterpret_cast<const float4*>(p.z + grow * D) + s * 32 + lane)
                                                    : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int u = 0; u < RB0; ++u) {
            const int r = r0 + u * RQ_WORKERS;
            if (r < BM) rq_emit_operand_row<SL, BF16, BM>(a_tile, margin_s, r, v[u], p.level_meta, lane);
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tensor-core reads
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_afull);

      for (int l = 0; l < p.L; ++l, ++aseq) {
        // ---------------- scan: scores of the level out of TMEM into candidate records ----------------
        if (scanner) {
          mbar_wait(bar_afull, aseq & 1);                    // every row's margin is in shared memory
          if (warp == 4) RQ_TR(1, l, 0);
          const int r = scan_row;
          const float margin = r >= 0 ? margin_s[r] : __int_as_float(0x7fc00000);
          uint2* rec = rec_s + ((r >= 0 ? r : 0) * 2 + cs) * RQ_RS;
          float best = kNegInf;
          float thr = margin == margin ? -3.0e38f : margin;  // lowest finite value: -inf chunks are never admitted
          int cnt = 0;
          for (int t = 0; t < p.code_tiles; ++t, ++tg) {
            const uint32_t b = tg & 1;
            const Bias bias = bias_next;
            const int t_cur_ahead = t_ahead;
            t_ahead = la_next(l_ahead);
            bias_next = load_bias(l_ahead, t_ahead);
            mbar_wait(bar_tfull + 8 * b, (tg >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + tcol + b * BN;
            uint32_t v[32];
            const uint32_t code_t = static_cast<uint32_t>(t * BN + cs * WCOLS);
#pragma unroll
            for (int ch = 0; ch < WCOLS / 32; ++ch) {
              TC_LD32(taddr + ch * 32, v);
              tc_wait_ld();
              if (ch == WCOLS / 32 - 1) {
                if (t_cur_ahead >= 0) preload(bias, b);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tempty + 8 * b);
              }
              rq_scan(v, code_t + ch * 32, margin, best, thr, cnt, rec);
            }
          }
          if (r >= 0) {
            cnt_s[r * 2 + cs] = cnt;
            best_s[r * 2 + cs] = best;
          }
          if (warp == 4) RQ_TR(1, l, 1);
        }
        if (warp == 2) RQ_TR(2, l, 0);
        rq_bar_workers();                                    // records of every row are complete
        if (warp == 4) RQ_TR(1, l, 2);
        if (warp == 2) RQ_TR(2, l, 1);
        // ---------------- decide the rows ----------------
        // One thread per row (the slice-0 scanning lanes) prunes the row's records against the final threshold.  A single
        // surviving code is the certified arg max (the great majority of rows); everything else goes on the hard list.
        // (Two other shapes of this phase were built and measured slower at the stage-2 batch, 0.136-0.141 ms against
        // 0.121: a single dynamic work queue over all rows, hard ones first; and the resolving warp finishing its hard
        // row at once, with no barrier before the certified rows.  The separate tight loops below win.)
        if (scanner && cs == 0 && scan_row >= 0) {
          const int r = scan_row;
          const int64_t grow = row0 + r;
          if (grow < p.n_rows) {
            const int c0 = cnt_s[r * 2], c1 = cnt_s[r * 2 + 1];
            const float b0 = best_s[r * 2], b1 = best_s[r * 2 + 1];
            const float mg = margin_s[r];
            const bool bad = c0 > RQ_RS || c1 > RQ_RS || (c0 == 0 && b0 != kNegInf) || (c1 == 0 && b1 != kNegInf) || !(mg == mg);
            int nhit = 0;
            uint32_t f = 0;
            if (!bad) {
              const float thr = fmaxf(b0, b1) - mg;
              for (int i = 0; i < c0; ++i) {
                const uint2 e = rec_s[(r * 2) * RQ_RS + i];
                if (__uint_as_float(e.y) >= thr) { nhit += __popc(e.x & 0xffu); f = e.x; }
              }
              for (int i = 0; i < c1; ++i) {
                const uint2 e = rec_s[(r * 2 + 1) * RQ_RS + i];
                if (__uint_as_float(e.y) >= thr) { nhit += __popc(e.x & 0xffu); f = e.x; }
              }
            }
            if (nhit == 1) {
              const uint32_t gid = static_cast<uint32_t>(l * p.K_per) + ((f >> 8) << 3) + (__ffs(f & 0xffu) - 1);
              res_s[l * BM + r] = gid;
              p.idx_out[static_cast<int64_t>(l) * p.n_rows + grow] = gid;
            } else {
              order_s[atomicAdd(&ctl_s[0], 1)] = r;
            }
          } else {
            res_s[l * BM + r] = static_cast<uint32_t>(l * p.K_per);
          }
        }
        rq_bar_workers();                                    // the hard list is complete
        if (warp == 4) RQ_TR(1, l, 3);
        const int n_hard = ctl_s[0];
        const float* res_src = l == 0 ? p.z + row0 * D : my_scratch;        // fp32 residual rows of this tile
        // Hard rows, taken dynamically by all fourteen warps: lanes 0 .. 2 RQ_RS - 1 hold the records of the two slices.
        for (;;) {
          int hi = 0;
          if (lane == 0) hi = atomicAdd(&ctl_s[1], 1);
          hi = __shfl_sync(0xffffffffu, hi, 0);
          if (hi >= n_hard) break;
          const int r = order_s[hi];
          const int64_t grow = row0 + r;
          const float* res_row = res_src + r * D;            // fp32 residual row
          float4 v[SL];
#pragma unroll
          for (int s = 0; s < SL; ++s) v[s] = reinterpret_cast<const float4*>(res_row)[s * 32 + lane];
          const int c0 = cnt_s[r * 2], c1 = cnt_s[r * 2 + 1];
          const float b0 = best_s[r * 2], b1 = best_s[r * 2 + 1];
          const float mg = margin_s[r];
          const bool bad = c0 > RQ_RS || c1 > RQ_RS || (c0 == 0 && b0 != kNegInf) || (c1 == 0 && b1 != kNegInf) || !(mg == mg);
          const float thr = fmaxf(b0, b1) - mg;
          const int sl = lane / RQ_RS, slot = lane % RQ_RS;
          uint2 ent = make_uint2(0u, 0xff800000u);
          const bool have = lane < 2 * RQ_RS && slot < (sl == 0 ? c0 : c1) && !bad;
          if (have) ent = rec_s[(r * 2 + sl) * RQ_RS + slot];
          const bool hit = have && __uint_as_float(ent.y) >= thr;
          const uint32_t mk = hit ? (ent.x & 0xffu) : 0u;
          int pre = __popc(mk);                              // inclusive prefix sum of the per-lane code counts
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, pre, o);
            if (lane >= o) pre += u;
          }
          const int ncodes = __shfl_sync(0xffffffffu, pre, 31);
          uint32_t pick;
          if (bad || ncodes < 1 || ncodes > RQ_LIST) {       // exhaustive exact search
            pick = rq_exhaustive<SL, BF16>(res_row, p.E, p.E_lp, p.ee_half, p.K_per, l, lane);
            if (lane == 0 && p.counters) atomicAdd(p.counters, 1);
          } else {                                            // exact re-rank of the survivors
            int at = pre - __popc(mk);
            uint32_t m = mk;
            while (m) {
              my_list[at++] = ((ent.x >> 8) << 3) + (__ffs(m) - 1);
              m &= m - 1;
            }
            __syncwarp();
            double top = -1e300;
            uint32_t top_idx = 0xffffffffu;
            constexpr int CB = SL <= 2 ? 4 : 2;               // candidates whose code rows are in flight together
            for (int c0i = 0; c0i < ncodes; c0i += CB) {
              float4 ev[CB][SL];
              uint32_t code[CB];
#pragma unroll
              for (int u = 0; u < CB; ++u) {
                code[u] = my_list[c0i + u < ncodes ? c0i + u : ncodes - 1];
                rq_load_code_row<SL, BF16>(ev[u], p.E, p.E_lp, static_cast<int64_t>(l) * p.K_per + code[u], lane);
              }
#pragma unroll
              for (int u = 0; u < CB; ++u) {
                const double sc = rq_exact_score<SL, BF16>(v, ev[u]);
                if (c0i + u < ncodes && (sc > top || (sc == top && code[u] < top_idx))) { top = sc; top_idx = code[u]; }
              }
            }
            pick = top_idx;
            __syncwarp();
          }
          if (lane == 0) {
            const uint32_t gid = static_cast<uint32_t>(l * p.K_per) + pick;
            res_s[l * BM + r] = gid;
            p.idx_out[static_cast<int64_t>(l) * p.n_rows + grow] = gid;
          }
        }
        rq_bar_workers();                                    // every row of the level has its code
        if (warp == 2) {
          RQ_TR(2, l, 2);
          if (lane == 0) {
            if (p.counters && n_hard) atomicAdd(p.counters + 1, n_hard);
            ctl_s[0] = 0; ctl_s[1] = 0;                      // for the next level (ordered by its first worker barrier)
          }
        }

        const float* meta_next = p.level_meta + (l + 1 < p.L ? l + 1 : l) * VQB200_LEVEL_META_FLOATS;
        if (l + 1 < p.L) {
          // ---------------- next residual fl(r - e) (models/vq_vae.py:258) ----------------
          // A row (two at D <= 256) at a time per warp, the residual row and its code row in flight together; the new
          // residual goes to the CTA's scratch tile (L2), to the operand tile and into the next level's margin.
          constexpr int RB = SL <= 2 ? 2 : 1;               // the 128-register budget holds 32 float4 of row data without spilling
          for (int r0 = w; r0 < BM; r0 += RQ_WORKERS * RB) {
            float4 v[RB][SL], e4[RB][SL];
            int rr[RB];
#pragma unroll
            for (int u = 0; u < RB; ++u) {
              rr[u] = r0 + u * RQ_WORKERS;
              if (rr[u] < BM) {
                const bool valid = row0 + rr[u] < p.n_rows;
                const uint32_t gid = res_s[l * BM + rr[u]];
#pragma unroll
                for (int s = 0; s < SL; ++s) {
                  v[u][s] = valid ? reinterpret_cast<const float4*>(res_src + rr[u] * D)[s * 32 + lane]
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
                  e4[u][s] = __ldg(reinterpret_cast<const float4*>(p.E + static_cast<int64_t>(gid) * D) + s * 32 + lane);
                }
              }
            }
#pragma unroll
            for (int u = 0; u < RB; ++u) {
              if (rr[u] >= BM) continue;
              const bool valid = row0 + rr[u] < p.n_rows;
              if (valid) run_add(static_cast<int>(res_s[l * BM + rr[u]]), v[u]);
#pragma unroll
              for (int s = 0; s < SL; ++s) {
                if (valid) {
                  v[u][s].x = __fsub_rn(v[u][s].x, e4[u][s].x); v[u][s].y = __fsub_rn(v[u][s].y, e4[u][s].y);
                  v[u][s].z = __fsub_rn(v[u][s].z, e4[u][s].z); v[u][s].w = __fsub_rn(v[u][s].w, e4[u][s].w);
                }
                reinterpret_cast<float4*>(my_scratch + rr[u] * D)[s * 32 + lane] = v[u][s];
              }
              rq_emit_operand_row<SL, BF16, BM>(a_tile, margin_s, rr[u], v[u], meta_next, lane);
            }
          }
        } else {
          // ---------------- outputs: z_q = ((E[i0] + E[i1]) + ...) in level order (:261), z_q_st (:263) ----------------
          // Per row, SS float4 slices at a time: the z slice and the slice of EVERY level's code row are loaded
          // together (one L2 round trip per slice group, not one per level).
          constexpr int SS = SL % 2 == 0 ? 2 : 1;
          for (int r = w; r < BM; r += RQ_WORKERS) {
            const int64_t grow = row0 + r;
            if (grow >= p.n_rows) continue;
            if (!SCATTER && p.hist && lane < p.L) atomicAdd(p.hist + res_s[lane * BM + r], 1);
            if (SCATTER) {
              float4 v[SL];
#pragma unroll
              for (int s = 0; s < SL; ++s) v[s] = reinterpret_cast<const float4*>(res_src + r * D)[s * 32 + lane];
              run_add(static_cast<int>(res_s[l * BM + r]), v);
            }
#pragma unroll
            for (int s0 = 0; s0 < SL; s0 += SS) {
              float4 zz[SS], q[SS];
#pragma unroll
              for (int u = 0; u < SS; ++u)
                zz[u] = ld_stream(reinterpret_cast<const float4*>(p.z + grow * D) + (s0 + u) * 32 + lane);
              constexpr int LG = 4;                            // levels whose code-row slices are in flight together
              for (int l0 = 0; l0 < p.L; l0 += LG) {
                float4 c[LG][SS];
#pragma unroll
                for (int j = 0; j < LG; ++j) {
                  if (l0 + j < p.L) {
                    const uint32_t g = res_s[(l0 + j) * BM + r];
#pragma unroll
                    for (int u = 0; u < SS; ++u)
                      c[j][u] = __ldg(reinterpret_cast<const float4*>(p.E + static_cast<int64_t>(g) * D) + (s0 + u) * 32 + lane);
                  }
                }
#pragma unroll
                for (int j = 0; j < LG; ++j) {
                  if (l0 + j < p.L) {
#pragma unroll
                    for (int u = 0; u < SS; ++u) {
                      if (l0 + j == 0) q[u] = c[j][u];
                      else {
                        q[u].x = __fadd_rn(q[u].x, c[j][u].x); q[u].y = __fadd_rn(q[u].y, c[j][u].y);
                        q[u].z = __fadd_rn(q[u].z, c[j][u].z); q[u].w = __fadd_rn(q[u].w, c[j][u].w);
                      }
                    }
                  }
                }
              }
#pragma unroll
              for (int u = 0; u < SS; ++u) {
                float4 df;
                df.x = __fsub_rn(q[u].x, zz[u].x); df.y = __fsub_rn(q[u].y, zz[u].y);
                df.z = __fsub_rn(q[u].z, zz[u].z); df.w = __fsub_rn(q[u].w, zz[u].w);
                if (p.zq_out) st_stream(reinterpret_cast<float4*>(p.zq_out + grow * D) + (s0 + u) * 32 + lane, q[u]);
                if (p.zq_st_out)
                  st_stream(reinterpret_cast<float4*>(p.zq_st_out + grow * D) + (s0 + u) * 32 + lane,
                            make_float4(__fadd_rn(zz[u].x, df.x), __fadd_rn(zz[u].y, df.y), __fadd_rn(zz[u].z, df.z),
                                        __fadd_rn(zz[u].w, df.w)));
                err_acc = fmaf(df.x, df.x, err_acc); err_acc = fmaf(df.y, df.y, err_acc);
                err_acc = fmaf(df.z, df.z, err_acc); err_acc = fmaf(df.w, df.w, err_acc);
              }
            }
          }
        }
        run_flush();
        if (warp == 4) RQ_TR(1, l, 4);
        if (warp == 2) RQ_TR(2, l, 3);
        if (l + 1 < p.L) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_afull);
        } else {
          rq_bar_workers();                                  // res_s / records are reused by the next tile
        }
      }
    }
    if (p.sqerr_sum) {
      const double e = warp_sum(static_cast<double>(err_acc));
      if (lane == 0 && e != 0.0) atomicAdd(p.sqerr_sum, e);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2u * BN) : "memory");
  }
  if (p.stats_out) {
    // every CTA's histogram / squared-error reductions precede its ticket; the last one sees them all
    __shared__ bool s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(p.counters + 2, 1) == static_cast<int>(gridDim.x) - 1;
    __syncthreads();
    if (s_last) {
      __threadfence();
      stats_finalize_block(p.hist, p.K_per * p.L, p.count_add, p.sqerr_sum, p.inv_elems, p.ep_usage, p.ep_cnt, p.stats_out);
    }
  }
}

// ------------------------------------------------------------------------------------ launch (one translation unit per D)
// dynamic shared memory the kernel may ask for: the 227 KB of the SM minus its static shared memory (statistics tail)
constexpr int RQ_SMEM_LIMIT = TC_SMEM_LIMIT - 1024;
struct RqConfig { int BM, BN, stages, grid, smem; };

template <int SL, int BM, int BN, bool SCATTER>
static inline int launch_rq(const CUtensorMap& map_e, const RvqParams& p, bool bf, int grid, int smem, cudaStream_t s) {
  static bool attr_done_dev[64] = {};
  bool& attr_done = attr_done_dev[current_device_slot()];
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(rvq_fused_kernel<SL, false, BM, BN, SCATTER>, cudaFuncAttributeMaxDynamicSharedMemorySize, RQ_SMEM_LIMIT);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(rvq_fused_kernel<SL, true, BM, BN, SCATTER>, cudaFuncAttributeMaxDynamicSharedMemorySize, RQ_SMEM_LIMIT);
    if (e != cudaSuccess) return status_of(e);
    attr_done = true;
  }
  timing_mark_begin(s);
  if (bf) rvq_fused_kernel<SL, true, BM, BN, SCATTER><<<grid, RQ_THREADS, smem, s>>>(map_e, p);
  else rvq_fused_kernel<SL, false, BM, BN, SCATTER><<<grid, RQ_THREADS, smem, s>>>(map_e, p);
  timing_mark_end(s);
  return status_of(cudaGetLastError());
}

template <int SL, bool SCATTER>
static inline int launch_rq_shape(const CUtensorMap& map_e, const RvqParams& p, bool bf, const RqConfig& c, cudaStream_t s) {
  if (c.BM == 64) return launch_rq<SL, 64, 256, SCATTER>(map_e, p, bf, c.grid, c.smem, s);
  if constexpr (SL == 4) return launch_rq<SL, 128, 128, SCATTER>(map_e, p, bf, c.grid, c.smem, s);
  else return launch_rq<SL, 128, 256, SCATTER>(map_e, p, bf, c.grid, c.smem, s);
}


// The kernels of one D = 128 SL are instantiated in their own translation unit (vq_rvq_fused_sl<SL>.cu: the four build in
// parallel); vq_rvq_fused.cu calls them through these plain functions.
int rq_launch_sl1(const CUtensorMap& map_e, const RvqParams& p, bool bf, const RqConfig& c, bool scatter, cudaStream_t s);
int rq_launch_sl2(const CUtensorMap& map_e, const RvqParams& p, bool bf, const RqConfig& c, bool scatter, cudaStream_t s);
int rq_launch_sl3(const CUtensorMap& map_e, const RvqParams& p, bool bf, const RqConfig& c, bool scatter, cudaStream_t s);
int rq_launch_sl4(const CUtensorMap& map_e, const RvqParams& p, bool bf, const RqConfig& c, bool scatter, cudaStream_t s);
#define RQ_DEFINE_LAUNCH_SL(SLV)                                                                                          \
  int rq_launch_sl##SLV(const CUtensorMap& map_e, const RvqParams& p, bool bf, const RqConfig& c, bool scatter,           \
                        cudaStream_t s) {                                                                                 \
    return scatter ? launch_rq_shape<SLV, true>(map_e, p, bf, c, s) : launch_rq_shape<SLV, false>(map_e, p, bf, c, s);   \
  }

}  // namespace vqb
