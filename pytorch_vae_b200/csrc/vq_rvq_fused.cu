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
  // ---- training mode (TRAIN): the local EMA update of models/vq_vae.py:251 / :77-89 after EVERY level, inside the kernel
  float* E_rw;                  // = E
  uint16_t* plane_bf16;         // [K_total, D] bf16 operand plane
  uint16_t* plane_f16;          // [K_total, D] fp16 operand plane
  float* ee_rw;                 // [2][K_total]  |e|^2/2 of the fp32 codes / of the bf16-rounded codes
  float* level_meta_rw;         // [L][8] the cache's copy (left in its canonical state at the end)
  float* meta_buf;              // [2][RQ_MAXL][8] level norms of the refreshed codebook, double-buffered by level parity
  float* ema_cs;                // [K_total]
  float* ema_emb;               // [K_total, D]
  float* seg_sum;               // [K_total, D] zero on entry
  float* seg_cnt;               // [K_total]    zero on entry
  unsigned* grid_bar;           // zero on entry
  float decay, omd, eps;
  int K_total;
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
__device__ __forceinline__ float rq_margin(bool bf16, int D, float ss, float sse, const float* meta) {
  float m;
  if (bf16) m = 2.f * static_cast<float>(D + 32) * 1.1920929e-7f * (sqrtf(ss) * 1.0001f) * meta[2] + 1e-30f;
  else m = admission_margin_fp32(ss, sse, meta[0], meta[4], meta[5], D);
  if (meta[1] != 0.f || !(ss < __int_as_float(0x7f800000)) || !(sse < __int_as_float(0x7f800000)))
    m = __int_as_float(0x7fc00000);   // NaN: exhaustive search
  return m;
}
// DEFER: the level norms are not known yet (training: the codebook is refreshed first) -- leave |r|^2 and |r - r~|^2 in
// norms_out[r][2]; the margins are formed once the refreshed norms exist.
template <int SL, bool BF16, int BM, bool DEFER = false>
__device__ __forceinline__ void rq_emit_operand_row(uint8_t* a_tile, float* margin_s, int r, const float4 (&v)[SL],
                                                    const float* meta, int lane, float* norms_out = nullptr) {
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
    if (DEFER) { norms_out[r * 2] = ss; norms_out[r * 2 + 1] = sse; }
    else margin_s[r] = rq_margin(BF16, D, ss, sse, meta);
  }
}
// one code row of the exact re-rank (fp32 codes; bf16_input mode: the rounded plane)
template <bool CG>
__device__ __forceinline__ float4 rq_ld4(const float4* p) { return CG ? __ldcg(p) : __ldg(p); }
template <int SL, bool BF16, bool CG = false>
__device__ __forceinline__ void rq_load_code_row(float4 (&ev)[SL], const float* E, const uint16_t* E_lp, int64_t gid, int lane) {
  constexpr int D = SL * 128;
#pragma unroll
  for (int s = 0; s < SL; ++s) {
    if (BF16) {
      const uint2 b = CG ? __ldcg(reinterpret_cast<const uint2*>(E_lp + gid * D + s * 128 + lane * 4))
                         : *reinterpret_cast<const uint2*>(E_lp + gid * D + s * 128 + lane * 4);
      ev[s] = make_float4(__uint_as_float(b.x << 16), __uint_as_float(b.x & 0xffff0000u),
                          __uint_as_float(b.y << 16), __uint_as_float(b.y & 0xffff0000u));
    } else {
      ev[s] = rq_ld4<CG>(reinterpret_cast<const float4*>(E + gid * D + s * 128 + lane * 4));
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
  dot = warp_sum(dot);
  ee = warp_sum(ee);
  return dot - 0.5 * ee;
}
// exhaustive exact search of one level by the warp: d' = |e|^2/2 - r.e (fp32), packed (key, index) minimum -- the rule
// of search_simt_kernel (lowest index on ties, a NaN distance wins).  Rare: kept out of line; it re-loads the row.
// first_zero (training): K_per - local index of the level's first all-zero code (0 = none); the other all-zero codes are
// skipped, as the refreshed cache marks them with |e|^2/2 = +inf (see codebook_refresh_kernel).
template <int SL, bool BF16, bool CG = false>
__device__ __noinline__ uint32_t rq_exhaustive(const float* row, const float* E, const uint16_t* E_lp, const float* ee_half,
                                               int K_per, int l, int lane, int first_zero = 0) {
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
      eeh[u] = CG ? __ldcg(ee_half + gid) : ee_half[gid];
      if (CG && eeh[u] == 0.f && first_zero > 0 && K_per - k != first_zero) eeh[u] = __int_as_float(0x7f800000);
      rq_load_code_row<SL, BF16, CG>(e4[u], E, E_lp, gid, lane);
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

// ---- training mode helpers
__device__ __forceinline__ unsigned rq_ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Barrier across the (co-resident: cooperative launch, one CTA per SM) grid, entered by the fourteen worker warps of
// every CTA.  Everything written before it is visible to every thread after it (gpu-scope fences on both sides; the
// second one also drops this SM's L1 lines).
__device__ __forceinline__ void rq_grid_barrier(unsigned* ctr, unsigned target) {
  __threadfence();
  rq_bar_workers();
  if (threadIdx.x == 64) {
    atomicAdd(ctr, 1u);
    unsigned spins = 0;
    while (rq_ld_acquire(ctr) < target)
      if (++spins > (1u << 22)) __trap();                 // a protocol bug must fault, never hang the GPU
    __threadfence();
  }
  rq_bar_workers();
}

// EMA update (models/vq_vae.py:85-89, for ALL codes of ALL levels) and cache refresh of the code rows gw, gw + GW, ... by
// one warp each: the arithmetic of codebook_refresh_kernel<1> (vq_rowops.cu), reading the segment sums the row passes
// of this level reduced into seg_sum / seg_cnt and zeroing them for the next level.  Level norms are max-reduced into
// meta_acc (shared memory, [RQ_MAXL][8] ints).
template <int SL>
__device__ __forceinline__ void rq_refresh_codes(const RvqParams& p, int gw, int GW, int lane, int* meta_acc) {
  constexpr int D4 = SL * 32;
  float4* EM = reinterpret_cast<float4*>(p.ema_emb);
  float4* SG = reinterpret_cast<float4*>(p.seg_sum);
  float4* E4 = reinterpret_cast<float4*>(p.E_rw);
  uint2* PB = reinterpret_cast<uint2*>(p.plane_bf16);
  uint2* PH = reinterpret_cast<uint2*>(p.plane_f16);
  for (int row = gw; row < p.K_total; row += GW) {
    // :85 cs.mul_(decay).add_(n * (1 - decay)) -- two roundings, no fma contraction
    const float cs = __fadd_rn(__fmul_rn(__ldcg(p.ema_cs + row), p.decay), __fmul_rn(__ldcg(p.seg_cnt + row), p.omd));
    __syncwarp();
    if (lane == 0) { p.ema_cs[row] = cs; p.seg_cnt[row] = 0.f; }
    const float denom = __fadd_rn(cs, p.eps);
    double acc = 0.0, accb = 0.0, accd = 0.0, acch = 0.0, accdh = 0.0;
    bool bad = false;
#pragma unroll
    for (int sl = 0; sl < SL; ++sl) {
      const int64_t o = static_cast<int64_t>(row) * D4 + sl * 32 + lane;
      const float4 m = __ldcg(EM + o), sm = __ldcg(SG + o);
      SG[o] = make_float4(0.f, 0.f, 0.f, 0.f);
      float4 n, e;
      n.x = __fadd_rn(__fmul_rn(m.x, p.decay), __fmul_rn(sm.x, p.omd));
      n.y = __fadd_rn(__fmul_rn(m.y, p.decay), __fmul_rn(sm.y, p.omd));
      n.z = __fadd_rn(__fmul_rn(m.z, p.decay), __fmul_rn(sm.z, p.omd));
      n.w = __fadd_rn(__fmul_rn(m.w, p.decay), __fmul_rn(sm.w, p.omd));
      EM[o] = n;
      e.x = __fdiv_rn(n.x, denom); e.y = __fdiv_rn(n.y, denom);       // :88 E = es / (cs + eps)
      e.z = __fdiv_rn(n.z, denom); e.w = __fdiv_rn(n.w, denom);
      E4[o] = e;
      const __nv_bfloat16 b0 = __float2bfloat16_rn(e.x), b1 = __float2bfloat16_rn(e.y),
                          b2 = __float2bfloat16_rn(e.z), b3 = __float2bfloat16_rn(e.w);
      uint2 pk;
      pk.x = static_cast<uint32_t>(__bfloat16_as_ushort(b0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(b1)) << 16);
      pk.y = static_cast<uint32_t>(__bfloat16_as_ushort(b2)) | (static_cast<uint32_t>(__bfloat16_as_ushort(b3)) << 16);
      PB[o] = pk;
      const float f0 = __bfloat162float(b0), f1 = __bfloat162float(b1), f2 = __bfloat162float(b2), f3 = __bfloat162float(b3);
      acc += static_cast<double>(e.x) * e.x + static_cast<double>(e.y) * e.y + static_cast<double>(e.z) * e.z +
             static_cast<double>(e.w) * e.w;
      accb += static_cast<double>(f0) * f0 + static_cast<double>(f1) * f1 + static_cast<double>(f2) * f2 +
              static_cast<double>(f3) * f3;
      {
        const double d0 = e.x - f0, d1 = e.y - f1, d2 = e.z - f2, d3 = e.w - f3;
        accd += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
      }
      {
        const uint16_t h0 = f16_bits_flush(e.x), h1 = f16_bits_flush(e.y), h2 = f16_bits_flush(e.z), h3 = f16_bits_flush(e.w);
        uint2 ph;
        ph.x = static_cast<uint32_t>(h0) | (static_cast<uint32_t>(h1) << 16);
        ph.y = static_cast<uint32_t>(h2) | (static_cast<uint32_t>(h3) << 16);
        PH[o] = ph;
        const double g0 = f16_bits_to_float(h0), g1 = f16_bits_to_float(h1), g2 = f16_bits_to_float(h2), g3 = f16_bits_to_float(h3);
        acch += g0 * g0 + g1 * g1 + g2 * g2 + g3 * g3;
        const double d0 = e.x - g0, d1 = e.y - g1, d2 = e.z - g2, d3 = e.w - g3;
        accdh += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
      }
      bad |= !(isfinite(e.x) && isfinite(e.y) && isfinite(e.z) && isfinite(e.w));
    }
    acc = warp_sum(acc); accb = warp_sum(accb); accd = warp_sum(accd); acch = warp_sum(acch); accdh = warp_sum(accdh);
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) {
      p.ee_rw[row] = static_cast<float>(0.5 * acc);
      p.ee_rw[p.K_total + row] = static_cast<float>(0.5 * accb);
      const float kInfF = __int_as_float(0x7f800000);
      const float n0 = static_cast<float>(sqrt(acc)) * 1.0000002f, n1 = static_cast<float>(sqrt(accb)) * 1.0000002f;
      const float n3 = static_cast<float>(sqrt(accd)) * 1.0000002f;
      const float n4 = static_cast<float>(sqrt(acch)) * 1.0000002f, n5 = static_cast<float>(sqrt(accdh)) * 1.0000002f;
      int* ma = meta_acc + (row / p.K_per) * VQB200_LEVEL_META_FLOATS;
      if (n0 == n0 && n0 < kInfF) atomicMax(ma + 0, __float_as_int(n0));
      if (n1 == n1 && n1 < kInfF) atomicMax(ma + 2, __float_as_int(n1));
      if (n3 == n3 && n3 < kInfF) atomicMax(ma + 3, __float_as_int(n3));
      if (n4 == n4) atomicMax(ma + 4, __float_as_int(n4));
      if (n5 == n5) atomicMax(ma + 5, __float_as_int(n5));
      if (acc == 0.0) atomicMax(ma + 7, p.K_per - (row % p.K_per));     // the LOWEST all-zero code of the level
      if (bad || !(acc == acc) || isinf(static_cast<float>(acc))) ma[1] = __float_as_int(1.0f);
    }
  }
}

template <int SL, bool BF16, int BM, int BN, bool TRAIN>
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
  const uint32_t bar_afull = bar0 + 20 * 8, bar_go = bar0 + 21 * 8, tmem_slot = bar0 + 22 * 8;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(bar_tfull + 8 * b, 1); mbar_init(bar_tempty + 8 * b, RQ_NEPI); }
    mbar_init(bar_afull, RQ_WORKERS);
    mbar_init(bar_go, 1);
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
      for (int l = 0; l < p.L; ++l) {
        if (TRAIN) {                                         // the level's operand plane is final (refreshed by the whole grid)
          mbar_wait(bar_go, l & 1);
          asm volatile("fence.proxy.async.global;" ::: "memory");
        }
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

    // level norms of the codebook level l is searched against: the cache's (eval; training: level 0), or the buffer the
    // refresh after level l - 1 filled
    auto meta_of = [&](int l) -> const float* {
      return (TRAIN && l > 0 ? p.meta_buf + (l & 1) * (RQ_MAXL * VQB200_LEVEL_META_FLOATS) : p.level_meta) +
             l * VQB200_LEVEL_META_FLOATS;
    };
    // ---- bias pre-load machinery of the scanning warps (as in search_tc_kernel)
    constexpr int BPL = WCOLS / 32;
    struct Bias { float v[BPL]; };
    auto load_bias = [&](int l, int t) -> Bias {
      Bias r;
      int first_zero = 0;
      if (TRAIN) first_zero = __float_as_int(__ldcg(meta_of(l) + 7));
#pragma unroll
      for (int j = 0; j < BPL; ++j) {
        const int c = t * BN + cs * WCOLS + lane * BPL + j;
        if (TRAIN) {
          float e = (t >= 0 && c < p.K_per) ? __ldcg(p.ee_half + l * p.K_per + c) : -kNegInf;
          if (e == 0.f && first_zero > 0 && p.K_per - c != first_zero) e = -kNegInf;   // duplicate all-zero code: skipped
          r.v[j] = -e;
        } else {
          r.v[j] = (t >= 0 && c < p.K_per) ? -p.ee_half[l * p.K_per + c] : kNegInf;
        }
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
    if (scanner && !TRAIN) {
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
    unsigned gbar = 0;                                      // grid barriers passed (training)
    if (TRAIN && warp == 2 && lane == 0) mbar_arrive(bar_go);   // level 0 reads the codebook as it was handed in
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
          if (TRAIN) {
            // the accumulator buffers of the level's first two tiles are armed here, once the refreshed |e|^2/2 exist
            // (eval arms them two tiles ahead, across the level boundary)
            for (int t = 0; t < 2 && t < p.code_tiles; ++t) {
              preload(load_bias(l, t), (tg + t) & 1);
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar_tempty + 8 * ((tg + t) & 1));
            }
          }
          for (int t = 0; t < p.code_tiles; ++t, ++tg) {
            const uint32_t b = tg & 1;
            Bias bias;
            int t_cur_ahead;
            if (TRAIN) {
              t_cur_ahead = t + 2 < p.code_tiles ? t + 2 : -1;
              bias = load_bias(l, t_cur_ahead);
            } else {
              bias = bias_next;
              t_cur_ahead = t_ahead;
              t_ahead = la_next(l_ahead);
              bias_next = load_bias(l_ahead, t_ahead);
            }
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
                if (!TRAIN || t_cur_ahead >= 0) {             // training: the next level arms its own first buffers
                  tc_fence_before();
                  __syncwarp();
                  if (lane == 0) mbar_arrive(bar_tempty + 8 * b);
                }
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
        // (A single dynamic work queue over all rows -- hard ones first, residual update right after the decision -- was
        // measured and is no faster: the phase is bound by instruction issue and L2 latency, not by imbalance.)
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
            pick = rq_exhaustive<SL, BF16, TRAIN>(res_row, p.E, p.E_lp, p.ee_half, p.K_per, l, lane,
                                                  TRAIN ? __float_as_int(__ldcg(meta_of(l) + 7)) : 0);
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
                rq_load_code_row<SL, BF16, TRAIN>(ev[u], p.E, p.E_lp, static_cast<int64_t>(l) * p.K_per + code[u], lane);
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
        if (TRAIN) {
          // ---------------- training: residual update + z_q accumulation + EMA segment sums, then the refresh ----------------
          // z_q is accumulated level by level from the codes as they are BEFORE this level's update (:248 -> :251); the
          // residual rows fl(r) enter the segment sums of their codes (:80-83), runs of one code summed in registers
          // first (a collapsed level sends every row to one code: thousands of reductions on 128 addresses otherwise).
          const bool last = l + 1 == p.L;
          float4 run[SL];
          int run_gid = -1, run_cnt = 0;
          auto flush = [&]() {
            if (run_gid < 0) return;
#pragma unroll
            for (int s = 0; s < SL; ++s)
              red_add_v4(p.seg_sum + static_cast<int64_t>(run_gid) * D + s * 128 + lane * 4, run[s]);
            if (lane == 0) {
              atomicAdd(p.seg_cnt + run_gid, static_cast<float>(run_cnt));
              if (p.hist) atomicAdd(p.hist + run_gid, run_cnt);
            }
          };
          for (int r = w; r < BM; r += RQ_WORKERS) {
            const int64_t grow = row0 + r;
            if (grow >= p.n_rows) {
              if (lane == 0) { best_s[r * 2] = 0.f; best_s[r * 2 + 1] = 0.f; }
              continue;
            }
            const int gid = static_cast<int>(res_s[l * BM + r]);
            float4 v[SL], e4[SL];
#pragma unroll
            for (int s = 0; s < SL; ++s) {
              v[s] = reinterpret_cast<const float4*>(res_src + r * D)[s * 32 + lane];
              e4[s] = __ldcg(reinterpret_cast<const float4*>(p.E + static_cast<int64_t>(gid) * D) + s * 32 + lane);
            }
            if (gid != run_gid) {
              flush();
              run_gid = gid; run_cnt = 0;
#pragma unroll
              for (int s = 0; s < SL; ++s) run[s] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            ++run_cnt;
#pragma unroll
            for (int s = 0; s < SL; ++s) { run[s].x += v[s].x; run[s].y += v[s].y; run[s].z += v[s].z; run[s].w += v[s].w; }
#pragma unroll
            for (int s = 0; s < SL; ++s) {
              float4* zq_p = reinterpret_cast<float4*>(p.zq_out + grow * D) + s * 32 + lane;
              float4 q = e4[s];
              if (l > 0) {
                const float4 a = __ldcg(zq_p);
                q.x = __fadd_rn(a.x, q.x); q.y = __fadd_rn(a.y, q.y); q.z = __fadd_rn(a.z, q.z); q.w = __fadd_rn(a.w, q.w);
              }
              *zq_p = q;
              if (!last) {
                v[s].x = __fsub_rn(v[s].x, e4[s].x); v[s].y = __fsub_rn(v[s].y, e4[s].y);
                v[s].z = __fsub_rn(v[s].z, e4[s].z); v[s].w = __fsub_rn(v[s].w, e4[s].w);
                reinterpret_cast<float4*>(my_scratch + r * D)[s * 32 + lane] = v[s];
              } else {
                const float4 zz = ld_stream(reinterpret_cast<const float4*>(p.z + grow * D) + s * 32 + lane);
                float4 df;
                df.x = __fsub_rn(q.x, zz.x); df.y = __fsub_rn(q.y, zz.y); df.z = __fsub_rn(q.z, zz.z); df.w = __fsub_rn(q.w, zz.w);
                if (p.zq_st_out)
                  st_stream(reinterpret_cast<float4*>(p.zq_st_out + grow * D) + s * 32 + lane,
                            make_float4(__fadd_rn(zz.x, df.x), __fadd_rn(zz.y, df.y), __fadd_rn(zz.z, df.z), __fadd_rn(zz.w, df.w)));
                err_acc = fmaf(df.x, df.x, err_acc); err_acc = fmaf(df.y, df.y, err_acc);
                err_acc = fmaf(df.z, df.z, err_acc); err_acc = fmaf(df.w, df.w, err_acc);
              }
            }
            if (!last) rq_emit_operand_row<SL, BF16, BM, true>(a_tile, margin_s, r, v, nullptr, lane, best_s);
          }
          flush();
          // ---- every CTA has reduced its rows and is done reading the codebook: refresh ALL codes (:85-89)
          int* meta_acc = cnt_s;                             // [RQ_MAXL][8]: the records are dead until the next scan
          rq_grid_barrier(p.grid_bar, ++gbar * gridDim.x);
          const int wt = static_cast<int>(threadIdx.x) - 64;
          if (wt < RQ_MAXL * VQB200_LEVEL_META_FLOATS) {
            meta_acc[wt] = 0;
            // this level's norms have been read by everyone: their buffer is free for the refresh after the NEXT level
            if (blockIdx.x == 0) p.meta_buf[(l & 1) * (RQ_MAXL * VQB200_LEVEL_META_FLOATS) + wt] = 0.f;
          }
          rq_bar_workers();
          rq_refresh_codes<SL>(p, static_cast<int>(blockIdx.x) * RQ_WORKERS + w, static_cast<int>(gridDim.x) * RQ_WORKERS, lane, meta_acc);
          rq_bar_workers();
          float* meta_w = p.meta_buf + ((l + 1) & 1) * (RQ_MAXL * VQB200_LEVEL_META_FLOATS);   // norms of the refreshed codebook
          if (wt < p.L * VQB200_LEVEL_META_FLOATS && meta_acc[wt] != 0) {
            if ((wt & 7) == 1) meta_w[wt] = 1.0f;
            else atomicMax(reinterpret_cast<int*>(meta_w) + wt, meta_acc[wt]);
          }
          rq_grid_barrier(p.grid_bar, ++gbar * gridDim.x);
          // duplicate all-zero codes of a level leave the search (|e|^2/2 = +inf), as codebook_refresh_kernel marks them
          for (int row = static_cast<int>(blockIdx.x) * RQ_WORKERS + w; row < p.K_total; row += static_cast<int>(gridDim.x) * RQ_WORKERS) {
            if (lane == 0 && __ldcg(p.ee_rw + row) == 0.f) {
              const int first = __float_as_int(__ldcg(meta_w + (row / p.K_per) * VQB200_LEVEL_META_FLOATS + 7));
              if (first > 0 && p.K_per - (row % p.K_per) != first) {
                p.ee_rw[row] = -kNegInf;
                p.ee_rw[p.K_total + row] = -kNegInf;
              }
            }
          }
          if (!last) {
            if (wt < BM) {
              float mt[VQB200_LEVEL_META_FLOATS];
#pragma unroll
              for (int i = 0; i < 6; ++i) mt[i] = __ldcg(meta_w + (l + 1) * VQB200_LEVEL_META_FLOATS + i);
              margin_s[wt] = rq_margin(BF16, D, best_s[wt * 2], best_s[wt * 2 + 1], mt);
            }
          } else if (blockIdx.x == 0 && wt < p.L * VQB200_LEVEL_META_FLOATS) {
            p.level_meta_rw[wt] = meta_w[wt];                // the cache's copy, in its canonical place
          }
        } else if (l + 1 < p.L) {
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
            if (p.hist && lane < p.L) atomicAdd(p.hist + res_s[lane * BM + r], 1);
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
        if (warp == 4) RQ_TR(1, l, 4);
        if (warp == 2) RQ_TR(2, l, 3);
        if (l + 1 < p.L) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_afull);
          if (TRAIN && warp == 2 && lane == 0) mbar_arrive(bar_go);   // the TMA producer may fetch the refreshed level
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
}

// ------------------------------------------------------------------------------------ host side
static int rq_smem_bytes(int BM, int BN, int D, int stages) {
  return 1024 + BM * D * 2 + stages * BN * TC_KB * 2 + rq_misc_bytes(BM, BN) + 8 + 24 * 8 + 64;
}

struct RqConfig { int BM, BN, stages, grid, smem; };

// Tile shape of a launch.  BM = 64 when the batch has so few 128-row tiles that most of a second wave of SMs would
// idle (VQB200_RVQ_BM=64|128 overrides); BN = 256 wherever the operand tile leaves room for >= 3 ring stages.
// Training mode needs every row tile resident at once (the grid synchronises between levels): one tile per CTA.
static bool rq_config(int64_t N, int D, RqConfig* c, bool train = false) {
  const int64_t tiles128 = (N + 127) / 128;
  int bm = tiles128 * 4 <= kNumSMs * 3 ? 64 : 128;
  if (const char* e = std::getenv("VQB200_RVQ_BM")) {
    if (e[0] == '6') bm = 64;
    else if (e[0] == '1') bm = 128;
  }
  if (train) {
    if ((N + 63) / 64 > kNumSMs) bm = 128;
    if ((N + bm - 1) / bm > kNumSMs) return false;
  }
  int bn = (bm == 64 || D <= 384) ? 256 : 128;
  int stages = 8;
  while (stages > 3 && rq_smem_bytes(bm, bn, D, stages) > TC_SMEM_LIMIT) --stages;
  if (rq_smem_bytes(bm, bn, D, stages) > TC_SMEM_LIMIT) return false;
  const int64_t tiles = (N + bm - 1) / bm;
  c->BM = bm; c->BN = bn; c->stages = stages;
  c->grid = static_cast<int>(tiles < kNumSMs ? tiles : kNumSMs);
  c->smem = rq_smem_bytes(bm, bn, D, stages);
  return true;
}

// Above this many rows the level-by-level pipeline (CTA-pair tensor kernel, HBM-bound passes between levels) is the
// faster one: the persistent kernel serialises search and row passes inside a CTA (measured, profiles/README.md).
static int64_t rq_max_rows() {
  if (const char* e = std::getenv("VQB200_RVQ_FUSED_MAX_ROWS")) return std::atoll(e);
  return 65536;
}

bool rvq_fused_supported(int64_t N, int K_per, int D, int L) {
  const char* f = std::getenv("VQB200_FORCE_SIMT");
  if (f && f[0] == '1') return false;
  const char* g = std::getenv("VQB200_NO_RVQ_FUSED");
  if (g && g[0] == '1') return false;
  if (D % 128 != 0 || D < 128 || D > 512 || K_per < TC_BN || L < 2 || L > RQ_MAXL || N < 1 || N > rq_max_rows()) return false;
  RqConfig c;
  return rq_config(N, D, &c);
}

bool rvq_fused_train_supported(int64_t N, int K_per, int D, int L) {
  const char* g = std::getenv("VQB200_NO_RVQ_FUSED_TRAIN");
  if (g && g[0] == '1') return false;
  if (!rvq_fused_supported(N, K_per, D, L)) return false;
  RqConfig c;
  return rq_config(N, D, &c, true);
}

size_t rvq_fused_workspace_bytes(int64_t N, int D) {
  // sized for either tile shape (the environment switch may change between the size query and the launch)
  const int64_t t64 = (N + 63) / 64, t128 = (N + 127) / 128;
  const size_t a = static_cast<size_t>(t64 < kNumSMs ? t64 : kNumSMs) * 64, b = static_cast<size_t>(t128 < kNumSMs ? t128 : kNumSMs) * 128;
  return 256 + (a > b ? a : b) * D * 4;
}

// training workspace: [0, 1024) counters, grid barrier, level-norm buffers | seg_sum | seg_cnt | residual scratch tiles
static size_t rq_train_seg_bytes(int K_total, int D) {
  return (static_cast<size_t>(K_total) * D * 4 + static_cast<size_t>(K_total) * 4 + 255) / 256 * 256;
}
size_t rvq_fused_train_workspace_bytes(int64_t N, int K_per, int D, int L) {
  return 1024 + rq_train_seg_bytes(K_per * L, D) + rvq_fused_workspace_bytes(N, D);
}

template <int SL, int BM, int BN, bool TRAIN>
static int launch_rq(const CUtensorMap& map_e, const RvqParams& p, bool bf, int grid, int smem, cudaStream_t s) {
  static bool attr_done_dev[64] = {};
  bool& attr_done = attr_done_dev[current_device_slot()];
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(rvq_fused_kernel<SL, false, BM, BN, TRAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(rvq_fused_kernel<SL, true, BM, BN, TRAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT);
    if (e != cudaSuccess) return status_of(e);
    attr_done = true;
  }
  timing_mark_begin(s);
  cudaError_t e = cudaSuccess;
  if (TRAIN) {
    // the grid synchronises between levels: every CTA must be resident (one per SM, grid <= SM count) -- cooperative launch
    void* args[2] = {const_cast<CUtensorMap*>(&map_e), const_cast<RvqParams*>(&p)};
    const void* fn = bf ? reinterpret_cast<const void*>(rvq_fused_kernel<SL, true, BM, BN, TRAIN>)
                        : reinterpret_cast<const void*>(rvq_fused_kernel<SL, false, BM, BN, TRAIN>);
    e = cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(RQ_THREADS), args, smem, s);
  } else {
    if (bf) rvq_fused_kernel<SL, true, BM, BN, TRAIN><<<grid, RQ_THREADS, smem, s>>>(map_e, p);
    else rvq_fused_kernel<SL, false, BM, BN, TRAIN><<<grid, RQ_THREADS, smem, s>>>(map_e, p);
  }
  timing_mark_end(s);
  return status_of(e != cudaSuccess ? e : cudaGetLastError());
}

template <int SL, bool TRAIN>
static int launch_rq_shape(const CUtensorMap& map_e, const RvqParams& p, bool bf, const RqConfig& c, cudaStream_t s) {
  if (c.BM == 64) return launch_rq<SL, 64, 256, TRAIN>(map_e, p, bf, c.grid, c.smem, s);
  if constexpr (SL == 4) return launch_rq<SL, 128, 128, TRAIN>(map_e, p, bf, c.grid, c.smem, s);
  else return launch_rq<SL, 128, 256, TRAIN>(map_e, p, bf, c.grid, c.smem, s);
}

static int rq_launch_common(RvqParams& p, const RqConfig& c, bool train, const void* E_lp, int K_per, int L, int D, bool bf,
                            cudaStream_t s) {
  p.row_tiles = static_cast<int>((p.n_rows + c.BM - 1) / c.BM);
  p.code_tiles = (K_per + c.BN - 1) / c.BN;
  p.stages = c.stages;
  // kind::f16 instruction descriptor: D = fp32, A = B = bf16 (bf16_input) or fp16, K-major, N = BN, M = BM
  p.idesc = (1u << 4) | (bf ? ((1u << 7) | (1u << 10)) : 0u) | ((static_cast<uint32_t>(c.BN) >> 3) << 17) |
            ((static_cast<uint32_t>(c.BM) >> 4) << 24);
  CUtensorMap map_e;
  if (!make_tensor_map_2d(&map_e, E_lp, static_cast<int64_t>(K_per) * L, D, c.BN,
                          bf ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2))
    return VQB200_EDRIVER;
  const char* dbg = std::getenv("VQB200_DEBUG");
  p.trace = nullptr;
  if (dbg && dbg[0] == '5') {
    cudaMallocManaged(&p.trace, 3 * RQ_MAXL * 8 * sizeof(long long));
    cudaMemset(p.trace, 0, 3 * RQ_MAXL * 8 * sizeof(long long));
  }
  int st;
  if (train) {
    switch (D / 128) {
      case 1: st = launch_rq_shape<1, true>(map_e, p, bf, c, s); break;
      case 2: st = launch_rq_shape<2, true>(map_e, p, bf, c, s); break;
      case 3: st = launch_rq_shape<3, true>(map_e, p, bf, c, s); break;
      default: st = launch_rq_shape<4, true>(map_e, p, bf, c, s); break;
    }
  } else {
    switch (D / 128) {
      case 1: st = launch_rq_shape<1, false>(map_e, p, bf, c, s); break;
      case 2: st = launch_rq_shape<2, false>(map_e, p, bf, c, s); break;
      case 3: st = launch_rq_shape<3, false>(map_e, p, bf, c, s); break;
      default: st = launch_rq_shape<4, false>(map_e, p, bf, c, s); break;
    }
  }
  if (p.trace) {                                                 // per-role timeline of CTA 0's first tile (cycles)
    cudaStreamSynchronize(s);
    long long t0 = 0;
    for (int i = 0; i < 3 * RQ_MAXL * 8; ++i) if (p.trace[i] && (!t0 || p.trace[i] < t0)) t0 = p.trace[i];
    static const char* names[3] = {"mma ", "scan", "help"};
    fprintf(stderr, "[vqb200] rvq fused%s: BM=%d BN=%d stages=%d grid=%d smem=%d\n", train ? " (training)" : "", c.BM, c.BN,
            c.stages, c.grid, c.smem);
    for (int l = 0; l < L; ++l)
      for (int r = 0; r < 3; ++r) {
        fprintf(stderr, "[vqb200] rvq level %d %s:", l, names[r]);
        for (int ev = 0; ev < 8; ++ev) {
          const long long v = p.trace[((r * RQ_MAXL) + l) * 8 + ev];
          if (v) fprintf(stderr, " e%d=%lld", ev, v - t0);
        }
        fprintf(stderr, "\n");
      }
    cudaFree(p.trace);
  }
  return st;
}

int launch_rvq_fused(const float* z, int64_t N, int D, const float* E, const uint16_t* E_lp, const float* ee_half,
                     const float* level_meta, int K_per, int L, int mode, int64_t* idx_out, float* zq_out,
                     float* zq_st_out, double* sqerr_sum, int32_t* hist, void* workspace, size_t workspace_bytes,
                     cudaStream_t s) {
  if (!rvq_fused_supported(N, K_per, D, L)) return VQB200_ESHAPE;
  if (workspace_bytes < rvq_fused_workspace_bytes(N, D)) return VQB200_EWORKSPACE;
  const bool bf = mode == VQB200_MODE_BF16_INPUT;
  RqConfig c;
  if (!rq_config(N, D, &c)) return VQB200_ESHAPE;
  RvqParams p{};
  p.n_rows = N; p.D = D; p.K_per = K_per; p.L = L; p.mode = mode;
  p.z = z; p.E = E; p.E_lp = E_lp; p.ee_half = ee_half; p.level_meta = level_meta;
  uint8_t* w = static_cast<uint8_t*>(workspace);
  p.counters = reinterpret_cast<int*>(w);
  p.scratch = reinterpret_cast<float*>(w + 256);
  p.idx_out = idx_out; p.zq_out = zq_out; p.zq_st_out = zq_st_out; p.sqerr_sum = sqerr_sum; p.hist = hist;
  cudaError_t e = cudaMemsetAsync(p.counters, 0, 2 * sizeof(int), s);
  if (e != cudaSuccess) return status_of(e);
  return rq_launch_common(p, c, false, E_lp, K_per, L, D, bf, s);
}

int launch_rvq_fused_train(const float* z, int64_t N, int D, float* E, uint16_t* E_lp_planes, float* ee_half,
                           float* level_meta, int K_per, int L, int mode, float decay, float omd, float eps,
                           float* ema_cs, float* ema_emb, int64_t* idx_out, float* zq_out, float* zq_st_out,
                           double* sqerr_sum, int32_t* hist, void* workspace, size_t workspace_bytes, cudaStream_t s) {
  if (!rvq_fused_train_supported(N, K_per, D, L)) return VQB200_ESHAPE;
  if (workspace_bytes < rvq_fused_train_workspace_bytes(N, K_per, D, L)) return VQB200_EWORKSPACE;
  const bool bf = mode == VQB200_MODE_BF16_INPUT;
  const int K_total = K_per * L;
  RqConfig c;
  if (!rq_config(N, D, &c, true)) return VQB200_ESHAPE;
  RvqParams p{};
  p.n_rows = N; p.D = D; p.K_per = K_per; p.L = L; p.mode = mode;
  p.z = z; p.E = E; p.level_meta = level_meta;
  p.plane_bf16 = E_lp_planes;
  p.plane_f16 = E_lp_planes + static_cast<size_t>(K_total) * D;
  p.E_lp = bf ? p.plane_bf16 : p.plane_f16;
  p.ee_half = bf ? ee_half + K_total : ee_half;
  p.E_rw = E; p.ee_rw = ee_half; p.level_meta_rw = level_meta;
  p.ema_cs = ema_cs; p.ema_emb = ema_emb; p.decay = decay; p.omd = omd; p.eps = eps; p.K_total = K_total;
  uint8_t* w = static_cast<uint8_t*>(workspace);
  p.counters = reinterpret_cast<int*>(w);
  p.grid_bar = reinterpret_cast<unsigned*>(w + 64);
  p.meta_buf = reinterpret_cast<float*>(w + 256);              // 2 x RQ_MAXL x 8 floats = 512 bytes
  p.seg_sum = reinterpret_cast<float*>(w + 1024);
  p.seg_cnt = p.seg_sum + static_cast<size_t>(K_total) * D;
  const size_t seg_bytes = rq_train_seg_bytes(K_total, D);
  p.scratch = reinterpret_cast<float*>(w + 1024 + seg_bytes + 256);
  p.idx_out = idx_out; p.zq_out = zq_out; p.zq_st_out = zq_st_out; p.sqerr_sum = sqerr_sum; p.hist = hist;
  cudaError_t e = cudaMemsetAsync(w, 0, 1024 + seg_bytes, s);
  if (e != cudaSuccess) return status_of(e);
  return rq_launch_common(p, c, true, p.E_lp, K_per, L, D, bf, s);
}

}  // namespace vqb
