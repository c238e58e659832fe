// Fully fused quantizer forward for small code dimensions (D = 64 and D = 128): ONE persistent kernel reads the
// fp32 latents once and writes indices, z_q, z_q_st, the commitment partial sum and the usage histogram -- the
// algorithmic HBM traffic of the path (12 D + 8 bytes per latent) and nothing else.  The multi-kernel
// pipeline of vq_search_tc.cu + gather crosses HBM/L2 four times per latent; at small D that, not the tensor
// work, bounds the step.
//
// Per CTA (one per SM, persistent over BM-row tiles; BM = 256 at D = 64, 128 at D = 128), 16 warps at BM = 256:
//   warp 0       TMA producer: fp32 latent tile [BM x D] (prefetched one tile ahead) + 16-bit codebook blocks
//   warp 1       MMA issuer:   tcgen05.mma M=128 N=128 K=16 per row half, accumulators double-buffered in TMEM and
//                              pre-loaded with -|e|^2/2 so the accumulator IS the score
//   warps 2-9    epilogue:     scan scores out of TMEM; groups of 8 columns that reach the running admission
//                              threshold go, raw scores and all, into a 4-entry shared-memory ring per row (its
//                              ageing is a register shift chain); then per tile: prune against the final threshold,
//                              score multi-survivor rows exactly (fp64, lane groups, inputs from L2), write idx and
//                              the histogram, publish the tile's codes in shared memory
//   warps 10-15  converters:   fp32 tile (swizzled TMA layout) -> fp16 (fp32 mode) / bf16 UMMA operand tile
//                / output      (128-byte swizzle), measured conversion error -> admission margins; then, one tile
//                              behind the epilogue: z_q = E[idx], z_q_st = fl(z + fl(z_q - z)), sum (z_q - z)^2
// Exactness argument, margin and hand-back rules are those of vq_search_tc.cu / common.cuh.
#include <cstdio>
#include <initializer_list>
#include <cstdlib>

#include "tc_ptx.cuh"

namespace vqb {

// Tile heights.  BM = 256 (D = 64): one CTA per SM, 8 epilogue + 6 converter/output warps (16 warps cap every
// thread at 128 registers).  BM = 128: D = 128 (the fp32 staging tile of 256 rows would not fit), one CTA per SM;
// and, as a measurement switch at D = 64 (VQB200_FUSED_BM=128), TWO co-resident CTAs per SM with 256 TMEM columns
// each -- measured slower than BM = 256: the same number of epilogue warps per SM, twice the per-tile overheads.
// converter / output warps: 6 where the CTA has the SM to itself (16 warps at BM = 256, 12 at D = 128), 2 in the
// two-CTAs-per-SM variant (D = 64, BM = 128)
constexpr int fz_nconv(int D, int BM) { return (D == 64 && BM == 128) ? 2 : 6; }
constexpr int fz_nepi(int BM) { return BM / 32; }
constexpr int fz_threads(int D, int BM) { return 64 + fz_nepi(BM) * 32 + fz_nconv(D, BM) * 32; }   // 512 / 384 / 256
constexpr int FZ_PAIRS = 64;          // (row, code) pairs scored per warp pass
constexpr int fz_hist(int BM) { return BM == 256 ? 2048 : 1024; }   // codebooks up to this size: shared-memory histogram
constexpr int FZ_RING_BYTES = 4 * (2 * 32 * 16 + 32 * 8);           // per warp: FZ_RING entries x 32 lanes x 40 bytes
constexpr uint32_t FZ_EMPTY = 0xffffffffu;

struct FusedParams {
  int64_t n_rows;
  int K, row_tiles, code_tiles, stages, mode;
  const float* z;               // [n_rows, D] fp32
  const float* E;               // [K, D] fp32
  const __nv_bfloat16* Eb;      // [K, D] bf16
  const float* ee_half;         // [K] plane chosen by mode
  const float* level_meta;
  int64_t idx_offset;
  int64_t* idx_out;
  float* zq_out;
  float* zq_st_out;
  double* sqerr_sum;
  int* hist;                    // [K_total] (+ idx_offset applied)
  const uint8_t* row_mask;
  int* fb_rows;
  uint64_t* fb_packed;
  int* counters;
  long long* dbg;               // VQB200_DEBUG=2: per CTA [smid, globaltimer at start, at end]
};

// 32 score columns of one row.  Fast path: one compare against the admission threshold.  Slow path: every group
// of 8 columns whose maximum reaches the threshold is appended, raw scores and all, to the lane's ring of
// FZ_RING entries in shared memory (lane-interleaved 16-byte fields: conflict-free for any per-lane ring
// position).  No mask is built here -- at the end of the row tile the survivors are re-tested against the FINAL
// threshold.  Overwriting an entry remembers its maximum: if that could still matter the row is handed back.
constexpr int FZ_RING = 4;
__device__ __forceinline__ void fz_scan(const uint32_t (&v)[32], uint32_t code0, float margin, float& best, float& thr,
                                        uint4* ring, uint2* rmeta, int lane, int& cnt, float& lost,
                                        float (&age)[FZ_RING]) {
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
        uint4* e = ring + ((cnt & (FZ_RING - 1)) * 2) * 32 + lane;     // score fields at e[0], e[32]
        uint2* m = rmeta + (cnt & (FZ_RING - 1)) * 32 + lane;           // (group maximum, group id)
        lost = fmaxf(lost, age[FZ_RING - 1]);        // the entry being overwritten is the oldest: its maximum
#pragma unroll                                        // sits at the end of a register shift chain (no LDS on this path)
        for (int a = FZ_RING - 1; a > 0; --a) age[a] = age[a - 1];
        age[0] = gm[g];
        e[0] = make_uint4(v[g * 8 + 0], v[g * 8 + 1], v[g * 8 + 2], v[g * 8 + 3]);
        e[32] = make_uint4(v[g * 8 + 4], v[g * 8 + 5], v[g * 8 + 6], v[g * 8 + 7]);
        *m = make_uint2(__float_as_uint(gm[g]), (code0 >> 3) + g);
        ++cnt;
      }
    }
  }
}

template <int D, bool BF16, int FZ_BM>
__global__ void __launch_bounds__(fz_threads(D, FZ_BM), (D == 64 && FZ_BM == 128) ? 2 : 1)
quantize_fused_kernel(const __grid_constant__ CUtensorMap tmap_zf, const __grid_constant__ CUtensorMap tmap_e,
                      const FusedParams p) {
  constexpr int FZ_NEPI = fz_nepi(FZ_BM), FZ_THREADS = fz_threads(D, FZ_BM), FZ_HIST = fz_hist(FZ_BM);
  constexpr int FZ_NCONV = fz_nconv(D, FZ_BM);
  constexpr int HALVES = FZ_BM / 128;              // M = 128 accumulator tiles per row tile
  constexpr uint32_t TMEM_COLS = 2 * HALVES * TC_BN;   // two accumulator buffers
  constexpr int KBLK = D / TC_KB;                  // bf16 operand blocks along D
  constexpr int KB32 = D / 32;                     // fp32 staging slabs along D (32 floats = 128 bytes)
  constexpr uint32_t ZF_BYTES = FZ_BM * D * 4;
  constexpr uint32_t ZB_BYTES = FZ_BM * D * 2;
  constexpr int LPV = (D / 4) < 32 ? (D / 4) : 32; // lanes that cover one row with float4 slices
  constexpr int GROUPS = 32 / LPV;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t zf_smem = base;
  const uint32_t zb_smem = zf_smem + ZF_BYTES;
  const uint32_t e_smem = zb_smem + ZB_BYTES;               // ONE bf16 operand tile: shared memory goes to the rings
  const uint32_t misc = e_smem + static_cast<uint32_t>(p.stages) * TC_STAGE_BYTES;
  float* margin_s = reinterpret_cast<float*>(gen + (misc - base));                       // [2][256]
  float* ee_slots = margin_s + 2 * FZ_BM;                                                 // [8][128]
  uint32_t* pair_list = reinterpret_cast<uint32_t*>(ee_slots + FZ_NEPI * TC_BN);          // [8][64]
  double* pair_score = reinterpret_cast<double*>(pair_list + FZ_NEPI * FZ_PAIRS);         // [8][64]
  int* hist_s = reinterpret_cast<int*>(pair_score + FZ_NEPI * FZ_PAIRS);                  // [FZ_HIST]
  double* red_s = reinterpret_cast<double*>(hist_s + FZ_HIST);                            // [16]
  uint8_t* rings = reinterpret_cast<uint8_t*>(red_s + 16);                                // [NEPI][FZ_RING_BYTES]
  uint32_t* res_s = reinterpret_cast<uint32_t*>(rings + FZ_NEPI * FZ_RING_BYTES);         // [2][FZ_BM] resolved codes
  const uint32_t bar0 = misc + 2 * FZ_BM * 4 + FZ_NEPI * TC_BN * 4 + FZ_NEPI * FZ_PAIRS * 4 + FZ_NEPI * FZ_PAIRS * 8 +
                        FZ_HIST * 4 + 16 * 8 + FZ_NEPI * FZ_RING_BYTES + 2 * FZ_BM * 4;
  const uint32_t bar_full = bar0, bar_empty = bar0 + 8 * 8, bar_tfull = bar0 + 16 * 8, bar_tempty = bar0 + 18 * 8;
  const uint32_t bar_zffull = bar0 + 20 * 8, bar_zfempty = bar0 + 21 * 8;
  const uint32_t bar_zbfull = bar0 + 22 * 8, bar_zbempty = bar0 + 24 * 8;    // [2] each
  const uint32_t bar_resfull = bar0 + 26 * 8, bar_resempty = bar0 + 28 * 8;  // [2] each
  const uint32_t tmem_slot = bar0 + 30 * 8;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool smem_hist = p.hist != nullptr && p.K <= FZ_HIST;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_tfull + 8 * b, 1); mbar_init(bar_tempty + 8 * b, FZ_NEPI);
      mbar_init(bar_zbfull + 8 * b, FZ_NCONV); mbar_init(bar_zbempty + 8 * b, 1);
      mbar_init(bar_resfull + 8 * b, FZ_NEPI); mbar_init(bar_resempty + 8 * b, FZ_NCONV);
    }
    mbar_init(bar_zffull, 1);
    mbar_init(bar_zfempty, FZ_NCONV);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (smem_hist)
    for (int k = threadIdx.x; k < p.K; k += FZ_THREADS) hist_s[k] = 0;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

  const int n_items = p.row_tiles;
  if (p.dbg && threadIdx.x == 0) {
    unsigned id; asm volatile("mov.u32 %0, %%smid;" : "=r"(id));
    long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
    p.dbg[blockIdx.x * 3] = id; p.dbg[blockIdx.x * 3 + 1] = g;
  }
  float err_acc = 0.f;                              // sum (z_q - z)^2 over the rows this thread finalises

  if (warp == 0) {
    // ============================== TMA producer ==============================
    uint32_t stage = 0, phase = 0, it = 0;
    auto load_z = [&](int item, uint32_t seq) {
      mbar_wait(bar_zfempty, (seq & 1) ^ 1);
      if (elect_one()) {
        mbar_expect_tx(bar_zffull, ZF_BYTES);
        for (int kb = 0; kb < KB32; ++kb)
          tma_load_2d(zf_smem + kb * (FZ_BM * 128), &tmap_zf, bar_zffull, kb * 32, item * FZ_BM);
      }
      __syncwarp();
    };
    if (static_cast<int>(blockIdx.x) < n_items) load_z(blockIdx.x, 0);
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int nxt = item + gridDim.x;
      if (nxt < n_items) load_z(nxt, it + 1);       // the next tile's latents ride ahead of this tile's codebook blocks
      for (int t = 0; t < p.code_tiles; ++t) {
        for (int kb = 0; kb < KBLK; ++kb) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          if (elect_one()) {
            mbar_expect_tx(bar_full + 8 * stage, TC_STAGE_BYTES);
            tma_load_2d(e_smem + stage * TC_STAGE_BYTES, &tmap_e, bar_full + 8 * stage, kb * TC_KB, t * TC_BN);
          }
          __syncwarp();
          if (++stage == static_cast<uint32_t>(p.stages)) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    uint32_t stage = 0, phase = 0, it = 0, tg = 0;
    const uint32_t zb_lo = umma_desc_lo(zb_smem), e_lo = umma_desc_lo(e_smem);
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const uint32_t zb = 0;
      mbar_wait(bar_zbfull, it & 1);
      for (int t = 0; t < p.code_tiles; ++t, ++tg) {
        const uint32_t b = tg & 1;
        mbar_wait(bar_tempty + 8 * b, (tg >> 1) & 1);       // drained AND pre-loaded with -|e|^2/2
        tc_fence_after();
        for (int kb = 0; kb < KBLK; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a0 = zb_lo + ((zb * ZB_BYTES + kb * (FZ_BM * 128)) >> 4);
            const uint32_t b0 = e_lo + ((stage * TC_STAGE_BYTES) >> 4);
#pragma unroll
            for (int h = 0; h < HALVES; ++h) {
              const uint32_t d_tmem = tmem_base + b * (HALVES * TC_BN) + h * TC_BN;
#pragma unroll
              for (int k = 0; k < TC_KB / 16; ++k)
                tc_mma_bf16(d_tmem, umma_desc(a0 + h * ((128 * 128) >> 4) + k * 2), umma_desc(b0 + k * 2), BF16 ? kIdesc : kIdescF16, 1u);
            }
            tc_commit(bar_empty + 8 * stage);
            if (kb == KBLK - 1) tc_commit(bar_tfull + 8 * b);
          }
          __syncwarp();
          if (++stage == static_cast<uint32_t>(p.stages)) { stage = 0; phase ^= 1; }
        }
      }
      if (elect_one()) tc_commit(bar_zbempty);
      __syncwarp();
    }
  } else if (warp >= 2 + FZ_NEPI) {
    // ============================== converters: fp32 tile -> bf16 operand tile + margins ==============================
    const int ct = (warp - 2 - FZ_NEPI) * 32 + lane;     // 0..63: rows ct, ct+64, ct+128, ct+192 of the tile
    const float emax = BF16 ? p.level_meta[2] : p.level_meta[0];
    const float emax_lp = p.level_meta[4], rho_e = p.level_meta[5];
    const bool code_bad = p.level_meta[1] != 0.f;
    const float coef = 2.f * static_cast<float>(D + 32) * 1.1920929e-7f;      // bf16 mode (see zprep_kernel)
    // Output phase of a finished tile (z_q, z_q_st, squared error), run by these warps one tile behind the
    // epilogue: the epilogue warps publish the tile's resolved codes in shared memory and move on to the next
    // tile's scan, so the L2 / HBM round trips of the outputs are off the scan -> re-rank critical path.
    const bool want_out = p.zq_out || p.zq_st_out || p.sqerr_sum;
    auto emit = [&](int item_o, uint32_t ito) {
      constexpr int RPS = (FZ_NCONV * 32) / LPV;             // rows per step of the two warps together
      constexpr int OB = 8;       // steps whose loads are issued together (measured: 16 is slower, 4 warps cost the
                                  // epilogue its registers: 13-16 warps cap every thread at 128)
      constexpr int SL = D / (LPV * 4);
      const uint32_t rbuf = ito & 1;
      mbar_wait(bar_resfull + 8 * rbuf, (ito >> 1) & 1);
      const uint32_t* rs = res_s + rbuf * FZ_BM;
      const int64_t row0 = static_cast<int64_t>(item_o) * FZ_BM;
      const int ogl = ct % LPV, ogr = ct / LPV;
#pragma unroll 1
      for (int rb = 0; rb < FZ_BM; rb += RPS * OB) {
        float4 e4[OB][SL], z4[OB][SL];
        bool res[OB];
#pragma unroll
        for (int u = 0; u < OB; ++u) {
          const int rl = rb + u * RPS + ogr;
          const uint32_t code = rl < FZ_BM ? rs[rl] : FZ_EMPTY;
          const int64_t grow = row0 + rl;
          res[u] = code != FZ_EMPTY;
#pragma unroll
          for (int sl = 0; sl < SL; ++sl) {
            const int d = ogl * 4 + sl * LPV * 4;
            if (res[u]) {
              e4[u][sl] = __ldg(reinterpret_cast<const float4*>(p.E + static_cast<int64_t>(code) * D + d));
              z4[u][sl] = ld_stream(reinterpret_cast<const float4*>(p.z + grow * D + d));
            }
          }
        }
#pragma unroll
        for (int u = 0; u < OB; ++u) {
          if (!res[u]) continue;
          const int64_t grow = row0 + rb + u * RPS + ogr;
#pragma unroll
          for (int sl = 0; sl < SL; ++sl) {
            const int d = ogl * 4 + sl * LPV * 4;
            const float4 e = e4[u][sl], zz = z4[u][sl];
            float4 df;
            df.x = __fsub_rn(e.x, zz.x); df.y = __fsub_rn(e.y, zz.y);
            df.z = __fsub_rn(e.z, zz.z); df.w = __fsub_rn(e.w, zz.w);
            if (p.zq_out) st_stream(reinterpret_cast<float4*>(p.zq_out + grow * D + d), e);
            if (p.zq_st_out) {
              float4 o;
              o.x = __fadd_rn(zz.x, df.x); o.y = __fadd_rn(zz.y, df.y);
              o.z = __fadd_rn(zz.z, df.z); o.w = __fadd_rn(zz.w, df.w);
              st_stream(reinterpret_cast<float4*>(p.zq_st_out + grow * D + d), o);
            }
            err_acc = fmaf(df.x, df.x, err_acc); err_acc = fmaf(df.y, df.y, err_acc);
            err_acc = fmaf(df.z, df.z, err_acc); err_acc = fmaf(df.w, df.w, err_acc);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_resempty + 8 * rbuf);
    };
    uint32_t it = 0;
    int item_prev = -1;
    for (int item = blockIdx.x; item < n_items; item_prev = item, item += gridDim.x, ++it) {
      const uint32_t zb = 0, mb = it & 1;
      mbar_wait(bar_zffull, it & 1);
      mbar_wait(bar_zbempty, (it & 1) ^ 1);             // the MMAs of the previous tile have retired
#pragma unroll 1
      for (int rr = 0; rr < (FZ_BM + FZ_NCONV * 32 - 1) / (FZ_NCONV * 32); ++rr) {
        const int r = ct + rr * (FZ_NCONV * 32);
        if (r >= FZ_BM) break;
        const uint32_t sw = static_cast<uint32_t>(r & 7);
        float ss = 0.f, sse = 0.f;
#pragma unroll
        for (int q = 0; q < D / 8; ++q) {             // one 16-byte bf16 chunk (8 elements) per step
          const int c = q * 8;                        // first column of the chunk
          float4 lo, hi;
          {
            const int slab = c / 32, j = (c % 32) / 4;
            const uint8_t* src = gen + (zf_smem - base) + slab * (FZ_BM * 128) + r * 128;
            lo = *reinterpret_cast<const float4*>(src + ((static_cast<uint32_t>(j) ^ sw) << 4));
            hi = *reinterpret_cast<const float4*>(src + ((static_cast<uint32_t>(j + 1) ^ sw) << 4));
          }
          uint4 pk;
          float2 a, b2, c2, d2;                       // the stored operand values, back in fp32
          if (BF16) {
            const __nv_bfloat162 p0 = __floats2bfloat162_rn(lo.x, lo.y), p1 = __floats2bfloat162_rn(lo.z, lo.w),
                                 p2 = __floats2bfloat162_rn(hi.x, hi.y), p3 = __floats2bfloat162_rn(hi.z, hi.w);
            pk.x = *reinterpret_cast<const uint32_t*>(&p0); pk.y = *reinterpret_cast<const uint32_t*>(&p1);
            pk.z = *reinterpret_cast<const uint32_t*>(&p2); pk.w = *reinterpret_cast<const uint32_t*>(&p3);
            a = __bfloat1622float2(p0); b2 = __bfloat1622float2(p1); c2 = __bfloat1622float2(p2); d2 = __bfloat1622float2(p3);
          } else {                                    // fp32 mode: fp16 operands, subnormal results flushed
            pk.x = f16x2_bits_flush(lo.x, lo.y); pk.y = f16x2_bits_flush(lo.z, lo.w);
            pk.z = f16x2_bits_flush(hi.x, hi.y); pk.w = f16x2_bits_flush(hi.z, hi.w);
            a = f16x2_bits_to_float2(pk.x); b2 = f16x2_bits_to_float2(pk.y);
            c2 = f16x2_bits_to_float2(pk.z); d2 = f16x2_bits_to_float2(pk.w);
          }
          const int bslab = c / TC_KB, bj = (c % TC_KB) / 8;
          uint8_t* dst = gen + (zb_smem - base) + zb * ZB_BYTES + bslab * (FZ_BM * 128) + r * 128;
          *reinterpret_cast<uint4*>(dst + ((static_cast<uint32_t>(bj) ^ sw) << 4)) = pk;
          if (BF16) {
            ss += a.x * a.x + a.y * a.y + b2.x * b2.x + b2.y * b2.y + c2.x * c2.x + c2.y * c2.y + d2.x * d2.x + d2.y * d2.y;
          } else {
            ss += lo.x * lo.x + lo.y * lo.y + lo.z * lo.z + lo.w * lo.w + hi.x * hi.x + hi.y * hi.y + hi.z * hi.z + hi.w * hi.w;
            const float e0 = lo.x - a.x, e1 = lo.y - a.y, e2 = lo.z - b2.x, e3 = lo.w - b2.y,      // exact differences
                        e4 = hi.x - c2.x, e5 = hi.y - c2.y, e6 = hi.z - d2.x, e7 = hi.w - d2.y;
            sse += e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3 + e4 * e4 + e5 * e5 + e6 * e6 + e7 * e7;
          }
        }
        float m = BF16 ? coef * (sqrtf(ss) * 1.0001f) * emax + 1e-30f : admission_margin_fp32(ss, sse, emax, emax_lp, rho_e, D);
        if (code_bad || !(ss < __int_as_float(0x7f800000)) || !(sse < __int_as_float(0x7f800000)))
          m = __int_as_float(0x7fc00000);   // NaN: exact path
        margin_s[mb * FZ_BM + r] = m;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tensor-core reads
      __syncwarp();
      if (lane == 0) { mbar_arrive(bar_zbfull); mbar_arrive(bar_zfempty); }
      if (want_out && item_prev >= 0) emit(item_prev, it - 1);
    }
    if (want_out && item_prev >= 0) emit(item_prev, it - 1);
  } else {
    // ============================== epilogue + finalise ==============================
    const int we = warp - 2;
    const int quarter = warp & 3, half = we >> 2;
    const int row_in_tile = half * 128 + quarter * 32;
    float* ee_slot = ee_slots + we * TC_BN;
    uint4* ring = reinterpret_cast<uint4*>(rings + we * FZ_RING_BYTES);
    uint2* rmeta = reinterpret_cast<uint2*>(rings + we * FZ_RING_BYTES + FZ_RING * 2 * 32 * 16);
    uint32_t* plist = pair_list + we * FZ_PAIRS;
    double* pscore = pair_score + we * FZ_PAIRS;
    const float kNegInf = __int_as_float(0xff800000);
    const uint32_t tcol = (static_cast<uint32_t>(quarter * 32) << 16) + half * TC_BN;
    const int gl = lane % LPV, gi = lane / LPV;

    struct Bias { float v[4]; };
    auto load_bias = [&](int t) -> Bias {
      Bias r;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = t * TC_BN + lane * 4 + j;
        r.v[j] = (t >= 0 && c < p.K) ? -p.ee_half[c] : kNegInf;
      }
      return r;
    };
    auto preload = [&](const Bias& bias, uint32_t b) {
      __syncwarp();
      *reinterpret_cast<float4*>(ee_slot + lane * 4) = make_float4(bias.v[0], bias.v[1], bias.v[2], bias.v[3]);
      __syncwarp();
#pragma unroll 2
      for (int hh = 0; hh < TC_BN / 16; ++hh) {
        uint32_t w[16];
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 q4 = *reinterpret_cast<const float4*>(ee_slot + hh * 16 + j);
          w[j + 0] = __float_as_uint(q4.x); w[j + 1] = __float_as_uint(q4.y);
          w[j + 2] = __float_as_uint(q4.z); w[j + 3] = __float_as_uint(q4.w);
        }
        TC_ST16(tmem_base + tcol + b * (HALVES * TC_BN) + hh * 16, w);
      }
      tc_wait_st();
    };
    // the CTA's tile sequence is item-major, code tiles 0..code_tiles-1 inside; walk it two tiles ahead
    const int64_t total_tiles = static_cast<int64_t>((n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                                                     static_cast<int>(gridDim.x)) * p.code_tiles;
    int64_t la = 0;                                   // index into the sequence of the next tile to fetch
    int la_t = 0;                                     // its code tile (kept incrementally: no division per tile)
    auto la_next = [&]() -> int {
      if (la++ >= total_tiles) return -1;
      const int t = la_t;
      la_t = (la_t + 1 == p.code_tiles) ? 0 : la_t + 1;
      return t;
    };
    for (uint32_t b = 0; b < 2; ++b) {
      const int tt = la_next();
      if (tt >= 0) preload(load_bias(tt), b);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * b);
    }
    int t_ahead = la_next();
    Bias bias_next = load_bias(t_ahead);

    uint32_t tg = 0, it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int64_t row0w = static_cast<int64_t>(item) * FZ_BM + row_in_tile;
      const int64_t row = row0w + lane;
      const bool row_ok = row < p.n_rows;
      mbar_wait(bar_zbfull, it & 1);                          // margins of this tile are in shared memory
      const float margin = row_ok ? margin_s[(it & 1) * FZ_BM + row_in_tile + lane] : __int_as_float(0x7fc00000);
      float best = kNegInf;
      float thr = margin == margin ? -3.0e38f : margin;   // lowest finite value: -inf chunks are never admitted
      int rcnt = 0;
      float lost = kNegInf;
      float age[FZ_RING] = {kNegInf, kNegInf, kNegInf, kNegInf};

      for (int t = 0; t < p.code_tiles; ++t, ++tg) {
        const uint32_t b = tg & 1;
        const Bias bias = bias_next;
        const int t_cur_ahead = t_ahead;
        t_ahead = la_next();
        bias_next = load_bias(t_ahead);
        mbar_wait(bar_tfull + 8 * b, (tg >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + tcol + b * (HALVES * TC_BN);
        // One 32-column chunk in registers at a time (measured: double-buffering the TMEM loads buys nothing
        // here, the scan is bound by its own dependency chains).  After the LAST chunk is out of TMEM the
        // buffer is re-armed with the bias of its next tile and handed back; only then is that chunk scanned.
        uint32_t v[32];
#pragma unroll
        for (int ch = 0; ch < TC_BN / 32; ++ch) {
          TC_LD32(taddr + ch * 32, v);
          tc_wait_ld();
          if (ch == TC_BN / 32 - 1) {
            if (t_cur_ahead >= 0) preload(bias, b);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * b);
          }
          fz_scan(v, static_cast<uint32_t>(t * TC_BN + ch * 32), margin, best, thr, ring, rmeta, lane, rcnt, lost, age);
        }
      }

      // ---------------- finalise the 32 rows of this warp ----------------
      // survivors: ring entries whose maximum reaches the FINAL threshold; their admit masks are built now,
      // against that threshold, from the stored raw scores
      int ns = 0, ncodes = 0;
      uint32_t first = 0;
      uint32_t sx[FZ_RING];
      const bool overflow = lost >= thr;                        // an overwritten entry could still matter
#pragma unroll
      for (int s = 0; s < FZ_RING; ++s) {
        sx[s] = FZ_EMPTY;
        if (s < rcnt) {
          const uint4* e = ring + (s * 2) * 32 + lane;
          const uint2 meta = rmeta[s * 32 + lane];
          if (__uint_as_float(meta.x) >= thr) {
            const uint4 a = e[0], b4 = e[32];
            uint32_t mk = 0;
            mk |= (__uint_as_float(a.x) >= thr) ? 1u : 0u;   mk |= (__uint_as_float(a.y) >= thr) ? 2u : 0u;
            mk |= (__uint_as_float(a.z) >= thr) ? 4u : 0u;   mk |= (__uint_as_float(a.w) >= thr) ? 8u : 0u;
            mk |= (__uint_as_float(b4.x) >= thr) ? 16u : 0u; mk |= (__uint_as_float(b4.y) >= thr) ? 32u : 0u;
            mk |= (__uint_as_float(b4.z) >= thr) ? 64u : 0u; mk |= (__uint_as_float(b4.w) >= thr) ? 128u : 0u;
            sx[s] = (meta.y << 8) | mk;
            if (ns == 0) first = sx[s];
            ++ns;
            ncodes += __popc(mk);
          }
        }
      }
      bool resolved = false, multi = false;
      uint32_t my_idx = 0;
      if (row_ok) {
        if (overflow || ns == 0 || !(margin == margin)) {       // exact SIMT kernel takes the row
          const int pos = atomicAdd(p.counters, 1);
          p.fb_rows[pos] = static_cast<int>(row);
          p.fb_packed[row] = ~0ull;
        } else if (ns == 1 && ncodes == 1) {                    // certified by the error bound
          my_idx = ((first >> 8) << 3) + (__ffs(first & 0xffu) - 1);
          resolved = true;
        } else {
          multi = true;
        }
      }
      bool pending = multi;
      while (__any_sync(0xffffffffu, pending)) {
        const int c = pending ? ncodes : 0;
        int pre = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int u = __shfl_up_sync(0xffffffffu, pre, o);
          if (lane >= o) pre += u;
        }
        const bool take = pending && pre <= FZ_PAIRS;
        const int off0 = pre - c;
        if (take) {
          int at = off0;
#pragma unroll
          for (int s = 0; s < FZ_RING; ++s) {
            if (sx[s] != FZ_EMPTY) {
              uint32_t m = sx[s] & 0xffu;
              while (m) {
                plist[at++] = (static_cast<uint32_t>(lane) << 24) | (((sx[s] >> 8) << 3) + (__ffs(m) - 1));
                m &= m - 1;
              }
            }
          }
        }
        const unsigned tk = __ballot_sync(0xffffffffu, take);
        const int T = __shfl_sync(0xffffffffu, pre, 31 - __clz(tk));
        __syncwarp();
        // GROUPS pairs are scored concurrently by lane groups, SB such steps are batched so that their
        // L2 loads are all in flight before the first FMA (the chain is latency-, not bandwidth-bound)
        constexpr int SB = 4;
        constexpr int SL = D / (LPV * 4);                     // float4 slices per lane per row
        for (int pb = 0; pb < T; pb += GROUPS * SB) {
          float4 zq4[SB][SL], eq4[SB][SL];
          uint2 eb2[SB][SL];
          bool okb[SB];
#pragma unroll
          for (int u = 0; u < SB; ++u) {
            const int pi = pb + u * GROUPS + gi;
            okb[u] = pi < T;
            const uint32_t pr = okb[u] ? plist[pi] : 0u;
            const uint32_t code = pr & 0xffffffu;
            const int64_t grow = row0w + (pr >> 24);
#pragma unroll
            for (int sl = 0; sl < SL; ++sl) {
              const int d = gl * 4 + sl * LPV * 4;
              if (okb[u]) {
                zq4[u][sl] = *reinterpret_cast<const float4*>(p.z + grow * D + d);
                if (BF16)
                  eb2[u][sl] = *reinterpret_cast<const uint2*>(p.Eb + static_cast<int64_t>(code) * D + d);
                else
                  eq4[u][sl] = __ldg(reinterpret_cast<const float4*>(p.E + static_cast<int64_t>(code) * D + d));
              }
            }
          }
#pragma unroll
          for (int u = 0; u < SB; ++u) {
            double dot = 0.0, ee = 0.0;
            if (okb[u]) {
#pragma unroll
              for (int sl = 0; sl < SL; ++sl) {
                float zv[4] = {zq4[u][sl].x, zq4[u][sl].y, zq4[u][sl].z, zq4[u][sl].w};
                float ev[4];
                if (BF16) {
                  ev[0] = __uint_as_float(eb2[u][sl].x << 16); ev[1] = __uint_as_float(eb2[u][sl].x & 0xffff0000u);
                  ev[2] = __uint_as_float(eb2[u][sl].y << 16); ev[3] = __uint_as_float(eb2[u][sl].y & 0xffff0000u);
#pragma unroll
                  for (int q = 0; q < 4; ++q) zv[q] = bf16_round(zv[q]);
                } else {
                  ev[0] = eq4[u][sl].x; ev[1] = eq4[u][sl].y; ev[2] = eq4[u][sl].z; ev[3] = eq4[u][sl].w;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  dot = fma(static_cast<double>(zv[q]), static_cast<double>(ev[q]), dot);
                  ee = fma(static_cast<double>(ev[q]), static_cast<double>(ev[q]), ee);
                }
              }
            }
#pragma unroll
            for (int o = LPV >> 1; o > 0; o >>= 1) {
              dot += __shfl_xor_sync(0xffffffffu, dot, o);
              ee += __shfl_xor_sync(0xffffffffu, ee, o);
            }
            if (okb[u] && gl == 0) pscore[pb + u * GROUPS + gi] = dot - 0.5 * ee;
          }
        }
        __syncwarp();
        if (take) {                                             // lexicographic (score, lowest index)
          double top = -1e300;
          uint32_t top_idx = 0xffffffffu;
          for (int j = off0; j < off0 + c; ++j) {
            const double sc = pscore[j];
            const uint32_t cd = plist[j] & 0xffffffu;
            if (sc > top || (sc == top && cd < top_idx)) { top = sc; top_idx = cd; }
          }
          my_idx = top_idx;
          resolved = true;
        }
        pending = pending && !take;
        __syncwarp();
      }

      // ---------------- outputs ----------------
      if (resolved) {
        p.idx_out[row] = p.idx_offset + my_idx;
        if (p.hist && (!p.row_mask || p.row_mask[row])) {
          if (smem_hist) atomicAdd(hist_s + my_idx, 1);
          else atomicAdd(p.hist + p.idx_offset + my_idx, 1);
        }
      }
      if (p.zq_out || p.zq_st_out || p.sqerr_sum) {              // publish the codes; the output warps take over
        const uint32_t rbuf = it & 1, use = it >> 1;
        if (use >= 1) mbar_wait(bar_resempty + 8 * rbuf, (use - 1) & 1);
        res_s[rbuf * FZ_BM + row_in_tile + lane] = resolved ? my_idx : FZ_EMPTY;
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_resfull + 8 * rbuf);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (p.dbg && threadIdx.x == 0) {
    long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
    p.dbg[blockIdx.x * 3 + 2] = g;
  }
  if (smem_hist)
    for (int k = threadIdx.x; k < p.K; k += FZ_THREADS) {
      const int v = hist_s[k];
      if (v) atomicAdd(p.hist + p.idx_offset + k, v);
    }
  if (p.sqerr_sum) {
    const double e = warp_sum(static_cast<double>(err_acc));
    if (lane == 0) red_s[warp] = e;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < FZ_THREADS / 32; ++w) t += red_s[w];
      atomicAdd(p.sqerr_sum, t);
    }
  }
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// Rows handed back to the exact SIMT kernel: write their index and outputs once it has decided.
__global__ void __launch_bounds__(256)
fused_fixup_kernel(const int* __restrict__ fb_rows, const uint64_t* __restrict__ fb_packed,
                   const int* __restrict__ counters, const float* __restrict__ z, const float* __restrict__ E, int D,
                   int64_t* __restrict__ idx_out, float* __restrict__ zq_out, float* __restrict__ zq_st_out,
                   double* sqerr_sum, int* __restrict__ hist, const uint8_t* __restrict__ row_mask) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  const int n = counters[0];
  float err = 0.f;
  for (int i = warp; i < n; i += nwarps) {
    const int64_t row = fb_rows[i];
    const int64_t code = static_cast<int64_t>(fb_packed[row] & 0xffffffffull);   // idx_offset already applied
    if (lane == 0) {
      idx_out[row] = code;
      if (hist && (!row_mask || row_mask[row])) atomicAdd(hist + code, 1);
    }
    for (int d = lane * 4; d < D; d += 128) {
      const float4 e4 = __ldg(reinterpret_cast<const float4*>(E + code * D + d));
      const float4 z4 = *reinterpret_cast<const float4*>(z + row * D + d);
      float4 df;
      df.x = __fsub_rn(e4.x, z4.x); df.y = __fsub_rn(e4.y, z4.y); df.z = __fsub_rn(e4.z, z4.z); df.w = __fsub_rn(e4.w, z4.w);
      if (zq_out) *reinterpret_cast<float4*>(zq_out + row * D + d) = e4;
      if (zq_st_out)
        *reinterpret_cast<float4*>(zq_st_out + row * D + d) =
            make_float4(__fadd_rn(z4.x, df.x), __fadd_rn(z4.y, df.y), __fadd_rn(z4.z, df.z), __fadd_rn(z4.w, df.w));
      err = fmaf(df.x, df.x, err); err = fmaf(df.y, df.y, err); err = fmaf(df.z, df.z, err); err = fmaf(df.w, df.w, err);
    }
  }
  if (sqerr_sum) {
    const double e = warp_sum(static_cast<double>(err));
    if (lane == 0 && e != 0.0) atomicAdd(sqerr_sum, e);
  }
}

// ------------------------------------------------------------------------------------ host side
static size_t fz_align(size_t v) { return (v + 255) / 256 * 256; }

bool fused_supported(int64_t N, int K, int D) {
  const char* f = std::getenv("VQB200_FORCE_SIMT");
  if (f && f[0] == '1') return false;
  const char* g = std::getenv("VQB200_NO_FUSED");
  if (g && g[0] == '1') return false;
  const char* h = std::getenv("VQB200_FUSED_D128");        // D = 128 (BM = 128, one CTA per SM): opt-out switch
  const bool d128 = D == 128 && !(h && h[0] == '0');
  return (D == 64 || d128) && K >= TC_BN && K < (1 << 24) && N >= 4096 && N <= 0x7fffffff;
}

size_t fused_workspace_bytes(int64_t N) { return 256 + fz_align(static_cast<size_t>(N) * 4) + fz_align(static_cast<size_t>(N) * 8); }

static int fused_smem_bytes(int D, int stages, int BM) {
  const int nepi = fz_nepi(BM);
  return 1024 + BM * D * 4 + BM * D * 2 + stages * TC_STAGE_BYTES + 2 * BM * 4 + nepi * TC_BN * 4 +
         nepi * FZ_PAIRS * 4 + nepi * FZ_PAIRS * 8 + fz_hist(BM) * 4 + 16 * 8 + nepi * FZ_RING_BYTES + 2 * BM * 4 + 256;
}

// Tile height: 256 rows, one CTA per SM (default), or 128 rows with two co-resident CTAs (VQB200_FUSED_BM=128;
// measured slower: the same number of epilogue warps per SM, twice the fixed per-tile overheads).
static int fused_bm() {
  const char* e = std::getenv("VQB200_FUSED_BM");
  return (e && e[0] == '1') ? 128 : 256;
}

// D = 64: BM = 256, one CTA per SM (or BM = 128 with two co-resident CTAs, a measurement switch).
// D = 128: BM = 128, one CTA per SM (the fp32 staging tile of 256 rows would not fit).
template <int D, int BM>
static int launch_fused_cfg(const CUtensorMap& map_zf, const CUtensorMap& map_e, FusedParams& p, bool bf,
                            cudaStream_t s) {
  constexpr bool kPair = D == 64 && BM == 128;            // two CTAs of BM = 128 share one SM's 228 KB (1 KB reserved each)
  const int limit = kPair ? (233472 - 2 * 1024) / 2 : TC_SMEM_LIMIT;
  int stages = 8;
  while (stages > 2 && fused_smem_bytes(D, stages, BM) > limit) --stages;
  if (fused_smem_bytes(D, stages, BM) > limit) return VQB200_ESHAPE;
  const char* dbg = std::getenv("VQB200_DEBUG");
  static bool attr_done_dev[64] = {};
  bool& attr_done = attr_done_dev[current_device_slot()];
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(quantize_fused_kernel<D, false, BM>, cudaFuncAttributeMaxDynamicSharedMemorySize, limit);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(quantize_fused_kernel<D, true, BM>, cudaFuncAttributeMaxDynamicSharedMemorySize, limit);
    if (e != cudaSuccess) return status_of(e);
    // co-residency needs the largest shared-memory carve-out the SM offers
    cudaFuncSetAttribute(quantize_fused_kernel<D, false, BM>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(quantize_fused_kernel<D, true, BM>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (dbg && dbg[0] == '1') {                             // NB: the occupancy API reports 1 for any tcgen05.alloc kernel
      int nb = -1;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, quantize_fused_kernel<D, false, BM>, fz_threads(D, BM),
                                                    fused_smem_bytes(D, stages, BM));
      fprintf(stderr, "[vqb200] fused D=%d BM=%d stages=%d smem=%d B: occupancy API says %d CTA(s) per SM\n", D, BM,
              stages, fused_smem_bytes(D, stages, BM), nb);
    }
    attr_done = true;
  }
  p.stages = stages;
  p.row_tiles = static_cast<int>((p.n_rows + BM - 1) / BM);
  int slots = kNumSMs * (kPair ? 2 : 1);
  const char* genv = std::getenv("VQB200_FUSED_GRID");          // measurement switch
  if (genv && std::atoi(genv) > 0) slots = std::atoi(genv);
  const int grid = p.row_tiles < slots ? p.row_tiles : slots;
  const int smem = fused_smem_bytes(D, stages, BM);
  const bool trace = dbg && dbg[0] == '2';
  p.dbg = nullptr;
  if (trace) cudaMallocManaged(&p.dbg, static_cast<size_t>(grid) * 3 * sizeof(long long));
  timing_mark_begin(s);
  if (bf) quantize_fused_kernel<D, true, BM><<<grid, fz_threads(D, BM), smem, s>>>(map_zf, map_e, p);
  else quantize_fused_kernel<D, false, BM><<<grid, fz_threads(D, BM), smem, s>>>(map_zf, map_e, p);
  timing_mark_end(s);
  if (trace) {                                                   // CTAs that overlapped in time on one SM
    cudaStreamSynchronize(s);
    int overlap = 0;
    long long lo = p.dbg[1], hi = p.dbg[2];
    for (int i = 0; i < grid; ++i) {
      if (p.dbg[i * 3 + 1] < lo) lo = p.dbg[i * 3 + 1];
      if (p.dbg[i * 3 + 2] > hi) hi = p.dbg[i * 3 + 2];
      for (int j = i + 1; j < grid; ++j)
        if (p.dbg[i * 3] == p.dbg[j * 3] && p.dbg[i * 3 + 1] < p.dbg[j * 3 + 2] && p.dbg[j * 3 + 1] < p.dbg[i * 3 + 2]) ++overlap;
    }
    fprintf(stderr, "[vqb200] fused D=%d BM=%d grid=%d: %d co-resident CTA pairs, span %.1f us\n", D, BM, grid, overlap,
            (hi - lo) * 1e-3);
    cudaFree(p.dbg);
  }
  return status_of(cudaGetLastError());
}

int launch_quantize_fused(const float* z, int64_t N, int D, const float* E, const uint16_t* E_bf16, const float* ee_half,
                          const float* ee_half_bf16, const float* level_meta, int K, int mode, int64_t idx_offset,
                          int64_t* idx_out, float* zq_out, float* zq_st_out, double* sqerr_sum, int32_t* hist,
                          const uint8_t* row_mask, void* workspace, size_t workspace_bytes, cudaStream_t s) {
  if (!fused_supported(N, K, D)) return VQB200_ESHAPE;
  if (workspace_bytes < fused_workspace_bytes(N)) return VQB200_EWORKSPACE;
  const bool bf = mode == VQB200_MODE_BF16_INPUT;
  const int BM = D == 128 ? 128 : fused_bm();

  uint8_t* w = static_cast<uint8_t*>(workspace);
  int* counters = reinterpret_cast<int*>(w); w += 256;
  int* fb_rows = reinterpret_cast<int*>(w); w += fz_align(static_cast<size_t>(N) * 4);
  uint64_t* fb_packed = reinterpret_cast<uint64_t*>(w);

  CUtensorMap map_zf, map_e;
  if (!make_tensor_map_2d(&map_zf, z, N, D, BM, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4)) return VQB200_EDRIVER;
  if (!make_tensor_map_2d(&map_e, E_bf16, K, D, TC_BN,
                          bf ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2))
    return VQB200_EDRIVER;

  cudaError_t e = cudaMemsetAsync(counters, 0, 2 * sizeof(int), s);
  if (e != cudaSuccess) return status_of(e);

  FusedParams p;
  p.n_rows = N; p.K = K;
  p.row_tiles = 0;
  p.code_tiles = (K + TC_BN - 1) / TC_BN;
  p.stages = 0; p.mode = mode;
  p.z = z; p.E = E; p.Eb = reinterpret_cast<const __nv_bfloat16*>(E_bf16);
  p.ee_half = bf ? ee_half_bf16 : ee_half; p.level_meta = level_meta;
  p.idx_offset = idx_offset; p.idx_out = idx_out; p.zq_out = zq_out; p.zq_st_out = zq_st_out;
  p.sqerr_sum = sqerr_sum; p.hist = hist; p.row_mask = row_mask;
  p.fb_rows = fb_rows; p.fb_packed = fb_packed; p.counters = counters;
  const int ls = D == 128 ? launch_fused_cfg<128, 128>(map_zf, map_e, p, bf, s)
                          : BM == 256 ? launch_fused_cfg<64, 256>(map_zf, map_e, p, bf, s)
                                      : launch_fused_cfg<64, 128>(map_zf, map_e, p, bf, s);
  if (ls != VQB200_OK) return ls;

  const int st = launch_search_simt_list(z, fb_rows, counters, N, D, E, bf ? ee_half_bf16 : ee_half, K, bf ? 1 : 0,
                                         idx_offset, fb_packed, s);
  if (st != VQB200_OK) return st;
  fused_fixup_kernel<<<64, 256, 0, s>>>(fb_rows, fb_packed, counters, z, E, D, idx_out, zq_out, zq_st_out, sqerr_sum, hist,
                                        row_mask);
  return status_of(cudaGetLastError());
}

}  // namespace vqb
