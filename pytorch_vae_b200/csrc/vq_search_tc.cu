// Tensor-core nearest-code search for sm_100a: tcgen05.mma (bf16 x bf16 -> fp32 in TMEM), operands
// staged by TMA, and a fused per-row candidate/argmax epilogue read straight out of TMEM, so the
// N x K score matrix never reaches shared memory, L2 or HBM.
//
// Exactness (BASELINE.json north_star: reference-exact codes from fp32 inputs).  A single bf16 pass
// cannot decide near-ties, so the epilogue does not pick a winner; it keeps, per row, every code
// whose approximate score  s~ = z~.e~ - |e|^2/2  lies within a RIGOROUS error bound of the running
// maximum:
//     |z~.e~ - z.e| <= |z - z~| max|e~| + |z| max|e - e~|      [+ fp32 accumulation slack]
// (the ACTUAL rounding-error norms of the row and of the codebook, see admission_margin_fp32 in common.cuh)
// so the true argmax a satisfies  s~_a >= max s~ - margin,  margin = 2 x that bound x 1.04.
// A tiny re-rank kernel then evaluates the few survivors exactly (fp64 accumulation of the fp32
// inputs) and takes the lowest index among exact ties.  Rows whose list overflows, rows with
// non-finite values and non-finite codebooks are handed to the exact SIMT kernel.  In bf16-input
// mode the products are exact and only the accumulation order differs; the same machinery runs
// with a much smaller margin and re-ranks on the bf16-rounded values.
//
// Kernel anatomy (one CTA per SM, persistent over (row tile, code split) work items):
//   warp 0   : TMA producer  - z tile [BM x D] once per item, codebook blocks [128 x 64] in a ring
//   warp 1   : MMA issuer    - one thread issues tcgen05.mma M=128 N=128 K=16, accumulators in TMEM
//                              (2 buffers x BM/128 halves x 128 columns = all 512 columns at BM=256)
//   warps 2+ : epilogue      - tcgen05.ld 32 columns at a time, subtract |e|^2/2, running max and
//                              candidate append; overlaps the MMAs of the next code tile
#include <cstdio>
#include <cstdlib>

#include "tc_ptx.cuh"

namespace vqb {

constexpr int TC_CAND = 31;         // candidate records per (row, code split); slot 31 is write scratch
constexpr int TC_SLOTS = TC_CAND + 1;

// ------------------------------------------------------------------------------------ z pre-pass
// z fp32 -> 16-bit tensor-core operand (fp16 in fp32 mode, bf16 in bf16_input mode; RN) plus the per-row
// admission margin.  One warp per row.
// With E_sub / idx_sub / residual_out the row is first replaced by the residual fl(z - E_sub[idx_sub[row]])
// (models/vq_vae.py:258), which is also written out in fp32: the residual update of one RVQ level and the
// pre-pass of the next in ONE read of the row.
__global__ void __launch_bounds__(256, 3)
zprep_kernel(const float* __restrict__ z, int64_t n, int D, int mode, const float* __restrict__ level_meta,
             __nv_bfloat16* __restrict__ zb, float* __restrict__ margin, const float* __restrict__ E_sub,
             const int64_t* __restrict__ idx_sub, int K_sub, float* __restrict__ residual_out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const int D4 = D >> 2;
  const bool bfm = mode == VQB200_MODE_BF16_INPUT;
  const float emax = bfm ? level_meta[2] : level_meta[0];
  const float emax_lp = level_meta[4], rho_e = level_meta[5];
  const bool code_bad = level_meta[1] != 0.f;
  // fp32 mode: admission_margin_fp32 (common.cuh)
  // bf16 mode: inputs are exact, only accumulation order differs: 2 (D + 32) 2^-23
  const float coef = 2.f * static_cast<float>(D + 32) * 1.1920929e-7f;
  for (int64_t row = warp; row < n; row += nwarps) {
    float ss = 0.f, sse = 0.f;
    int64_t ksub = -1;
    if (E_sub) {
      ksub = idx_sub[row];
      if (ksub < 0 || ksub >= K_sub) ksub = -1;           // never produced by vqb200_search; leaves the row as is
    }
    // every load of the row (and of the code row being subtracted) is issued before the first conversion: at the
    // stage-2 shape (8192 rows x 512 floats, one launch per level) the kernel is latency-, not bandwidth-bound
    constexpr int ZP_SL = 4;                                 // float4 slices per lane held in flight: D <= 512
    for (int c0 = lane; c0 < D4; c0 += 32 * ZP_SL) {
      float4 vv[ZP_SL], ee4[ZP_SL];
#pragma unroll
      for (int u = 0; u < ZP_SL; ++u) {
        const int c = c0 + u * 32;
        if (c < D4) {
          vv[u] = ld_stream(reinterpret_cast<const float4*>(z) + row * D4 + c);
          if (E_sub && ksub >= 0) ee4[u] = __ldg(reinterpret_cast<const float4*>(E_sub) + ksub * D4 + c);
        }
      }
#pragma unroll
      for (int u = 0; u < ZP_SL; ++u) {
        const int c = c0 + u * 32;
        if (c >= D4) continue;
        float4 v = vv[u];
        if (E_sub) {
          if (ksub >= 0) {
            const float4 e = ee4[u];
            v.x = __fsub_rn(v.x, e.x); v.y = __fsub_rn(v.y, e.y); v.z = __fsub_rn(v.z, e.z); v.w = __fsub_rn(v.w, e.w);
          }
          st_stream(reinterpret_cast<float4*>(residual_out) + row * D4 + c, v);
        }
        uint16_t b0, b1, b2, b3;
        float f0, f1, f2, f3;
        if (bfm) {
          b0 = __bfloat16_as_ushort(__float2bfloat16_rn(v.x)); b1 = __bfloat16_as_ushort(__float2bfloat16_rn(v.y));
          b2 = __bfloat16_as_ushort(__float2bfloat16_rn(v.z)); b3 = __bfloat16_as_ushort(__float2bfloat16_rn(v.w));
          f0 = __uint_as_float(static_cast<uint32_t>(b0) << 16); f1 = __uint_as_float(static_cast<uint32_t>(b1) << 16);
          f2 = __uint_as_float(static_cast<uint32_t>(b2) << 16); f3 = __uint_as_float(static_cast<uint32_t>(b3) << 16);
        } else {
          const uint32_t q0 = f16x2_bits_flush(v.x, v.y), q1 = f16x2_bits_flush(v.z, v.w);
          b0 = static_cast<uint16_t>(q0); b1 = static_cast<uint16_t>(q0 >> 16);
          b2 = static_cast<uint16_t>(q1); b3 = static_cast<uint16_t>(q1 >> 16);
          const float2 g0 = f16x2_bits_to_float2(q0), g1 = f16x2_bits_to_float2(q1);
          f0 = g0.x; f1 = g0.y; f2 = g1.x; f3 = g1.y;
        }
        uint2 pk;
        pk.x = static_cast<uint32_t>(b0) | (static_cast<uint32_t>(b1) << 16);
        pk.y = static_cast<uint32_t>(b2) | (static_cast<uint32_t>(b3) << 16);
        reinterpret_cast<uint2*>(zb)[row * D4 + c] = pk;
        if (bfm) {
          ss += f0 * f0 + f1 * f1 + f2 * f2 + f3 * f3;
        } else {
          ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
          const float d0 = v.x - f0, d1 = v.y - f1, d2 = v.z - f2, d3 = v.w - f3;      // exact differences
          sse += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
        }
      }
    }
    ss = warp_sum(ss);
    sse = warp_sum(sse);
    if (lane == 0) {
      float m = bfm ? coef * (sqrtf(ss) * 1.0001f) * emax + 1e-30f : admission_margin_fp32(ss, sse, emax, emax_lp, rho_e, D);
      if (code_bad || !(ss < __int_as_float(0x7f800000)) || !(sse < __int_as_float(0x7f800000)))
        m = __int_as_float(0x7fc00000);   // NaN: exact path
      margin[row] = m;
    }
  }
}

int launch_residual_prep(const float* z, const float* E_full, const int64_t* idx, int64_t N, int D, int K_total,
                         int mode, const float* next_level_meta, float* residual_out, uint16_t* z16_out,
                         float* margin_out, cudaStream_t s) {
  if (N == 0) return VQB200_OK;
  int64_t blocks = (N + 7) / 8;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  zprep_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(z, N, D, mode, next_level_meta,
                                                             reinterpret_cast<__nv_bfloat16*>(z16_out), margin_out, E_full,
                                                             idx, K_total, residual_out);
  return status_of(cudaGetLastError());
}

// ------------------------------------------------------------------------------------ the tensor kernel
struct TcParams {
  int64_t n_rows;        // rows in this chunk
  int D, K;
  int row_tiles, ksplit, tiles_per_split, code_tiles;
  int stages;
  int zbufs;             // 1 or 2 z-tile buffers (2 when shared memory allows: next item's rows prefetch)
  uint32_t idesc;        // UMMA instruction descriptor: bf16 or fp16 operands (by mode), this kernel's M x N
  const float* ee_half;  // [K] (plane chosen by mode)
  const float* margin;   // [n_rows]
  uint2* cand;           // [n_rows][ksplit][TC_SLOTS] records {code group << 8 | admit mask, group max}
  int* cnt;              // [n_rows][ksplit]
  float* best;           // [n_rows][ksplit]
  long long* clk;        // VQB200_DEBUG=4: per CTA [clock64, globaltimer] at start and at end (effective SM clock)
};

// Scan 32 accumulator columns of one row.  The accumulator was pre-loaded with -|e|^2/2, so the TMEM words ARE
// the scores s = z.e - |e|^2/2: no shared-memory read and no subtract per column.  The whole chunk is skipped
// with ONE compare unless its maximum reaches the admission threshold thr = best - margin (rare once the
// running maximum has settled: ~ (1 + margin * best) / j per column j).  In the slow path each group of 8
// columns whose maximum passes emits one record {group index << 8 | 8-bit admit mask, group max}.
__device__ __forceinline__ void epi_scan(const uint32_t (&v)[32], uint32_t code0, float margin, float& best,
                                         float& thr, int& cnt, uint2* cand_row) {
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
      if (gm[g] >= thr) {                              // ~1 of the 4 groups: the mask is built only where needed
        uint32_t mk = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) mk |= (__uint_as_float(v[g * 8 + i]) >= thr) ? (1u << i) : 0u;
        if (cnt < TC_CAND) cand_row[cnt] = make_uint2((((code0 >> 3) + g) << 8) | mk, __float_as_uint(gm[g]));
        ++cnt;
      }
    }
  }
}

// Epilogue column split CS: CS warps share a row quarter and scan TC_BN/CS columns each.  BM=256 runs 8
// epilogue warps (CS=1: a 10-warp CTA keeps the full register budget, no spills); BM=128 (D=512) runs CS=2.
constexpr int tc_cs(int BM) { return BM == 128 ? 2 : 1; }

template <int BM>
__global__ void __launch_bounds__(64 + BM * tc_cs(BM), 1)
search_tc_kernel(const __grid_constant__ CUtensorMap tmap_z, const __grid_constant__ CUtensorMap tmap_e,
                 const TcParams p) {
  constexpr int NHALF = BM / 128;
  constexpr int TC_CS = tc_cs(BM);
  constexpr int NEPI = (BM / 32) * TC_CS;
  constexpr int WCOLS = TC_BN / TC_CS;           // columns one epilogue warp scans per tile
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int KBLK = p.D / TC_KB;
  const uint32_t z_bytes = static_cast<uint32_t>(BM) * p.D * 2;
  const uint32_t z_smem = base;                                    // zbufs x KBLK slabs of [BM rows x 128 B]
  const uint32_t e_smem = z_smem + p.zbufs * z_bytes;              // ring of [128 codes x 128 B]
  const uint32_t ee_smem = e_smem + static_cast<uint32_t>(p.stages) * TC_STAGE_BYTES;   // [NEPI][2][WCOLS] fp32
  const uint32_t bar0 = ee_smem + NEPI * 2 * WCOLS * 4;
  // barrier map (8 bytes each)
  const uint32_t bar_full = bar0;                         // [stages]
  const uint32_t bar_empty = bar0 + 8 * 8;                // [stages]
  const uint32_t bar_tfull = bar0 + 16 * 8;               // [2]
  const uint32_t bar_tempty = bar0 + 18 * 8;              // [2]
  const uint32_t bar_zfull = bar0 + 20 * 8;               // [2]
  const uint32_t bar_zempty = bar0 + 22 * 8;              // [2]
  const uint32_t tmem_slot = bar0 + 24 * 8;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));  // generic pointer to the aligned base

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(bar_tfull + 8 * b, 1); mbar_init(bar_tempty + 8 * b, NEPI); }
    for (int b = 0; b < 2; ++b) { mbar_init(bar_zfull + 8 * b, 1); mbar_init(bar_zempty + 8 * b, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

  const int n_items = p.row_tiles * p.ksplit;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    // The whole warp walks the loop (warp-uniform control flow keeps addresses in uniform registers);
    // one elected lane issues the copies.
    uint32_t stage = 0, phase = 0, it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int rt = item / p.ksplit, ks = item - rt * p.ksplit;
      const int t0 = ks * p.tiles_per_split;
      const int t1 = min(t0 + p.tiles_per_split, p.code_tiles);
      const uint32_t zb = p.zbufs == 2 ? (it & 1) : 0;
      const uint32_t zuse = p.zbufs == 2 ? (it >> 1) : it;
      mbar_wait(bar_zempty + 8 * zb, (zuse & 1) ^ 1);
      if (elect_one()) {
        mbar_expect_tx(bar_zfull + 8 * zb, z_bytes);
        for (int kb = 0; kb < KBLK; ++kb)
          tma_load_2d(z_smem + zb * z_bytes + kb * (BM * 128), &tmap_z, bar_zfull + 8 * zb, kb * TC_KB, rt * BM);
      }
      __syncwarp();
      for (int t = t0; t < t1; ++t) {
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
    // Warp-uniform loop, one elected lane issues.  Descriptors advance by adds on the low word only.
    uint32_t stage = 0, phase = 0, it = 0, tg = 0;
    const uint32_t z_lo = umma_desc_lo(z_smem), e_lo = umma_desc_lo(e_smem);
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int rt = item / p.ksplit, ks = item - rt * p.ksplit;
      const int t0 = ks * p.tiles_per_split;
      const int t1 = min(t0 + p.tiles_per_split, p.code_tiles);
      const uint32_t zb = p.zbufs == 2 ? (it & 1) : 0;
      const uint32_t zuse = p.zbufs == 2 ? (it >> 1) : it;
      (void)rt;
      mbar_wait(bar_zfull + 8 * zb, zuse & 1);
      for (int t = t0; t < t1; ++t, ++tg) {
        const uint32_t b = tg & 1;
        mbar_wait(bar_tempty + 8 * b, (tg >> 1) & 1);   // buffer drained AND pre-loaded with -|e|^2/2
        tc_fence_after();
        for (int kb = 0; kb < KBLK; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a0 = z_lo + ((zb * z_bytes + kb * (BM * 128)) >> 4);
            const uint32_t b0 = e_lo + ((stage * TC_STAGE_BYTES) >> 4);
#pragma unroll
            for (int h = 0; h < NHALF; ++h) {
              const uint32_t d_tmem = tmem_base + b * (NHALF * TC_BN) + h * TC_BN;
#pragma unroll
              for (int k = 0; k < TC_KB / 16; ++k)
                tc_mma_bf16(d_tmem, umma_desc(a0 + h * ((128 * 128) >> 4) + k * 2), umma_desc(b0 + k * 2), p.idesc, 1u);
            }
            tc_commit(bar_empty + 8 * stage);          // frees the smem stage when these MMAs retire
            if (kb == KBLK - 1) tc_commit(bar_tfull + 8 * b);   // accumulator tile complete
          }
          __syncwarp();
          if (++stage == static_cast<uint32_t>(p.stages)) { stage = 0; phase ^= 1; }
        }
      }
      if (elect_one()) tc_commit(bar_zempty + 8 * zb);  // z tile may be overwritten
      __syncwarp();
    }
  } else {
    // ============================== epilogue ==============================
    const int we = warp - 2;                           // 0 .. NEPI-1
    const int quarter = warp & 3;                      // TMEM lane quarter this warp may touch
    const int half = (we >> 2) % NHALF;                // which 128-row accumulator
    const int cs = (we >> 2) / NHALF;                  // which column slice of the tile
    float* ee_slot = reinterpret_cast<float*>(gen + (ee_smem - base)) + we * 2 * WCOLS;   // [WCOLS] staging, per warp
    const float kNegInf = __int_as_float(0xff800000);
    const uint32_t tcol = (static_cast<uint32_t>(quarter * 32) << 16) + half * TC_BN + cs * WCOLS;

    // -|e|^2/2 of this warp's column slice of code tile t: WCOLS/32 values per lane (padding codes get -inf)
    constexpr int BPL = WCOLS / 32;
    struct Bias { float v[BPL]; };
    auto load_bias = [&](int t) -> Bias {
      Bias r;
#pragma unroll
      for (int j = 0; j < BPL; ++j) {
        const int c = t * TC_BN + cs * WCOLS + lane * BPL + j;
        r.v[j] = (t >= 0 && c < p.K) ? -p.ee_half[c] : kNegInf;
      }
      return r;
    };
    // write the bias of one code tile into this warp's lanes/columns of accumulator buffer b
    auto preload = [&](const Bias& bias, uint32_t b) {
      __syncwarp();
#pragma unroll
      for (int j = 0; j < BPL; ++j) ee_slot[lane * BPL + j] = bias.v[j];
      __syncwarp();
#pragma unroll
      for (int hh = 0; hh < WCOLS / 16; ++hh) {
        uint32_t w[16];
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 q4 = *reinterpret_cast<const float4*>(ee_slot + hh * 16 + j);   // broadcast read
          w[j + 0] = __float_as_uint(q4.x); w[j + 1] = __float_as_uint(q4.y);
          w[j + 2] = __float_as_uint(q4.z); w[j + 3] = __float_as_uint(q4.w);
        }
        TC_ST16(tmem_base + tcol + b * (NHALF * TC_BN) + hh * 16, w);
      }
      tc_wait_st();
    };
    // walk the CTA's tile sequence two steps ahead of the tile being scanned (the next user of a buffer)
    int la_item = blockIdx.x, la_t = 0, la_t1 = 0;
    auto la_open = [&]() {
      if (la_item < n_items) {
        const int ks = la_item % p.ksplit;
        la_t = ks * p.tiles_per_split;
        la_t1 = min(la_t + p.tiles_per_split, p.code_tiles);
      }
    };
    auto la_next = [&]() -> int {                       // code tile index of the next tile in sequence, -1 at the end
      while (la_item < n_items && la_t >= la_t1) { la_item += gridDim.x; la_open(); }
      if (la_item >= n_items) return -1;
      return la_t++;
    };
    la_open();
    // prologue: both buffers receive the bias of the first two tiles, then are handed to the MMA warp
    for (uint32_t b = 0; b < 2; ++b) {
      const int tt = la_next();
      if (tt >= 0) preload(load_bias(tt), b);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * b);
    }
    int t_ahead = la_next();                            // tile that will reuse the buffer of the first scanned tile
    Bias bias_next = load_bias(t_ahead);

    uint32_t tg = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int rt = item / p.ksplit, ks = item - rt * p.ksplit;
      const int t0 = ks * p.tiles_per_split;
      const int t1 = min(t0 + p.tiles_per_split, p.code_tiles);
      const int64_t row = static_cast<int64_t>(rt) * BM + half * 128 + quarter * 32 + lane;
      const bool row_ok = row < p.n_rows;
      const float margin = row_ok ? p.margin[row] : __int_as_float(0x7fc00000);
      const int64_t sub = row_ok ? (row * p.ksplit + ks) * TC_CS + cs : 0;   // this warp's candidate sub-list
      uint2* cand_row = p.cand + sub * TC_SLOTS;
      float best = kNegInf;
      // admission threshold best - margin (NaN: never admits).  It starts at the lowest FINITE value: a chunk
      // of -inf scores (padding, de-duplicated dead codes) must not be admitted, or a slice made of such
      // columns only would overflow its record list with them.
      float thr = margin == margin ? -3.0e38f : margin;
      int cnt = 0;

      for (int t = t0; t < t1; ++t, ++tg) {
        const uint32_t b = tg & 1;
        const Bias bias = bias_next;                    // for tile t_ahead (fetched one iteration ago)
        const int t_cur_ahead = t_ahead;
        t_ahead = la_next();
        bias_next = load_bias(t_ahead);                 // global loads in flight across the wait below
        mbar_wait(bar_tfull + 8 * b, (tg >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + tcol + b * (NHALF * TC_BN);
        // One 32-column chunk in registers at a time (the 18-warp CTA is capped at 96 registers per thread).
        // After the LAST chunk is out of TMEM the buffer is re-armed with the bias of its next tile and
        // handed back to the MMA warp; only then is that chunk scanned.
        uint32_t v[32];
        const uint32_t code_t = static_cast<uint32_t>(t * TC_BN + cs * WCOLS);
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
          epi_scan(v, code_t + ch * 32, margin, best, thr, cnt, cand_row);
        }
      }
      if (row_ok) {
        p.cnt[sub] = cnt;
        p.best[sub] = best;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------ CTA-pair variant
// Same search, issued as tcgen05.mma.cta_group::2: a cluster of two CTAs (one SM pair) computes a
// [256 rows x 256 codes] tile per MMA sweep; each CTA stages ITS 128 rows of z and ITS 128 codes of the tile,
// and receives its 128 rows x 256 columns of accumulator in its own TMEM.  Per CTA a K=16 step moves
// 4 KB (A) + 4 KB (B) of shared-memory operands per 128 cycles instead of 8 KB per 64 cycles in the 1-CTA
// kernel, whose M=128 x N=128 MMAs sit exactly at the 128 B/clk shared-memory limit and so run at ~60 %.
// Protocol: the leader CTA (rank 0) issues every MMA.  Both CTAs' TMA loads credit the LEADER's full / z-full
// barriers; MMA completions are committed with a multicast arrive to the barriers at the same offsets in both
// CTAs; both CTAs' epilogue warps arrive (remotely for the peer) on the leader's tmem-empty barriers.
// ---- side jobs: HBM-bound streaming work of the NEIGHBOURING chunks, run by extra warps of the persistent tensor
// kernel.  The tensor kernel fills every SM (one CTA, ~218 KB of shared memory), so no other kernel co-resides with
// it and the chunk pipeline's side passes -- the pre-pass of the next chunk (fp32 -> 16-bit operand + margins,
// 6 D bytes per row) and the gather / straight-through / loss / histogram pass of the previous chunk (12 D bytes per
// row), both at the HBM roofline when run alone -- used to sit BETWEEN the tensor kernels: 4.5 ms of a 16.4 ms step
// at K = 8192, D = 256, N = 2^22.  Six extra warps per CTA stream that work while the tensor pipe is busy: it needs
// ~1.5 TB/s of the 6.5 TB/s HBM and ~2 % of the issue slots, and touches neither the TMA / MMA / epilogue protocol
// nor their registers (setmaxnreg gives the epilogue warp groups 168 registers, everything else 88).
struct SideJobs {
  int D, mode;
  // gather of an EARLIER chunk (its indices are final: re-rank and hand-back ran before this launch)
  const float* g_z; const int64_t* g_idx; int64_t g_rows;
  const float* g_E; int g_K_total;
  float* g_zq; float* g_zq_st; double* g_sqerr; int32_t* g_hist; const uint8_t* g_mask;
  // pre-pass of a LATER chunk
  const float* p_z; int64_t p_rows; __nv_bfloat16* p_zb; float* p_margin; const float* level_meta;
};

// One warp per row, SL float4 slices per lane (D = 128 SL), RB rows in flight.
template <int SL>
__device__ __forceinline__ void side_gather(const SideJobs& sj, int64_t w0, int64_t nw, int lane) {
  constexpr int RB = SL == 1 ? 6 : (SL == 2 ? 3 : (SL == 3 ? 2 : 1));   // 48 data registers in flight (32 at D = 512)
  const int D4 = SL * 32;
  const float4* Z = reinterpret_cast<const float4*>(sj.g_z);
  const float4* E = reinterpret_cast<const float4*>(sj.g_E);
  float4* ZQ = reinterpret_cast<float4*>(sj.g_zq);
  float4* ST = reinterpret_cast<float4*>(sj.g_zq_st);
  float err = 0.f;
  for (int64_t r0 = w0 * RB; r0 < sj.g_rows; r0 += nw * RB) {
    int k[RB];
    float4 v[RB][SL], e[RB][SL];
#pragma unroll
    for (int u = 0; u < RB; ++u) {
      const int64_t row = r0 + u;
      k[u] = -1;
      if (row < sj.g_rows) {
        const int64_t kk = sj.g_idx[row];
        k[u] = (kk >= 0 && kk < sj.g_K_total) ? static_cast<int>(kk) : -1;
#pragma unroll
        for (int s = 0; s < SL; ++s) v[u][s] = ld_stream(Z + row * D4 + s * 32 + lane);
      }
    }
#pragma unroll
    for (int u = 0; u < RB; ++u)
      if (k[u] >= 0) {
#pragma unroll
        for (int s = 0; s < SL; ++s) e[u][s] = __ldg(E + static_cast<int64_t>(k[u]) * D4 + s * 32 + lane);
      }
#pragma unroll
    for (int u = 0; u < RB; ++u) {
      if (k[u] < 0) continue;
      const int64_t row = r0 + u;
#pragma unroll
      for (int s = 0; s < SL; ++s) {
        const float4 q = e[u][s], zz = v[u][s];
        float4 df;
        df.x = __fsub_rn(q.x, zz.x); df.y = __fsub_rn(q.y, zz.y); df.z = __fsub_rn(q.z, zz.z); df.w = __fsub_rn(q.w, zz.w);
        if (ZQ) st_stream(ZQ + row * D4 + s * 32 + lane, q);
        if (ST) st_stream(ST + row * D4 + s * 32 + lane,
                          make_float4(__fadd_rn(zz.x, df.x), __fadd_rn(zz.y, df.y), __fadd_rn(zz.z, df.z), __fadd_rn(zz.w, df.w)));
        err = fmaf(df.x, df.x, err); err = fmaf(df.y, df.y, err); err = fmaf(df.z, df.z, err); err = fmaf(df.w, df.w, err);
      }
      if (sj.g_hist && lane == 0 && (!sj.g_mask || sj.g_mask[row])) atomicAdd(sj.g_hist + k[u], 1);
    }
  }
  if (sj.g_sqerr) {
    const double t = warp_sum(static_cast<double>(err));
    if (lane == 0 && t != 0.0) atomicAdd(sj.g_sqerr, t);
  }
}

template <int SL>
__device__ __forceinline__ void side_prep(const SideJobs& sj, int64_t w0, int64_t nw, int lane) {
  constexpr int RB = SL == 1 ? 6 : (SL == 2 ? 3 : (SL == 3 ? 2 : 1));   // 48 data registers in flight (32 at D = 512)
  const int D4 = SL * 32, D = SL * 128;
  const bool bfm = sj.mode == VQB200_MODE_BF16_INPUT;
  const float emax = bfm ? sj.level_meta[2] : sj.level_meta[0];
  const float emax_lp = sj.level_meta[4], rho_e = sj.level_meta[5];
  const bool code_bad = sj.level_meta[1] != 0.f;
  const float coef = 2.f * static_cast<float>(D + 32) * 1.1920929e-7f;
  const float4* Z = reinterpret_cast<const float4*>(sj.p_z);
  uint2* ZB = reinterpret_cast<uint2*>(sj.p_zb);
  for (int64_t r0 = w0 * RB; r0 < sj.p_rows; r0 += nw * RB) {
    float4 v[RB][SL];
#pragma unroll
    for (int u = 0; u < RB; ++u)
      if (r0 + u < sj.p_rows) {
#pragma unroll
        for (int s = 0; s < SL; ++s) v[u][s] = ld_stream(Z + (r0 + u) * D4 + s * 32 + lane);
      }
#pragma unroll
    for (int u = 0; u < RB; ++u) {
      const int64_t row = r0 + u;
      if (row >= sj.p_rows) continue;                       // warp-uniform
      float ss = 0.f, sse = 0.f;
#pragma unroll
      for (int s = 0; s < SL; ++s) {
        const float4 x = v[u][s];
        uint2 pk;
        float f0, f1, f2, f3;
        if (bfm) {
          const __nv_bfloat162 a = __floats2bfloat162_rn(x.x, x.y), b = __floats2bfloat162_rn(x.z, x.w);
          pk.x = *reinterpret_cast<const uint32_t*>(&a); pk.y = *reinterpret_cast<const uint32_t*>(&b);
          f0 = __uint_as_float(pk.x << 16); f1 = __uint_as_float(pk.x & 0xffff0000u);
          f2 = __uint_as_float(pk.y << 16); f3 = __uint_as_float(pk.y & 0xffff0000u);
          ss += f0 * f0 + f1 * f1 + f2 * f2 + f3 * f3;
        } else {
          pk.x = f16x2_bits_flush(x.x, x.y); pk.y = f16x2_bits_flush(x.z, x.w);
          const float2 g0 = f16x2_bits_to_float2(pk.x), g1 = f16x2_bits_to_float2(pk.y);
          f0 = g0.x; f1 = g0.y; f2 = g1.x; f3 = g1.y;
          ss += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
          const float d0 = x.x - f0, d1 = x.y - f1, d2 = x.z - f2, d3 = x.w - f3;      // exact differences
          sse += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
        }
        ZB[row * D4 + s * 32 + lane] = pk;
      }
      ss = warp_sum(ss);
      sse = warp_sum(sse);
      if (lane == 0) {
        float m = bfm ? coef * (sqrtf(ss) * 1.0001f) * emax + 1e-30f : admission_margin_fp32(ss, sse, emax, emax_lp, rho_e, D);
        if (code_bad || !(ss < __int_as_float(0x7f800000)) || !(sse < __int_as_float(0x7f800000)))
          m = __int_as_float(0x7fc00000);   // NaN: exact path
        sj.p_margin[row] = m;
      }
    }
  }
}

__device__ __forceinline__ void run_side_jobs(const SideJobs& sj, int side_warp, int n_side, int lane) {
  const int64_t w0 = static_cast<int64_t>(blockIdx.x) * n_side + side_warp;
  const int64_t nw = static_cast<int64_t>(gridDim.x) * n_side;
  const int SL = sj.D >> 7;
  if (sj.g_rows > 0) {
    if (SL == 1) side_gather<1>(sj, w0, nw, lane);
    else if (SL == 2) side_gather<2>(sj, w0, nw, lane);
    else if (SL == 3) side_gather<3>(sj, w0, nw, lane);
    else side_gather<4>(sj, w0, nw, lane);
  }
  if (sj.p_rows > 0) {
    if (SL == 1) side_prep<1>(sj, w0, nw, lane);
    else if (SL == 2) side_prep<2>(sj, w0, nw, lane);
    else if (SL == 3) side_prep<3>(sj, w0, nw, lane);
    else side_prep<4>(sj, w0, nw, lane);
  }
}

constexpr int P2_ROWS = 128;        // rows per CTA
constexpr int P2_BN = 256;          // codes per pair tile (each CTA stages 128 of them)
constexpr int P2_CS = 2;            // column slices: 8 epilogue warps = 4 lane quarters x 2 slices of 128 columns
constexpr int P2_NEPI = 8;
constexpr int P2_WCOLS = P2_BN / P2_CS;

// REGS caps the registers per thread.  At 128 (a few spills on the slow path) the ten warps of a CTA leave
// 4096 registers free in every scheduler partition, which is what lets ONE 256-thread block of the chunk
// pipeline's side kernels (pre-pass, re-rank, gather: <= 64 registers) co-reside with this persistent kernel
// (an experiment switch: see launch_search_tc for the measurement; the default build is uncapped).
// SIDE = 0: ten warps (TMA, MMA, 8 epilogue).  SIDE = 6: sixteen warps in four warp groups -- {TMA, MMA, side, side},
// {epilogue x4}, {epilogue x4}, {side x4} -- launched at 128 registers per thread; the epilogue groups then take 168
// and the others give back down to 88 (setmaxnreg), and the side warps run the SideJobs.
constexpr int P2_SIDE = 6;
template <int SIDE>
__device__ __forceinline__ void search_tc2_body(const CUtensorMap& tmap_z, const CUtensorMap& tmap_e, const TcParams& p,
                                                const SideJobs& sj) {
  constexpr int EPI0 = SIDE > 0 ? 4 : 2;               // first epilogue warp (a warp may touch TMEM lanes 32 (w % 4) ...)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int KBLK = p.D / TC_KB;
  const uint32_t z_bytes = static_cast<uint32_t>(P2_ROWS) * p.D * 2;
  const uint32_t z_smem = base;
  const uint32_t e_smem = z_smem + p.zbufs * z_bytes;
  const uint32_t ee_smem = e_smem + static_cast<uint32_t>(p.stages) * TC_STAGE_BYTES;   // [8][128] fp32 staging
  const uint32_t bar0 = ee_smem + P2_NEPI * P2_WCOLS * 4;
  const uint32_t bar_full = bar0, bar_empty = bar0 + 8 * 8, bar_tfull = bar0 + 16 * 8, bar_tempty = bar0 + 18 * 8;
  const uint32_t bar_zfull = bar0 + 20 * 8, bar_zempty = bar0 + 22 * 8, tmem_slot = bar0 + 24 * 8;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_tfull + 8 * b, 1); mbar_init(bar_tempty + 8 * b, 2 * P2_NEPI);   // both CTAs' epilogue warps
      mbar_init(bar_zfull + 8 * b, 1); mbar_init(bar_zempty + 8 * b, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {      // the same warp of BOTH CTAs allocates for the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                  // peer barriers are initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));
  const int n_items = p.row_tiles;                     // 256-row tiles, one per pair step
  if (p.clk && threadIdx.x == 0) {
    long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
    p.clk[blockIdx.x * 4 + 0] = clock64(); p.clk[blockIdx.x * 4 + 1] = g;
  }
  if (SIDE > 0) {
    if (warp >= EPI0 && warp < EPI0 + P2_NEPI) asm volatile("setmaxnreg.inc.sync.aligned.u32 160;");
    else asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
  }

  if (warp == 0) {
    // ============================== TMA producer (both CTAs) ==============================
    uint32_t stage = 0, phase = 0, it = 0;
    for (int item = pair; item < n_items; item += n_pairs, ++it) {
      const uint32_t zb = p.zbufs == 2 ? (it & 1) : 0;
      const uint32_t zuse = p.zbufs == 2 ? (it >> 1) : it;
      mbar_wait(bar_zempty + 8 * zb, (zuse & 1) ^ 1);
      if (elect_one()) {
        if (leader) mbar_expect_tx(bar_zfull + 8 * zb, 2 * z_bytes);           // its own rows + the peer's
        const uint32_t zf = map_to_cta(bar_zfull + 8 * zb, 0);
        for (int kb = 0; kb < KBLK; ++kb)
          tma_load_2d_pair(z_smem + zb * z_bytes + kb * (P2_ROWS * 128), &tmap_z, zf, kb * TC_KB,
                           item * 2 * P2_ROWS + static_cast<int>(crank) * P2_ROWS);
      }
      __syncwarp();
      for (int t = 0; t < p.code_tiles; ++t) {
        for (int kb = 0; kb < KBLK; ++kb) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          if (elect_one()) {
            if (leader) mbar_expect_tx(bar_full + 8 * stage, 2 * TC_STAGE_BYTES);
            tma_load_2d_pair(e_smem + stage * TC_STAGE_BYTES, &tmap_e, map_to_cta(bar_full + 8 * stage, 0), kb * TC_KB,
                             t * P2_BN + static_cast<int>(crank) * (P2_BN / 2));
          }
          __syncwarp();
          if (++stage == static_cast<uint32_t>(p.stages)) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer (leader CTA only) ==============================
    if (leader) {
      uint32_t stage = 0, phase = 0, it = 0, tg = 0;
      const uint32_t z_lo = umma_desc_lo(z_smem), e_lo = umma_desc_lo(e_smem);
      for (int item = pair; item < n_items; item += n_pairs, ++it) {
        const uint32_t zb = p.zbufs == 2 ? (it & 1) : 0;
        const uint32_t zuse = p.zbufs == 2 ? (it >> 1) : it;
        mbar_wait(bar_zfull + 8 * zb, zuse & 1);
        for (int t = 0; t < p.code_tiles; ++t, ++tg) {
          const uint32_t b = tg & 1;
          mbar_wait(bar_tempty + 8 * b, (tg >> 1) & 1);
          tc_fence_after();
          for (int kb = 0; kb < KBLK; ++kb) {
            mbar_wait(bar_full + 8 * stage, phase);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t a0 = z_lo + ((zb * z_bytes + kb * (P2_ROWS * 128)) >> 4);
              const uint32_t b0 = e_lo + ((stage * TC_STAGE_BYTES) >> 4);
              const uint32_t d_tmem = tmem_base + b * P2_BN;
#pragma unroll
              for (int k = 0; k < TC_KB / 16; ++k)
                tc_mma_bf16_pair(d_tmem, umma_desc(a0 + k * 2), umma_desc(b0 + k * 2), p.idesc, 1u);
              tc_commit_pair(bar_empty + 8 * stage);
              if (kb == KBLK - 1) tc_commit_pair(bar_tfull + 8 * b);
            }
            __syncwarp();
            if (++stage == static_cast<uint32_t>(p.stages)) { stage = 0; phase ^= 1; }
          }
        }
        if (elect_one()) tc_commit_pair(bar_zempty + 8 * zb);
        __syncwarp();
      }
    }
  } else if (SIDE > 0 && (warp < EPI0 || warp >= EPI0 + P2_NEPI)) {
    // ============================== side warps: streaming work of the neighbouring chunks ==============================
    run_side_jobs(sj, warp < EPI0 ? warp - 2 : warp - (EPI0 + P2_NEPI) + 2, SIDE, lane);
  } else {
    // ============================== epilogue (both CTAs, own TMEM) ==============================
    const int we = warp - EPI0;
    const int quarter = warp & 3;
    const int cs = we >> 2;
    float* ee_slot = reinterpret_cast<float*>(gen + (ee_smem - base)) + we * P2_WCOLS;
    const float kNegInf = __int_as_float(0xff800000);
    const uint32_t tcol = (static_cast<uint32_t>(quarter * 32) << 16) + cs * P2_WCOLS;
    const uint32_t tempty0 = map_to_cta(bar_tempty, 0);        // the leader's tmem-empty barriers
    constexpr int BPL = P2_WCOLS / 32;
    struct Bias { float v[BPL]; };
    auto load_bias = [&](int t) -> Bias {
      Bias r;
#pragma unroll
      for (int j = 0; j < BPL; ++j) {
        const int c = t * P2_BN + cs * P2_WCOLS + lane * BPL + j;
        r.v[j] = (t >= 0 && c < p.K) ? -p.ee_half[c] : kNegInf;
      }
      return r;
    };
    auto preload = [&](const Bias& bias, uint32_t b) {
      __syncwarp();
#pragma unroll
      for (int j = 0; j < BPL; ++j) ee_slot[lane * BPL + j] = bias.v[j];
      __syncwarp();
#pragma unroll
      for (int hh = 0; hh < P2_WCOLS / 16; ++hh) {
        uint32_t w[16];
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 q4 = *reinterpret_cast<const float4*>(ee_slot + hh * 16 + j);
          w[j + 0] = __float_as_uint(q4.x); w[j + 1] = __float_as_uint(q4.y);
          w[j + 2] = __float_as_uint(q4.z); w[j + 3] = __float_as_uint(q4.w);
        }
        TC_ST16(tmem_base + tcol + b * P2_BN + hh * 16, w);
      }
      tc_wait_st();
    };
    const int64_t total_tiles = static_cast<int64_t>((n_items - pair + n_pairs - 1) / n_pairs) * p.code_tiles;
    int64_t la = 0;
    int la_t = 0;                                       // code tile of sequence position la (no division per tile)
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
      if (lane == 0) mbar_arrive_cluster(tempty0 + 8 * b);
    }
    int t_ahead = la_next();
    Bias bias_next = load_bias(t_ahead);

    uint32_t tg = 0;
    for (int item = pair; item < n_items; item += n_pairs) {
      const int64_t row = static_cast<int64_t>(item) * 2 * P2_ROWS + crank * P2_ROWS + quarter * 32 + lane;
      const bool row_ok = row < p.n_rows;
      const float margin = row_ok ? p.margin[row] : __int_as_float(0x7fc00000);
      const int64_t sub = row_ok ? row * P2_CS + cs : 0;
      uint2* cand_row = p.cand + sub * TC_SLOTS;
      float best = kNegInf;
      float thr = margin == margin ? -3.0e38f : margin;   // lowest finite value: -inf chunks are never admitted
      int cnt = 0;
      for (int t = 0; t < p.code_tiles; ++t, ++tg) {
        const uint32_t b = tg & 1;
        const Bias bias = bias_next;
        const int t_cur_ahead = t_ahead;
        t_ahead = la_next();
        bias_next = load_bias(t_ahead);
        mbar_wait(bar_tfull + 8 * b, (tg >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + tcol + b * P2_BN;
        uint32_t v[32];
        const uint32_t code_t = static_cast<uint32_t>(t * P2_BN + cs * P2_WCOLS);
#pragma unroll
        for (int ch = 0; ch < P2_WCOLS / 32; ++ch) {
          TC_LD32(taddr + ch * 32, v);
          tc_wait_ld();
          if (ch == P2_WCOLS / 32 - 1) {
            if (t_cur_ahead >= 0) preload(bias, b);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tempty0 + 8 * b);
          }
          epi_scan(v, code_t + ch * 32, margin, best, thr, cnt, cand_row);
        }
      }
      if (row_ok) {
        p.cnt[sub] = cnt;
        p.best[sub] = best;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (p.clk && threadIdx.x == 0) {
    long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
    p.clk[blockIdx.x * 4 + 2] = clock64(); p.clk[blockIdx.x * 4 + 3] = g;
  }
  cluster_sync_all();                                  // neither CTA leaves while the peer may still touch it
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

template <int REGS>
__global__ void __cluster_dims__(2, 1, 1) __maxnreg__(REGS)
search_tc2_kernel(const __grid_constant__ CUtensorMap tmap_z, const __grid_constant__ CUtensorMap tmap_e,
                  const TcParams p) {
  SideJobs none;
  none.g_rows = 0; none.p_rows = 0;
  search_tc2_body<0>(tmap_z, tmap_e, p, none);
}

// the same with six side warps (sixteen warps: 128 registers per thread at launch, re-divided by setmaxnreg)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 256 + 32 * P2_SIDE, 1)
search_tc2_side_kernel(const __grid_constant__ CUtensorMap tmap_z, const __grid_constant__ CUtensorMap tmap_e,
                       const TcParams p, const SideJobs sj) {
  search_tc2_body<P2_SIDE>(tmap_z, tmap_e, p, sj);
}

// ------------------------------------------------------------------------------------ exact re-rank
// One kernel, two granularities.  Each warp takes 32 rows.
//  (1) every THREAD prunes its row's candidate records against the final maximum -- loads only, no
//      stores, so they pipeline.  One surviving code = certified by the error bound, written at once;
//      overflow / nothing admitted / non-finite = queued for the exact SIMT kernel.
//  (2) rows with several survivors are scored by the whole WARP, one row at a time: the surviving codes are
//      expanded into a small shared list, then 32/(D/4) codes are scored concurrently by lane groups
//      (fp64 accumulation of the fp32 -- or bf16-rounded -- inputs); (score, index) is reduced
//      lexicographically so the lowest index wins exact ties.
constexpr int RR_LIST = 64;   // surviving codes per row handled in-kernel; more -> exact SIMT kernel

template <bool BF16>
__global__ void __launch_bounds__(256, 3)
rerank_kernel(const float* __restrict__ z, const __nv_bfloat16* __restrict__ zb, const float* __restrict__ E,
              const __nv_bfloat16* __restrict__ Eb, int64_t n, int D, int nsub, int rpw,
              const float* __restrict__ margin, const uint2* __restrict__ cand, const int* __restrict__ cnt,
              const float* __restrict__ best, int64_t idx_offset, int64_t* __restrict__ idx_out,
              int* __restrict__ fb_rows, uint64_t* __restrict__ fb_packed, int* __restrict__ counters) {
  // rpw = rows per warp step (32 for large n; fewer when n is small so that enough warps are in flight to
  // hide the dependent-load latency of the prune / score chains)
  __shared__ uint32_t s_list[8][RR_LIST];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  // lane groups for the exact scores: lpv lanes cover one row with float4 slices
  int lpv = D >> 2;
  if (lpv > 32) lpv = 32;
  while (lpv & (lpv - 1)) lpv &= lpv - 1;               // largest power of two <= D/4 (D % 4 == 0)
  const int groups = 32 / lpv, gl = lane % lpv, gi = lane / lpv;

  constexpr int RR_NS = 4;                              // slices whose records the one-row-per-warp path keeps in registers
  const bool pre_ok = nsub <= RR_NS;
  for (int64_t base = warp * rpw; base < n; base += nwarps * rpw) {
    const int64_t row = base + lane;
    const bool mine = lane < rpw && row < n;
    uint2 ents[RR_NS];
    int cnts[RR_NS] = {0, 0, 0, 0};
    int ns = 0, ncodes = 0;
    uint32_t first = 0;
    float thr = __int_as_float(0x7fc00000);
    bool fallback = false;
    if (rpw == 1 && base < n) {
      // One row per warp (small n): the whole warp prunes it -- lane j takes record j of each slice (one coalesced
      // 256-byte load per slice) instead of lane 0 walking up to 31 records per slice through dependent loads.
      // With up to RR_NS slices the records are fetched SPECULATIVELY, together with the counts, the slice maxima and
      // the margin (slots beyond a slice's count hold stale workspace bytes and are masked below): one memory round
      // trip instead of one per slice plus two -- at 8192 rows per launch this kernel is a chain of L2 latencies.
      const int64_t r = base;
      float bs = __int_as_float(0xff800000);
      int c_l = 0;
      if (pre_ok) {
#pragma unroll
        for (int sb = 0; sb < RR_NS; ++sb)
          if (sb < nsub && lane < TC_CAND) ents[sb] = cand[(r * nsub + sb) * TC_SLOTS + lane];
      }
      if (lane < nsub) { c_l = cnt[r * nsub + lane]; bs = best[r * nsub + lane]; }
      bool bad = lane < nsub && (c_l < 0 || c_l > TC_CAND || (c_l == 0 && bs != __int_as_float(0xff800000)));
      bad = __any_sync(0xffffffffu, bad);
      float bmax = bs;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) bmax = fmaxf(bmax, __shfl_xor_sync(0xffffffffu, bmax, o));
      const float mg = margin[r];
      if (!bad && mg == mg) {
        thr = bmax - mg;
#pragma unroll
        for (int sb = 0; sb < RR_NS; ++sb) cnts[sb] = sb < nsub ? __shfl_sync(0xffffffffu, c_l, sb) : 0;
        for (int sb = 0; sb < nsub; ++sb) {
          const int c = __shfl_sync(0xffffffffu, c_l, sb);
          uint2 ent = make_uint2(0u, 0xff800000u);
          if (pre_ok) {
#pragma unroll
            for (int q = 0; q < RR_NS; ++q) if (q == sb && lane < c) ent = ents[q];
          } else if (lane < c) {
            ent = cand[(r * nsub + sb) * TC_SLOTS + lane];
          }
          const bool hit = lane < c && __uint_as_float(ent.y) >= thr;
          const unsigned hits = __ballot_sync(0xffffffffu, hit);
          if (hits) {
            if (ns == 0) first = __shfl_sync(0xffffffffu, ent.x, __ffs(hits) - 1);
            ns += __popc(hits);
            int pc = hit ? __popc(ent.x & 0xffu) : 0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) pc += __shfl_xor_sync(0xffffffffu, pc, o);
            ncodes += pc;
          }
        }
      }
      if (lane == 0) {                                   // ns, ncodes, first, thr are warp-uniform here
        if (ns == 1 && ncodes == 1) {
          idx_out[r] = idx_offset + ((first >> 8) << 3) + (__ffs(first & 0xffu) - 1);
        } else if (ns < 1 || ncodes > RR_LIST) {
          fallback = true;
          const int pos = atomicAdd(counters + 0, 1);
          fb_rows[pos] = static_cast<int>(r);
          fb_packed[r] = ~0ull;
        }
      }
    } else if (mine) {
      float bmax = __int_as_float(0xff800000);
      bool bad = false;
      for (int sb = 0; sb < nsub; ++sb) {
        const int c = cnt[row * nsub + sb];
        const float bs = best[row * nsub + sb];
        // an empty list is legitimate only for a slice whose columns are all -inf (masked dead codes, padding)
        bad |= (c < 0 || c > TC_CAND || (c == 0 && bs != __int_as_float(0xff800000)));
        bmax = fmaxf(bmax, bs);
      }
      const float mg = margin[row];
      if (!bad && mg == mg) {
        thr = bmax - mg;
        for (int sb = 0; sb < nsub; ++sb) {
          const int c = cnt[row * nsub + sb];
          const uint2* src = cand + (row * nsub + sb) * TC_SLOTS;
#pragma unroll 4
          for (int j = 0; j < c; ++j) {
            const uint2 ent = src[j];
            if (__uint_as_float(ent.y) >= thr) {       // the group's maximum is still within the margin
              if (ns == 0) first = ent.x;
              ++ns;
              ncodes += __popc(ent.x & 0xffu);
            }
          }
        }
      }
      if (ns == 1 && ncodes == 1) {
        idx_out[row] = idx_offset + ((first >> 8) << 3) + (__ffs(first & 0xffu) - 1);
      } else if (ns < 1 || ncodes > RR_LIST) {
        fallback = true;
        const int pos = atomicAdd(counters + 0, 1);
        fb_rows[pos] = static_cast<int>(row);
        fb_packed[row] = ~0ull;
      }
    }
    const bool multi = mine && !fallback && !(ns == 1 && ncodes == 1);
    unsigned todo = __ballot_sync(0xffffffffu, multi);
    while (todo) {
      const int src_lane = __ffs(todo) - 1;
      todo &= todo - 1;
      const int64_t r = base + src_lane;
      const float thr_r = __shfl_sync(0xffffffffu, thr, src_lane);
      // expand the surviving records of row r into the shared code list
      int n_c = 0;
      for (int sb = 0; sb < nsub; ++sb) {
        int c;
        uint2 ent = make_uint2(0u, 0xff800000u);
        if (rpw == 1 && pre_ok) {                      // the records are still in registers
          c = 0;
#pragma unroll
          for (int q = 0; q < RR_NS; ++q) if (q == sb) { c = cnts[q]; if (lane < c) ent = ents[q]; }
        } else {
          c = cnt[r * nsub + sb];
          if (lane < c) ent = cand[(r * nsub + sb) * TC_SLOTS + lane];
        }
        const uint32_t mk = (lane < c && __uint_as_float(ent.y) >= thr_r) ? (ent.x & 0xffu) : 0u;
        int pre = __popc(mk);                          // exclusive prefix sum of the per-lane code counts
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, pre, o);
          if (lane >= o) pre += v;
        }
        const int tot = __shfl_sync(0xffffffffu, pre, 31);
        int at = n_c + pre - __popc(mk);
        uint32_t m = mk;
        while (m) {
          s_list[wib][at++] = ((ent.x >> 8) << 3) + (__ffs(m) - 1);
          m &= m - 1;
        }
        n_c += tot;
      }
      __syncwarp();
      // Exact scores.  Fast path (fp32 inputs, D <= 512 with a full warp per row, or smaller rows): this lane's slices
      // of the row are loaded ONCE, and all slices of a candidate's code row are in flight together -- the
      // chain is bound by L2 latency (one launch per residual level at 8192 rows), not by arithmetic.
      constexpr int RR_SL = 4, RR_CB = 1;   // slices per lane, candidates per lane group in flight (registers)
      const bool fast = !BF16 && D <= lpv * 4 * RR_SL;
      float4 zr[RR_SL];
      if (fast) {
#pragma unroll
        for (int sl = 0; sl < RR_SL; ++sl) {
          const int d = gl * 4 + sl * lpv * 4;
          if (d < D) zr[sl] = *reinterpret_cast<const float4*>(z + r * D + d);
        }
      }
      double top = -1e300;
      uint32_t top_idx = 0xffffffffu;
      for (int c0 = 0; c0 < n_c; c0 += groups * RR_CB) {
        uint32_t codes[RR_CB];
        bool oks[RR_CB];
        float4 er[RR_CB][RR_SL];
#pragma unroll
        for (int u = 0; u < RR_CB; ++u) {
          const int ci = c0 + u * groups + gi;
          oks[u] = ci < n_c;
          codes[u] = oks[u] ? s_list[wib][ci] : 0u;
          if (fast && oks[u]) {
#pragma unroll
            for (int sl = 0; sl < RR_SL; ++sl) {
              const int d = gl * 4 + sl * lpv * 4;
              if (d < D) er[u][sl] = __ldg(reinterpret_cast<const float4*>(E + static_cast<int64_t>(codes[u]) * D + d));
            }
          }
        }
#pragma unroll
        for (int u = 0; u < RR_CB; ++u) {
          const bool ok = oks[u];
          const uint32_t code = codes[u];
          double dot = 0.0, ee = 0.0;
          if (ok && fast) {
#pragma unroll
            for (int sl = 0; sl < RR_SL; ++sl) {
              if (gl * 4 + sl * lpv * 4 < D) {
                const float zv[4] = {zr[sl].x, zr[sl].y, zr[sl].z, zr[sl].w};
                const float ev[4] = {er[u][sl].x, er[u][sl].y, er[u][sl].z, er[u][sl].w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  dot = fma(static_cast<double>(zv[q]), static_cast<double>(ev[q]), dot);
                  ee = fma(static_cast<double>(ev[q]), static_cast<double>(ev[q]), ee);
                }
              }
            }
          } else if (ok) {
            for (int d = gl * 4; d < D; d += lpv * 4) {
              float zv[4], ev[4];
              if (BF16) {
                const uint2 a = *reinterpret_cast<const uint2*>(zb + r * D + d);
                const uint2 b = *reinterpret_cast<const uint2*>(Eb + static_cast<int64_t>(code) * D + d);
                zv[0] = __uint_as_float(a.x << 16); zv[1] = __uint_as_float(a.x & 0xffff0000u);
                zv[2] = __uint_as_float(a.y << 16); zv[3] = __uint_as_float(a.y & 0xffff0000u);
                ev[0] = __uint_as_float(b.x << 16); ev[1] = __uint_as_float(b.x & 0xffff0000u);
                ev[2] = __uint_as_float(b.y << 16); ev[3] = __uint_as_float(b.y & 0xffff0000u);
              } else {
                *reinterpret_cast<float4*>(zv) = *reinterpret_cast<const float4*>(z + r * D + d);
                *reinterpret_cast<float4*>(ev) = __ldg(reinterpret_cast<const float4*>(E + static_cast<int64_t>(code) * D + d));
              }
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                dot = fma(static_cast<double>(zv[q]), static_cast<double>(ev[q]), dot);
                ee = fma(static_cast<double>(ev[q]), static_cast<double>(ev[q]), ee);
              }
            }
          }
          for (int o = lpv >> 1; o > 0; o >>= 1) {       // reduce inside the lane group
            dot += __shfl_xor_sync(0xffffffffu, dot, o);
            ee += __shfl_xor_sync(0xffffffffu, ee, o);
          }
          double sc = ok ? dot - 0.5 * ee : -1e300;
          uint32_t cd = ok ? code : 0xffffffffu;
          for (int o = lpv; o < 32; o <<= 1) {           // lexicographic (score, lowest index) across groups
            const double so = __shfl_xor_sync(0xffffffffu, sc, o);
            const uint32_t co = __shfl_xor_sync(0xffffffffu, cd, o);
            if (so > sc || (so == sc && co < cd)) { sc = so; cd = co; }
          }
          if (sc > top || (sc == top && cd < top_idx)) { top = sc; top_idx = cd; }
        }
      }
      if (lane == 0) idx_out[r] = idx_offset + top_idx;
      __syncwarp();
    }
  }
}

__global__ void fb_unpack_kernel(const int* __restrict__ fb_rows, const uint64_t* __restrict__ fb_packed,
                                 const int* __restrict__ counters, int64_t* __restrict__ idx_out) {
  const int n = counters[0];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int row = fb_rows[i];
    idx_out[row] = static_cast<int64_t>(fb_packed[row] & 0xffffffffull);
  }
}

// ------------------------------------------------------------------------------------ host side
static bool make_map(CUtensorMap* m, const void* base, int64_t rows, int D, int box_rows, bool f16) {
  return make_tensor_map_2d(m, base, rows, D, box_rows,
                            f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2);
}

struct TcPlan {
  int BM, stages, smem_bytes, zbufs;
  int64_t chunk_rows;
  int ksplit_max;
};

static bool tc_plan(int64_t N, int K, int D, TcPlan* pl) {
  if (D % TC_KB != 0 || D < 64 || D > 512 || K < TC_BN || N < 64) return false;
  pl->BM = (D <= 256) ? 256 : 128;
  const int TC_CS = tc_cs(pl->BM);
  const int nepi = (pl->BM / 32) * TC_CS;
  const int ztile = pl->BM * D * 2;
  pl->zbufs = (2 * ztile + 4 * TC_STAGE_BYTES + nepi * 2 * (TC_BN / TC_CS) * 4 + 2048 <= TC_SMEM_LIMIT) ? 2 : 1;
  const int fixed = 1024 + pl->zbufs * ztile + nepi * 2 * (TC_BN / TC_CS) * 4 + 256;
  int st = (TC_SMEM_LIMIT - fixed) / TC_STAGE_BYTES;
  if (st > 8) st = 8;
  if (st < 2) return false;
  pl->stages = st;
  pl->smem_bytes = fixed + st * TC_STAGE_BYTES;
  // rows per pass.  Small D is HBM-bound: keep one chunk's fp32 rows + bf16 copy within the 126 MB L2 so
  // the pre-pass output and the re-rank / gather re-reads hit L2.  Large D is tensor-bound: big chunks
  // (many waves per launch), bounded only by the workspace.
  int64_t chunk;
  if (D <= 128) {
    chunk = (static_cast<int64_t>(96) << 20) / (static_cast<int64_t>(D) * 6);
    chunk = (chunk / 4096) * 4096;
    if (chunk < 65536) chunk = 65536;
  } else {
    chunk = 1 << 20;
  }
  if (chunk > (1 << 20)) chunk = 1 << 20;
  pl->chunk_rows = chunk;
  pl->ksplit_max = 8;
  return true;
}

// Split the code range over CTAs when there are too few row tiles to fill the SMs; never leaves a
// split empty.  Returns the number of splits and sets *tiles_per_split.
static int pick_ksplit(int64_t rows, int BM, int code_tiles, int ksplit_max, int* tiles_per_split) {
  const int64_t row_tiles = (rows + BM - 1) / BM;
  int want = 1;
  while (want * 2 <= ksplit_max && want * 2 <= code_tiles && row_tiles * want * 2 <= kNumSMs) want *= 2;
  const int tps = (code_tiles + want - 1) / want;
  *tiles_per_split = tps;
  return (code_tiles + tps - 1) / tps;
}

bool tc_supported(int64_t N, int K, int D) {
  const char* f = std::getenv("VQB200_FORCE_SIMT");
  if (f && f[0] == '1') return false;
  TcPlan pl;
  return tc_plan(N, K, D, &pl);
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
// (row, split) slots: code splits are only used while row_tiles * ksplit <= #SMs
static size_t tc_slots(int64_t rows, int BM) {
  const size_t few = static_cast<size_t>(kNumSMs) * BM;
  (void)BM;
  return (static_cast<size_t>(rows) > few ? static_cast<size_t>(rows) : few) * 2;   // x column slices (<= 2)
}

// Workspace of ONE chunk; two chunks' worth is laid out when the rows span several chunks (software pipeline).
static size_t tc_set_bytes(int64_t rows, int D, int BM) {
  const size_t slots = tc_slots(rows, BM);
  size_t b = 256;                                                   // counters
  b += align_up(static_cast<size_t>(rows) * D * 2, 256);            // zb
  b += align_up(static_cast<size_t>(rows) * 4, 256);                // margin
  b += align_up(slots * 4, 256) * 2;                                // cnt, best
  b += align_up(slots * TC_SLOTS * 8, 256);                         // cand
  b += align_up(static_cast<size_t>(rows) * 4, 256);                // fb_rows
  b += align_up(static_cast<size_t>(rows) * 8, 256);                // fb_packed
  return b;
}

size_t tc_workspace_bytes(int64_t N, int K, int D) {
  TcPlan pl;
  if (!tc_plan(N, K, D, &pl)) return 0;
  const int64_t rows = N < pl.chunk_rows ? N : pl.chunk_rows;
  return tc_set_bytes(rows, D, pl.BM) * (N > pl.chunk_rows ? 2 : 1);
}

int tc_launches(int64_t N, int K, int D);

// Helper streams for the chunk pipeline (created once per device, never destroyed).
struct TcPipe {
  bool ready = false;
  cudaStream_t s1 = nullptr, s2 = nullptr, stc = nullptr;
  cudaEvent_t fork = nullptr, prep[2] = {nullptr, nullptr}, tc[2] = {nullptr, nullptr}, done[2] = {nullptr, nullptr};
};
static TcPipe* tc_pipe() {
  static TcPipe pipes[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  TcPipe& p = pipes[dev];
  if (!p.ready) {
    // The tensor kernel gets its own HIGH-priority stream: when its CTAs and side-kernel blocks are both
    // pending, the CTAs are placed first and the side blocks fill what is left of each SM.
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    bool ok = cudaStreamCreateWithPriority(&p.s1, cudaStreamNonBlocking, lo) == cudaSuccess &&
              cudaStreamCreateWithPriority(&p.s2, cudaStreamNonBlocking, lo) == cudaSuccess &&
              cudaStreamCreateWithPriority(&p.stc, cudaStreamNonBlocking, hi) == cudaSuccess &&
              cudaEventCreateWithFlags(&p.fork, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 2 && ok; ++i)
      ok = cudaEventCreateWithFlags(&p.prep[i], cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&p.tc[i], cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&p.done[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) return nullptr;
    p.ready = true;
  }
  return &p;
}

struct TcSet {
  int* counters; __nv_bfloat16* zb; float* margin; int* cnt; float* best; uint2* cand; int* fb_rows; uint64_t* fb_packed;
};
static TcSet carve(uint8_t* w, int64_t cap, int D, int BM) {
  const size_t slots = tc_slots(cap, BM);
  TcSet t;
  t.counters = reinterpret_cast<int*>(w); w += 256;      // [0] rows handed to the SIMT kernel
  t.zb = reinterpret_cast<__nv_bfloat16*>(w); w += align_up(static_cast<size_t>(cap) * D * 2, 256);
  t.margin = reinterpret_cast<float*>(w); w += align_up(static_cast<size_t>(cap) * 4, 256);
  t.cnt = reinterpret_cast<int*>(w); w += align_up(slots * 4, 256);
  t.best = reinterpret_cast<float*>(w); w += align_up(slots * 4, 256);
  t.cand = reinterpret_cast<uint2*>(w); w += align_up(slots * TC_SLOTS * 8, 256);
  t.fb_rows = reinterpret_cast<int*>(w); w += align_up(static_cast<size_t>(cap) * 4, 256);
  t.fb_packed = reinterpret_cast<uint64_t*>(w);
  return t;
}

#define VQ_CUDA(call)                         \
  do {                                        \
    cudaError_t e_ = (call);                  \
    if (e_ != cudaSuccess) return status_of(e_); \
  } while (0)

// Does the side-job pipeline apply?  Several chunks, every chunk large enough for the CTA-pair kernel, D a multiple
// of 128 (one warp covers a row with whole float4 slices), not switched off (VQB200_TC2_SIDE=0).
static bool tc_pair_allowed(int K, int D, int* stages2_out, int* zbufs2_out) {
  const char* env2 = std::getenv("VQB200_TC2");
  const bool want2 = !(env2 && env2[0] == '0');
  const int z2 = P2_ROWS * D * 2;
  const int zbufs2 = (2 * z2 + 4 * TC_STAGE_BYTES + P2_NEPI * P2_WCOLS * 4 + 2048 <= TC_SMEM_LIMIT) ? 2 : 1;
  int stages2 = (TC_SMEM_LIMIT - (1024 + zbufs2 * z2 + P2_NEPI * P2_WCOLS * 4 + 256)) / TC_STAGE_BYTES;
  if (stages2 > 8) stages2 = 8;
  if (stages2_out) *stages2_out = stages2;
  if (zbufs2_out) *zbufs2_out = zbufs2;
  return want2 && stages2 >= 3 && K >= P2_BN;
}
static constexpr int64_t kPairMinRows = static_cast<int64_t>(kNumSMs / 2) * 2 * P2_ROWS;

bool tc_side_pipeline(int64_t N, int K, int D) {
  TcPlan pl;
  if (!tc_plan(N, K, D, &pl)) return false;
  // OPT-IN (VQB200_TC2_SIDE=1 | p | g | i).  Measured on the B200 pool (profiles/r02_c3_side_variants.txt, c3,
  // 2^22 rows): the tensor kernel runs power-capped at ~1.39 GHz effective (clock64 / globaltimer), 4.2 M cycles per
  // 2^20-row launch; with the side warps streaming it needs 5.6 M cycles at the same clock (4.0 ms instead of 3.4)
  // and the step goes from 17.4 to 18.3 ms -- everything that shares an SM with the epilogue costs more than the
  // overlap returns, exactly as the co-resident side KERNELS did in round 1.  Kept as a tested switch.
  const char* es = std::getenv("VQB200_TC2_SIDE");
  if (!es || es[0] == '0') return false;
  if (N <= pl.chunk_rows || D % 128 != 0) return false;
  const int64_t last = N - ((N - 1) / pl.chunk_rows) * pl.chunk_rows;
  return tc_pair_allowed(K, D, nullptr, nullptr) && pl.chunk_rows >= kPairMinRows && last >= kPairMinRows;
}

int tc_launches(int64_t N, int K, int D) {
  TcPlan pl;
  if (!tc_plan(N, K, D, &pl)) return 0;
  const int chunks = static_cast<int>((N + pl.chunk_rows - 1) / pl.chunk_rows);
  // per chunk: pre-pass, tcgen05 search, re-rank, SIMT hand-back, unpack; with the side-job pipeline only the first
  // chunk has a pre-pass of its own
  return tc_side_pipeline(N, K, D) ? chunks * 4 + 1 : chunks * 5;
}

// The search over N rows runs chunk by chunk through pre-pass -> tcgen05 kernel -> re-rank (+ exact hand-back)
// [-> gather].  Two schedules:
//  * side-job pipeline (several chunks, CTA-pair kernel, D % 128 == 0): ONE stream; the tensor kernel of chunk i
//    carries, on six extra warps, the pre-pass of chunk i+1 and the gather of chunk i-1 (see SideJobs), so only the
//    first pre-pass, the small re-rank kernels and the last gather run outside a tensor kernel;
//  * otherwise: the stages of neighbouring chunks on three streams (they serialise against the persistent tensor
//    kernel, but the launch gaps overlap), joined back into the caller's stream before returning.
int launch_search_tc(const float* z, int64_t N, int D, const float* E, const uint16_t* E_bf16, const float* ee_half,
                     const float* ee_half_bf16, const float* level_meta, int K, int mode, int64_t idx_offset,
                     int64_t* idx_out, void* workspace, size_t workspace_bytes, cudaStream_t s, const GatherArgs* ga,
                     const PrepArgs* prep) {
  TcPlan pl;
  if (!tc_plan(N, K, D, &pl)) return VQB200_ESHAPE;
  if (workspace_bytes < tc_workspace_bytes(N, K, D)) return VQB200_EWORKSPACE;
  const bool bf = mode == VQB200_MODE_BF16_INPUT;
  const int64_t cap = N < pl.chunk_rows ? N : pl.chunk_rows;
  const int n_chunks = static_cast<int>((N + pl.chunk_rows - 1) / pl.chunk_rows);
  const bool side = !prep && tc_side_pipeline(N, K, D);
  // measurement switches: VQB200_TC2_SIDE=i keeps the sixteen-warp kernel but leaves its side warps idle (the
  // pre-pass and the gather run as separate kernels), =g / =p give them only the gather / only the pre-pass
  const char* esv = std::getenv("VQB200_TC2_SIDE");
  const char sv = esv ? esv[0] : '0';
  const bool side_gather_on = sv != 'i' && sv != 'p', side_prep_on = sv != 'i' && sv != 'g';
  const char* edbg = std::getenv("VQB200_DEBUG");
  const bool clk_dbg = edbg && edbg[0] == '4';
  TcPipe* pipe = (n_chunks > 1 && !side) ? tc_pipe() : nullptr;
  const bool piped = pipe != nullptr;
  // VQB200_TC_PRIO=1 runs the tensor kernel on a high-priority helper stream (measured: 17.9 ms against 17.4 ms
  // per c3 step on the caller's stream -- nothing co-resides with the persistent kernel, so priority only adds
  // event hops); the default keeps it on the caller's stream.
  const char* envp = std::getenv("VQB200_TC_PRIO");
  const bool prio = piped && envp && envp[0] == '1';
  cudaStream_t s_prep = piped ? pipe->s1 : s, s_rr = piped ? pipe->s2 : s, s_tc = prio ? pipe->stc : s;

  const size_t set_bytes = tc_set_bytes(cap, D, pl.BM);
  TcSet sets[2];
  sets[0] = carve(static_cast<uint8_t*>(workspace), cap, D, pl.BM);
  sets[1] = n_chunks > 1 ? carve(static_cast<uint8_t*>(workspace) + set_bytes, cap, D, pl.BM) : sets[0];

  CUtensorMap map_e;
  if (!make_map(&map_e, E_bf16, K, D, TC_BN, !bf)) return VQB200_EDRIVER;
  const int code_tiles = (K + TC_BN - 1) / TC_BN;

  // CTA-pair kernel (cta_group::2): large row counts only (no code splits), z tile of 128 rows must fit
  int stages2 = 0, zbufs2 = 1;
  const int z2 = P2_ROWS * D * 2;
  const bool use2 = tc_pair_allowed(K, D, &stages2, &zbufs2) && cap >= kPairMinRows;
  // VQB200_TC2_REGS=128 selects the register-capped build under which one side-KERNEL block per SM co-resides
  // with the persistent tensor kernel.  Measured (c3, 2^22 rows): 19.2 ms per step against 17.4-17.9 ms -- the
  // side kernels' warps take issue slots and LSU bandwidth from the epilogue (tensor kernel 2.9 -> 3.8 ms), which
  // costs more than the overlap returns.  (The side-job pipeline above is the design that replaced it.)
  const char* envr = std::getenv("VQB200_TC2_REGS");
  const bool slim = envr && envr[0] == '1' && envr[1] == '2';
  static bool attr2_done_dev[64] = {};
  bool& attr2_done = attr2_done_dev[current_device_slot()];
  if (use2 && !attr2_done) {
    VQ_CUDA(cudaFuncSetAttribute(search_tc2_kernel<168>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
    VQ_CUDA(cudaFuncSetAttribute(search_tc2_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
    VQ_CUDA(cudaFuncSetAttribute(search_tc2_side_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
    attr2_done = true;
  }
  static bool attr_done_dev[64][2] = {};
  bool (&attr_done)[2] = attr_done_dev[current_device_slot()];
  if (!attr_done[pl.BM == 256]) {
    VQ_CUDA(pl.BM == 256
        ? cudaFuncSetAttribute(search_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT)
        : cudaFuncSetAttribute(search_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
    attr_done[pl.BM == 256] = true;
  }
  if (piped) {
    VQ_CUDA(cudaEventRecord(pipe->fork, s));
    VQ_CUDA(cudaStreamWaitEvent(s_prep, pipe->fork, 0));
    VQ_CUDA(cudaStreamWaitEvent(s_rr, pipe->fork, 0));
    if (prio) VQ_CUDA(cudaStreamWaitEvent(s_tc, pipe->fork, 0));
  }
  auto chunk_rows_of = [&](int c) -> int64_t {
    const int64_t r0 = static_cast<int64_t>(c) * pl.chunk_rows;
    return (N - r0) < pl.chunk_rows ? (N - r0) : pl.chunk_rows;
  };
  auto launch_prep = [&](int c, cudaStream_t st) {
    const int64_t rows = chunk_rows_of(c);
    int64_t blocks = (rows + 7) / 8;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    const TcSet& w = sets[n_chunks > 1 ? (c & 1) : 0];
    zprep_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(z + static_cast<int64_t>(c) * pl.chunk_rows * D, rows, D,
                                                                 mode, level_meta, w.zb, w.margin, nullptr, nullptr, 0, nullptr);
  };
  auto launch_gather_chunk = [&](int c, cudaStream_t st) -> int {
    const int64_t r0 = static_cast<int64_t>(c) * pl.chunk_rows;
    return launch_gather(z + r0 * D, ga->E_full, idx_out + r0, chunk_rows_of(c), D, ga->K_total,
                         ga->zq_out ? ga->zq_out + r0 * D : nullptr, 0, ga->zq_st_out ? ga->zq_st_out + r0 * D : nullptr,
                         nullptr, ga->sqerr_sum, ga->hist, ga->row_mask ? ga->row_mask + r0 : nullptr, st);
  };

  int pending_gather = -1;                             // side pipeline: chunk whose gather has not been issued yet
  bool next_prepped = false;                           // side pipeline: the pre-pass of this chunk rode on the previous kernel
  int ci = 0;
  for (int64_t r0 = 0; r0 < N; r0 += pl.chunk_rows, ++ci) {
    const int64_t rows = chunk_rows_of(ci);
    const float* zc = z + r0 * D;
    const int b = n_chunks > 1 ? (ci & 1) : 0;
    const TcSet& w = sets[b];

    // ---- stage 1: pre-pass (needs the set free: the re-rank of chunk ci-2 has finished with it)
    if (piped && ci >= 2) VQ_CUDA(cudaStreamWaitEvent(s_prep, pipe->done[b], 0));
    VQ_CUDA(cudaMemsetAsync(w.counters, 0, 2 * sizeof(int), s_prep));
    if (!prep && !(side && next_prepped)) launch_prep(ci, s_prep);
    VQ_CUDA(cudaGetLastError());
    const __nv_bfloat16* zb_c = prep ? reinterpret_cast<const __nv_bfloat16*>(prep->z16) + r0 * D : w.zb;
    const float* mg_c = prep ? prep->margin + r0 : w.margin;
    if (piped) {
      VQ_CUDA(cudaEventRecord(pipe->prep[b], s_prep));
      VQ_CUDA(cudaStreamWaitEvent(s_tc, pipe->prep[b], 0));
    }

    // ---- stage 2: tensor-core candidates (caller's stream)
    CUtensorMap map_z;
    TcParams p;
    p.n_rows = rows; p.D = D; p.K = K;
    p.ee_half = bf ? ee_half_bf16 : ee_half;
    p.margin = mg_c; p.cand = w.cand; p.cnt = w.cnt; p.best = w.best;
    p.clk = nullptr;
    if (clk_dbg) cudaMallocManaged(&p.clk, kNumSMs * 4 * sizeof(long long));
    const bool pair_now = use2 && rows >= kPairMinRows;
    int nsub;
    if (pair_now) {
      if (!make_map(&map_z, zb_c, rows, D, P2_ROWS, !bf)) return VQB200_EDRIVER;
      p.idesc = bf ? kIdescPair : kIdescPairF16;
      p.row_tiles = static_cast<int>((rows + 2 * P2_ROWS - 1) / (2 * P2_ROWS));
      p.ksplit = 1; p.tiles_per_split = (K + P2_BN - 1) / P2_BN;
      p.code_tiles = (K + P2_BN - 1) / P2_BN;
      p.stages = stages2; p.zbufs = zbufs2;
      nsub = P2_CS;
      const int pairs = p.row_tiles < kNumSMs / 2 ? p.row_tiles : kNumSMs / 2;
      const int smem2 = 1024 + zbufs2 * z2 + stages2 * TC_STAGE_BYTES + P2_NEPI * P2_WCOLS * 4 + 256;
      if (side) {
        SideJobs sj;
        sj.D = D; sj.mode = mode;
        sj.g_rows = 0; sj.p_rows = 0;
        sj.g_z = nullptr; sj.g_idx = nullptr; sj.g_E = nullptr; sj.g_K_total = 0; sj.g_zq = nullptr; sj.g_zq_st = nullptr;
        sj.g_sqerr = nullptr; sj.g_hist = nullptr; sj.g_mask = nullptr;
        sj.p_z = nullptr; sj.p_zb = nullptr; sj.p_margin = nullptr; sj.level_meta = level_meta;
        if (ga && pending_gather >= 0 && !side_gather_on) {
          const int gs = launch_gather_chunk(pending_gather, s);
          if (gs != VQB200_OK) return gs;
          pending_gather = -1;
        }
        if (ga && pending_gather >= 0) {               // gather of the previous chunk: its indices are final
          const int64_t g0 = static_cast<int64_t>(pending_gather) * pl.chunk_rows;
          sj.g_z = z + g0 * D; sj.g_idx = idx_out + g0; sj.g_rows = chunk_rows_of(pending_gather);
          sj.g_E = ga->E_full; sj.g_K_total = ga->K_total;
          sj.g_zq = ga->zq_out ? ga->zq_out + g0 * D : nullptr;
          sj.g_zq_st = ga->zq_st_out ? ga->zq_st_out + g0 * D : nullptr;
          sj.g_sqerr = ga->sqerr_sum; sj.g_hist = ga->hist;
          sj.g_mask = ga->row_mask ? ga->row_mask + g0 : nullptr;
          pending_gather = -1;
        }
        next_prepped = false;
        if (ci + 1 < n_chunks && side_prep_on) {       // pre-pass of the next chunk into the other workspace set
          const TcSet& wn = sets[(ci + 1) & 1];
          sj.p_z = z + (r0 + pl.chunk_rows) * D; sj.p_rows = chunk_rows_of(ci + 1);
          sj.p_zb = wn.zb; sj.p_margin = wn.margin;
          next_prepped = true;
        }
        timing_mark_begin(s_tc);
        search_tc2_side_kernel<<<2 * pairs, 64 + P2_NEPI * 32 + P2_SIDE * 32, smem2, s_tc>>>(map_z, map_e, p, sj);
        timing_mark_end(s_tc);
      } else {
        timing_mark_begin(s_tc);
        if (slim) search_tc2_kernel<128><<<2 * pairs, 64 + P2_NEPI * 32, smem2, s_tc>>>(map_z, map_e, p);
        else search_tc2_kernel<168><<<2 * pairs, 64 + P2_NEPI * 32, smem2, s_tc>>>(map_z, map_e, p);
        timing_mark_end(s_tc);
      }
    } else {
      if (!make_map(&map_z, zb_c, rows, D, pl.BM, !bf)) return VQB200_EDRIVER;
      p.idesc = bf ? kIdesc : kIdescF16;
      p.row_tiles = static_cast<int>((rows + pl.BM - 1) / pl.BM);
      p.ksplit = pick_ksplit(rows, pl.BM, code_tiles, pl.ksplit_max, &p.tiles_per_split);
      p.code_tiles = code_tiles;
      p.stages = pl.stages;
      p.zbufs = pl.zbufs;
      nsub = p.ksplit * tc_cs(pl.BM);
      const int items = p.row_tiles * p.ksplit;
      const int grid = items < kNumSMs ? items : kNumSMs;
      timing_mark_begin(s_tc);
      if (pl.BM == 256)
        search_tc_kernel<256><<<grid, 64 + 256 * tc_cs(256), pl.smem_bytes, s_tc>>>(map_z, map_e, p);
      else
        search_tc_kernel<128><<<grid, 64 + 128 * tc_cs(128), pl.smem_bytes, s_tc>>>(map_z, map_e, p);
      timing_mark_end(s_tc);
    }
    VQ_CUDA(cudaGetLastError());
    if (clk_dbg) {                                     // effective SM clock of this launch: cycles / wall time per CTA
      cudaStreamSynchronize(s_tc);
      double mhz = 0.0, us = 0.0;
      const int nb = pair_now ? 2 * (p.row_tiles < kNumSMs / 2 ? p.row_tiles : kNumSMs / 2) : 0;
      for (int i = 0; i < nb; ++i) {
        const double dc = static_cast<double>(p.clk[i * 4 + 2] - p.clk[i * 4 + 0]);
        const double dt = static_cast<double>(p.clk[i * 4 + 3] - p.clk[i * 4 + 1]);
        mhz += dc / dt * 1e3 / nb; us += dt * 1e-3 / nb;
      }
      if (nb) fprintf(stderr, "[vqb200] pair kernel chunk %d: %.1f us per CTA, effective SM clock %.0f MHz\n", ci, us, mhz);
      cudaFree(p.clk);
    }
    if (piped) {
      VQ_CUDA(cudaEventRecord(pipe->tc[b], s_tc));
      VQ_CUDA(cudaStreamWaitEvent(s_rr, pipe->tc[b], 0));
    }

    // ---- stage 3: prune + exact re-rank, then the rows handed back to the exact SIMT kernel
    int rpw = 32;                                      // rows per warp step: keep >= ~8K warps in flight
    while (rpw > 1 && rows / rpw < 8192) rpw >>= 1;
    int64_t blocks = (rows + 8 * rpw - 1) / (8 * rpw);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    const __nv_bfloat16* Eb = reinterpret_cast<const __nv_bfloat16*>(E_bf16);
    if (bf)
      rerank_kernel<true><<<static_cast<unsigned>(blocks), 256, 0, s_rr>>>(zc, zb_c, E, Eb, rows, D, nsub, rpw, mg_c,
                                                                          w.cand, w.cnt, w.best, idx_offset, idx_out + r0,
                                                                          w.fb_rows, w.fb_packed, w.counters);
    else
      rerank_kernel<false><<<static_cast<unsigned>(blocks), 256, 0, s_rr>>>(zc, zb_c, E, Eb, rows, D, nsub, rpw, mg_c,
                                                                           w.cand, w.cnt, w.best, idx_offset, idx_out + r0,
                                                                           w.fb_rows, w.fb_packed, w.counters);
    VQ_CUDA(cudaGetLastError());
    const int st = launch_search_simt_list(zc, w.fb_rows, w.counters, rows, D, E, bf ? ee_half_bf16 : ee_half, K,
                                           bf ? 1 : 0, idx_offset, w.fb_packed, s_rr);
    if (st != VQB200_OK) return st;
    fb_unpack_kernel<<<64, 256, 0, s_rr>>>(w.fb_rows, w.fb_packed, w.counters, idx_out + r0);
    VQ_CUDA(cudaGetLastError());
    if (ga) {   // stage 4 (optional): gather / straight-through / loss / histogram of this chunk, behind its re-rank
      if (side) {
        pending_gather = ci;                           // rides on the next chunk's tensor kernel
      } else {
        const int gs = launch_gather_chunk(ci, s_rr);
        if (gs != VQB200_OK) return gs;
      }
    }
    if (piped) VQ_CUDA(cudaEventRecord(pipe->done[b], s_rr));
  }
  if (ga && pending_gather >= 0) {                     // the last chunk's gather has no tensor kernel to ride on
    const int gs = launch_gather_chunk(pending_gather, s);
    if (gs != VQB200_OK) return gs;
  }
  if (piped) {                                         // join: the caller's stream sees every chunk finished
    VQ_CUDA(cudaStreamWaitEvent(s, pipe->done[0], 0));
    if (n_chunks > 1) VQ_CUDA(cudaStreamWaitEvent(s, pipe->done[1], 0));
  }
  return VQB200_OK;
}

}  // namespace vqb
