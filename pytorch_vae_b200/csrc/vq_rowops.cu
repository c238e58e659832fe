// HBM-bound row kernels of the VQ path: codebook cache refresh (+ fused EMA finalize), the
// gather / straight-through / residual / loss / histogram pass, statistics, scatter-add,
// backward, and the index wire formats.  All are coalesced 128-bit streaming kernels sized
// in multiples of the SM count; none of them is GEMM-shaped.
#include "common.cuh"

namespace vqb {

// --------------------------------------------------------------------------------------------
// codebook refresh: one warp per code row
// --------------------------------------------------------------------------------------------
// MODE 0: derive the cache from E.  MODE 1: EMA update (models/vq_vae.py:85-89) then the cache.
// MODE 2: Lloyd step of the k-means initialiser -- E <- segment mean where the segment is non-empty (an empty
// cluster keeps its centroid), then the cache.
// MODE 1 with chain_phase != 0 (training forward of a residual codebook, vq_rvq_fused.cuh): the reference runs one EMA
// update per LEVEL and every update touches all K_total codes; for the codes of another level it is a decay-only step
// (empty one-hot columns: n = 0, s = 0).  Phase 1 applies to a code of level l the l decay-only steps of the earlier
// levels (the state the level is searched in; level 0 keeps its E untouched), phase 2 its own update from the segment
// sums followed by the L - 1 - l decay-only steps of the later levels.  Each step is the arithmetic of :85-88.
template <int MODE>
__global__ void __launch_bounds__(256)
codebook_refresh_kernel(const float* __restrict__ seg_sum, const float* __restrict__ seg_cnt, float decay,
                        float omd, float eps, int K_total, int D, int K_per, float* ema_cs, float* ema_emb,
                        float* E, uint16_t* __restrict__ E_bf16, float* __restrict__ ee_half,
                        float* __restrict__ level_meta, int chain_phase, int row_base) {
  const int lane = threadIdx.x & 31;
  const int row = row_base + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int D4 = D >> 2;
  __shared__ int s_red[8][6];
  if (row < K_total) {
  constexpr bool EMA = MODE == 1;
  float denom = 1.f;
  if (MODE == 2) denom = seg_cnt[row];
  // decay-only steps before / after the real one, and whether the real one runs
  const int lvl_row = row / K_per, n_levels = K_total / K_per;
  const int n_pre = (EMA && chain_phase == 1) ? lvl_row : 0;
  const int n_post = (EMA && chain_phase == 2) ? n_levels - 1 - lvl_row : 0;
  const bool real = EMA && chain_phase != 1;
  const bool touch = EMA && (real || n_pre > 0);           // phase 1 leaves level 0 exactly as it is
  const float zero_term = __fmul_rn(0.f, omd);             // n (1 - decay) of an empty column
  if (touch) {
    // models/vq_vae.py:85: cs.mul_(decay).add_(n * (1 - decay)) -- two roundings, no fma contraction
    float cs = ema_cs[row];
    for (int i = 0; i < n_pre; ++i) cs = __fadd_rn(__fmul_rn(cs, decay), zero_term);
    if (real) cs = __fadd_rn(__fmul_rn(cs, decay), __fmul_rn(seg_cnt[row], omd));
    for (int i = 0; i < n_post; ++i) cs = __fadd_rn(__fmul_rn(cs, decay), zero_term);
    __syncwarp();
    if (lane == 0) ema_cs[row] = cs;
    denom = __fadd_rn(cs, eps);
  }
  double acc = 0.0, accb = 0.0, accd = 0.0, acch = 0.0, accdh = 0.0;
  bool bad = false;
  uint16_t* E_f16 = E_bf16 + static_cast<int64_t>(K_total) * D;        // operand plane 1
  // four float4 slices of the row per lane and pass: every load of the pass is issued before the first use (one memory
  // round trip per 512 floats of the row instead of one per 128: the kernel is latency-, not bandwidth-bound)
  for (int c0 = lane; c0 < D4; c0 += 128) {
    float4 in_a[4], in_s[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = c0 + 32 * u;
      if (c < D4) {
        const int64_t o = static_cast<int64_t>(row) * D4 + c;
        if (touch) {
          in_a[u] = reinterpret_cast<const float4*>(ema_emb)[o];
          if (real) in_s[u] = reinterpret_cast<const float4*>(seg_sum)[o];
        } else if (MODE == 2 && denom > 0.f) {
          in_s[u] = reinterpret_cast<const float4*>(seg_sum)[o];
        } else {
          in_a[u] = reinterpret_cast<const float4*>(E)[o];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
    const int c = c0 + 32 * u;
    if (c >= D4) continue;
    const int64_t o = static_cast<int64_t>(row) * D4 + c;
    float4 e;
    if (touch) {
      float4 n = in_a[u];
      for (int i = 0; i < n_pre; ++i) {
        n.x = __fadd_rn(__fmul_rn(n.x, decay), zero_term); n.y = __fadd_rn(__fmul_rn(n.y, decay), zero_term);
        n.z = __fadd_rn(__fmul_rn(n.z, decay), zero_term); n.w = __fadd_rn(__fmul_rn(n.w, decay), zero_term);
      }
      if (real) {
        const float4 s = in_s[u];
        n.x = __fadd_rn(__fmul_rn(n.x, decay), __fmul_rn(s.x, omd));
        n.y = __fadd_rn(__fmul_rn(n.y, decay), __fmul_rn(s.y, omd));
        n.z = __fadd_rn(__fmul_rn(n.z, decay), __fmul_rn(s.z, omd));
        n.w = __fadd_rn(__fmul_rn(n.w, decay), __fmul_rn(s.w, omd));
      }
      for (int i = 0; i < n_post; ++i) {
        n.x = __fadd_rn(__fmul_rn(n.x, decay), zero_term); n.y = __fadd_rn(__fmul_rn(n.y, decay), zero_term);
        n.z = __fadd_rn(__fmul_rn(n.z, decay), zero_term); n.w = __fadd_rn(__fmul_rn(n.w, decay), zero_term);
      }
      reinterpret_cast<float4*>(ema_emb)[o] = n;
      e.x = __fdiv_rn(n.x, denom); e.y = __fdiv_rn(n.y, denom);       // :88 E = es / (cs + eps)
      e.z = __fdiv_rn(n.z, denom); e.w = __fdiv_rn(n.w, denom);
      reinterpret_cast<float4*>(E)[o] = e;
    } else if (MODE == 2 && denom > 0.f) {
      const float4 sm = in_s[u];
      e.x = __fdiv_rn(sm.x, denom); e.y = __fdiv_rn(sm.y, denom);
      e.z = __fdiv_rn(sm.z, denom); e.w = __fdiv_rn(sm.w, denom);
      reinterpret_cast<float4*>(E)[o] = e;
    } else {
      e = in_a[u];
    }
    const __nv_bfloat16 b0 = __float2bfloat16_rn(e.x), b1 = __float2bfloat16_rn(e.y),
                        b2 = __float2bfloat16_rn(e.z), b3 = __float2bfloat16_rn(e.w);
    uint2 pk;
    pk.x = static_cast<uint32_t>(__bfloat16_as_ushort(b0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(b1)) << 16);
    pk.y = static_cast<uint32_t>(__bfloat16_as_ushort(b2)) | (static_cast<uint32_t>(__bfloat16_as_ushort(b3)) << 16);
    reinterpret_cast<uint2*>(E_bf16)[o] = pk;
    const float f0 = __bfloat162float(b0), f1 = __bfloat162float(b1), f2 = __bfloat162float(b2),
                f3 = __bfloat162float(b3);
    acc += static_cast<double>(e.x) * e.x + static_cast<double>(e.y) * e.y +
           static_cast<double>(e.z) * e.z + static_cast<double>(e.w) * e.w;
    accb += static_cast<double>(f0) * f0 + static_cast<double>(f1) * f1 + static_cast<double>(f2) * f2 +
            static_cast<double>(f3) * f3;
    {   // |e - bf16(e)|^2: the ACTUAL rounding error of this row (each difference is exact in fp32)
      const double d0 = e.x - f0, d1 = e.y - f1, d2 = e.z - f2, d3 = e.w - f3;
      accd += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    }
    {   // fp16 plane (the fp32-mode operand) and its norms
      const uint16_t h0 = f16_bits_flush(e.x), h1 = f16_bits_flush(e.y), h2 = f16_bits_flush(e.z),
                     h3 = f16_bits_flush(e.w);
      uint2 ph;
      ph.x = static_cast<uint32_t>(h0) | (static_cast<uint32_t>(h1) << 16);
      ph.y = static_cast<uint32_t>(h2) | (static_cast<uint32_t>(h3) << 16);
      reinterpret_cast<uint2*>(E_f16)[o] = ph;
      const double g0 = f16_bits_to_float(h0), g1 = f16_bits_to_float(h1), g2 = f16_bits_to_float(h2),
                   g3 = f16_bits_to_float(h3);
      acch += g0 * g0 + g1 * g1 + g2 * g2 + g3 * g3;
      const double d0 = e.x - g0, d1 = e.y - g1, d2 = e.z - g2, d3 = e.w - g3;
      accdh += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    }
    bad |= !(isfinite(e.x) && isfinite(e.y) && isfinite(e.z) && isfinite(e.w));
    }
  }
  acc = warp_sum(acc);
  accb = warp_sum(accb);
  accd = warp_sum(accd);
  acch = warp_sum(acch);
  accdh = warp_sum(accdh);
  bad = __any_sync(0xffffffffu, bad);
  if (lane == 0) {
    ee_half[row] = static_cast<float>(0.5 * acc);
    ee_half[K_total + row] = static_cast<float>(0.5 * accb);
    // level norms (rounded UP a hair: they feed an error bound), dead-code marker, non-finite flag
    const float kInfF = __int_as_float(0x7f800000);
    const float n0 = static_cast<float>(sqrt(acc)) * 1.0000002f, n1 = static_cast<float>(sqrt(accb)) * 1.0000002f;
    const float n3 = static_cast<float>(sqrt(accd)) * 1.0000002f;          // max_k |e_k - bf16(e_k)|
    // fp16 plane: an overflowed element makes these +inf, which the pre-pass turns into "exact path for every row"
    const float n4 = static_cast<float>(sqrt(acch)) * 1.0000002f, n5 = static_cast<float>(sqrt(accdh)) * 1.0000002f;
    int v[6];
    v[0] = (n0 == n0 && n0 < kInfF) ? __float_as_int(n0) : 0;
    v[1] = (n1 == n1 && n1 < kInfF) ? __float_as_int(n1) : 0;
    v[2] = (n3 == n3 && n3 < kInfF) ? __float_as_int(n3) : 0;
    v[3] = (n4 == n4) ? __float_as_int(n4) : 0;
    v[4] = (n5 == n5) ? __float_as_int(n5) : 0;
    // an all-zero row (a dead code after an EMA update from zeroed buffers, models/vq_vae.py:52-53,88): the LOWEST
    // such index of the level is remembered as K_per - local index (0 = none) for the de-duplication below
    v[5] = (acc == 0.0) ? K_per - (row % K_per) : 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) s_red[threadIdx.x >> 5][i] = v[i];
    if (bad || !(acc == acc) || isinf(static_cast<float>(acc)))
      level_meta[(row / K_per) * VQB200_LEVEL_META_FLOATS + 1] = 1.0f;
  }
  }
  // one atomicMax per block and word instead of one per row (K_total x 6 atomics on a handful of L2 lines);
  // positive floats order like their bit patterns.  A block whose rows straddle two levels reduces per row.
  __syncthreads();
  {
    const int wpb = blockDim.x >> 5, row0 = row_base + blockIdx.x * wpb;
    const int rows_here = K_total - row0 < wpb ? K_total - row0 : wpb;
    const int slot[6] = {0, 2, 3, 4, 5, 7};
    if (rows_here > 0 && threadIdx.x < 6) {
      const bool one_level = row0 / K_per == (row0 + rows_here - 1) / K_per;
      int m = 0;
      for (int w = 0; w < rows_here; ++w) {
        const int x = s_red[w][threadIdx.x];
        if (one_level) m = x > m ? x : m;
        else if (x) atomicMax(reinterpret_cast<int*>(level_meta + ((row0 + w) / K_per) * VQB200_LEVEL_META_FLOATS + slot[threadIdx.x]), x);
      }
      if (one_level && m)
        atomicMax(reinterpret_cast<int*>(level_meta + (row0 / K_per) * VQB200_LEVEL_META_FLOATS + slot[threadIdx.x]), m);
    }
  }
  // De-duplicate dead codes.  Every all-zero row of a level has the same distance to any latent, and arg-min
  // takes the lowest index, so all but the first can never be chosen: their |e|^2/2 becomes +inf (score -inf).
  // Without this a collapsed codebook fills every row's candidate list with ties and sends the row to the exact
  // SIMT kernel (measured: 54 % of the stage-2 training step).  Done by the last block to finish (ticket in
  // level_meta[6] of level 0, zeroed by the launcher), so it costs no extra launch.
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0)
    s_last = atomicAdd(reinterpret_cast<int*>(level_meta + 6), 1) == static_cast<int>(gridDim.x) - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // (this tail is serial: it used to walk K_total / 256 dependent L2 round trips -- 20 of the kernel's 28 us at 4096
  // codes.  Now: the levels' markers once into shared memory; no level with an all-zero code -> nothing to do; else
  // eight rows' |e|^2/2 per thread in flight together.)
  __shared__ int s_first[VQB200_MAX_LEVELS];
  __shared__ int s_any;
  const int n_levels = K_total / K_per;
  if (threadIdx.x == 0) s_any = 0;
  __syncthreads();
  for (int l = threadIdx.x; l < n_levels && l < VQB200_MAX_LEVELS; l += blockDim.x) {
    const int f = __ldcg(reinterpret_cast<const int*>(level_meta + l * VQB200_LEVEL_META_FLOATS + 7));
    s_first[l] = f;
    if (f > 0) s_any = 1;
  }
  __syncthreads();
  if (!s_any) return;
  const float kInf = __int_as_float(0x7f800000);
  constexpr int TB = 8;
  for (int r0 = threadIdx.x; r0 < K_total; r0 += blockDim.x * TB) {
    float ee[TB];
#pragma unroll
    for (int u = 0; u < TB; ++u) {
      const int r = r0 + u * blockDim.x;
      ee[u] = r < K_total ? __ldcg(ee_half + r) : 1.f;
    }
#pragma unroll
    for (int u = 0; u < TB; ++u) {
      const int r = r0 + u * blockDim.x;
      if (r >= K_total) continue;
      const int lvl = r / K_per;
      const int first = lvl < VQB200_MAX_LEVELS ? s_first[lvl]
                                                : __ldcg(reinterpret_cast<const int*>(level_meta + lvl * VQB200_LEVEL_META_FLOATS + 7));
      if (first > 0 && ee[u] == 0.f && K_per - (r % K_per) != first) {
        ee_half[r] = kInf;
        ee_half[K_total + r] = kInf;
      }
    }
  }
}

int launch_codebook_refresh(int mode, const float* seg_sum, const float* seg_cnt, float decay, float omd,
                            float eps, int K_total, int D, int K_per, float* ema_cs, float* ema_emb, float* E,
                            uint16_t* E_bf16, float* ee_half, float* level_meta, cudaStream_t s, int chain_phase) {
  const int levels = K_total / K_per;
  const int wpb = 8;
  if (mode == 1 && chain_phase == 1) {
    // level 0 receives no decay-only step and its cache entries are current: its rows are not visited and its level
    // norms stay (only the ticket word of the last-block pass, level_meta[6], is cleared)
    if (levels < 2) return VQB200_OK;
    cudaError_t e = cudaMemsetAsync(level_meta + VQB200_LEVEL_META_FLOATS, 0, sizeof(float) * VQB200_LEVEL_META_FLOATS * (levels - 1), s);
    if (e == cudaSuccess) e = cudaMemsetAsync(level_meta + 6, 0, sizeof(float), s);
    if (e != cudaSuccess) return status_of(e);
    const int blocks1 = (K_total - K_per + wpb - 1) / wpb;
    codebook_refresh_kernel<1><<<blocks1, wpb * 32, 0, s>>>(seg_sum, seg_cnt, decay, omd, eps, K_total, D, K_per, ema_cs,
                                                           ema_emb, E, E_bf16, ee_half, level_meta, 1, K_per);
    return status_of(cudaGetLastError());
  }
  cudaError_t e = cudaMemsetAsync(level_meta, 0, sizeof(float) * VQB200_LEVEL_META_FLOATS * levels, s);
  if (e != cudaSuccess) return status_of(e);
  const int blocks = (K_total + wpb - 1) / wpb;
  if (mode == 1)
    codebook_refresh_kernel<1><<<blocks, wpb * 32, 0, s>>>(seg_sum, seg_cnt, decay, omd, eps, K_total, D,
                                                          K_per, ema_cs, ema_emb, E, E_bf16, ee_half, level_meta, chain_phase, 0);
  else if (mode == 2)
    codebook_refresh_kernel<2><<<blocks, wpb * 32, 0, s>>>(seg_sum, seg_cnt, 0.f, 0.f, 0.f, K_total, D, K_per,
                                                          nullptr, nullptr, E, E_bf16, ee_half, level_meta, 0, 0);
  else
    codebook_refresh_kernel<0><<<blocks, wpb * 32, 0, s>>>(nullptr, nullptr, 0.f, 0.f, 0.f, K_total, D, K_per,
                                                          nullptr, nullptr, E, E_bf16, ee_half, level_meta, 0, 0);
  return status_of(cudaGetLastError());
}

// --------------------------------------------------------------------------------------------
// gather + straight-through + residual + commitment partial sum + histogram
// --------------------------------------------------------------------------------------------
constexpr int ROW_THREADS = 256;

__device__ __forceinline__ void block_add_double(double v, double* dst) {
  __shared__ double part[ROW_THREADS / 32];
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    double t = (threadIdx.x < ROW_THREADS / 32) ? part[threadIdx.x] : 0.0;
    t = warp_sum(t);
    if (threadIdx.x == 0) atomicAdd(dst, t);
  }
}

constexpr int GATHER_ILP = 4;   // independent 128-bit items per thread kept in flight (HBM latency x bandwidth)

constexpr int HIST_SMEM_BINS = 2048;   // small codebooks: all counts land on a few L2 lines, privatise per CTA

template <bool ACC, bool SMEM_HIST>
__global__ void __launch_bounds__(ROW_THREADS, 3)
gather_kernel(const float4* __restrict__ z, const float4* __restrict__ E, const int64_t* __restrict__ idx,
              int64_t N, int D4, int d4_shift, int K_total, float4* zq_out, float4* __restrict__ zq_st_out,
              float4* __restrict__ residual_out, double* sqerr_sum, int32_t* __restrict__ hist,
              const uint8_t* __restrict__ row_mask) {
  __shared__ int sh_hist[SMEM_HIST ? HIST_SMEM_BINS : 1];
  if (SMEM_HIST) {
    for (int k = threadIdx.x; k < K_total; k += blockDim.x) sh_hist[k] = 0;
    __syncthreads();
  }
  const int64_t total = N * D4;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  float err = 0.f;
  for (int64_t i0 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i0 < total;
       i0 += stride * GATHER_ILP) {
    int k[GATHER_ILP], row[GATHER_ILP];       // rows per call and code ids both fit 31 bits (checked by the launcher)
    int c[GATHER_ILP];
    float4 v[GATHER_ILP], e[GATHER_ILP], prev[GATHER_ILP];
    bool ok[GATHER_ILP];
#pragma unroll
    for (int u = 0; u < GATHER_ILP; ++u) {                  // issue every load of the batch first
      const int64_t i = i0 + u * stride;
      ok[u] = i < total;
      if (!ok[u]) continue;
      if (d4_shift >= 0) { row[u] = static_cast<int>(i >> d4_shift); c[u] = static_cast<int>(i & ((1 << d4_shift) - 1)); }
      else { row[u] = static_cast<int>(i / D4); c[u] = static_cast<int>(i - static_cast<int64_t>(row[u]) * D4); }
      const int64_t kk = idx[row[u]];
      k[u] = (kk >= 0 && kk < K_total) ? static_cast<int>(kk) : -1;
      v[u] = ld_stream(z + i);
      if (ACC) prev[u] = zq_out[i];
    }
#pragma unroll
    for (int u = 0; u < GATHER_ILP; ++u) {
      ok[u] = ok[u] && k[u] >= 0;                           // out-of-range ids are never produced by vqb200_search
      if (ok[u]) e[u] = __ldg(E + static_cast<int64_t>(k[u]) * D4 + c[u]);
    }
#pragma unroll
    for (int u = 0; u < GATHER_ILP; ++u) {
      if (!ok[u]) continue;
      const int64_t i = i0 + u * stride;
      float4 df;
      df.x = __fsub_rn(e[u].x, v[u].x); df.y = __fsub_rn(e[u].y, v[u].y);
      df.z = __fsub_rn(e[u].z, v[u].z); df.w = __fsub_rn(e[u].w, v[u].w);
      if (zq_out) {
        float4 o = e[u];
        if (ACC) { o.x = __fadd_rn(prev[u].x, e[u].x); o.y = __fadd_rn(prev[u].y, e[u].y);
                   o.z = __fadd_rn(prev[u].z, e[u].z); o.w = __fadd_rn(prev[u].w, e[u].w); }
        st_stream(zq_out + i, o);
      }
      if (zq_st_out) {                                      // fl(z + fl(zq - z)): bitwise != zq
        float4 o;
        o.x = __fadd_rn(v[u].x, df.x); o.y = __fadd_rn(v[u].y, df.y);
        o.z = __fadd_rn(v[u].z, df.z); o.w = __fadd_rn(v[u].w, df.w);
        st_stream(zq_st_out + i, o);
      }
      if (residual_out)                                     // fl(z - zq) (= -df exactly)
        st_stream(residual_out + i, make_float4(-df.x, -df.y, -df.z, -df.w));
      err = fmaf(df.x, df.x, err); err = fmaf(df.y, df.y, err);
      err = fmaf(df.z, df.z, err); err = fmaf(df.w, df.w, err);
      if (hist && c[u] == 0 && (!row_mask || row_mask[row[u]])) {
        if (SMEM_HIST) atomicAdd(sh_hist + k[u], 1);
        else atomicAdd(hist + k[u], 1);
      }
    }
  }
  if (SMEM_HIST) {
    __syncthreads();
    for (int k = threadIdx.x; k < K_total; k += blockDim.x) {
      const int v = sh_hist[k];
      if (v) atomicAdd(hist + k, v);
    }
  }
  if (sqerr_sum) block_add_double(static_cast<double>(err), sqerr_sum);
}

int launch_gather(const float* z, const float* E, const int64_t* idx, int64_t N, int D, int K_total, float* zq_out,
                  int zq_accumulate, float* zq_st_out, float* residual_out, double* sqerr_sum, int32_t* hist,
                  const uint8_t* row_mask, cudaStream_t s) {
  if (N == 0) return VQB200_OK;
  if (N > 0x7fffffff) return VQB200_ESHAPE;
  const int D4 = D >> 2;
  int shift = -1;
  if ((D4 & (D4 - 1)) == 0) { shift = 0; while ((1 << shift) < D4) ++shift; }
  const int64_t total = N * D4;
  int64_t blocks = (total + ROW_THREADS - 1) / ROW_THREADS;
  blocks = (blocks + GATHER_ILP - 1) / GATHER_ILP;
  const int64_t cap = static_cast<int64_t>(kNumSMs) * 3 * 4;          // 3 resident CTAs/SM, four waves
  if (blocks > cap) blocks = cap;
  auto Z = reinterpret_cast<const float4*>(z);
  auto Ev = reinterpret_cast<const float4*>(E);
  const bool sh = hist != nullptr && K_total <= HIST_SMEM_BINS;
  if (sh && blocks > static_cast<int64_t>(kNumSMs) * 3) blocks = static_cast<int64_t>(kNumSMs) * 3;   // one wave: fewer flushes
  const unsigned g = static_cast<unsigned>(blocks);
  auto ZQ = reinterpret_cast<float4*>(zq_out);
  auto ST = reinterpret_cast<float4*>(zq_st_out);
  auto RS = reinterpret_cast<float4*>(residual_out);
#define VQ_GATHER(ACC_, SH_) \
  gather_kernel<ACC_, SH_><<<g, ROW_THREADS, 0, s>>>(Z, Ev, idx, N, D4, shift, K_total, ZQ, ST, RS, sqerr_sum, hist, row_mask)
  if (zq_accumulate) { if (sh) VQ_GATHER(true, true); else VQ_GATHER(true, false); }
  else { if (sh) VQ_GATHER(false, true); else VQ_GATHER(false, false); }
#undef VQ_GATHER
  return status_of(cudaGetLastError());
}

__global__ void __launch_bounds__(ROW_THREADS)
st_loss_kernel(const float4* __restrict__ z, const float4* __restrict__ zq, int64_t n4, float4* __restrict__ st,
               double* sqerr_sum) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  float err = 0.f;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = ld_stream(z + i), e = ld_stream(zq + i);
    float4 df;
    df.x = __fsub_rn(e.x, v.x); df.y = __fsub_rn(e.y, v.y); df.z = __fsub_rn(e.z, v.z); df.w = __fsub_rn(e.w, v.w);
    if (st) {
      float4 o;
      o.x = __fadd_rn(v.x, df.x); o.y = __fadd_rn(v.y, df.y); o.z = __fadd_rn(v.z, df.z); o.w = __fadd_rn(v.w, df.w);
      st_stream(st + i, o);
    }
    err = fmaf(df.x, df.x, err); err = fmaf(df.y, df.y, err);
    err = fmaf(df.z, df.z, err); err = fmaf(df.w, df.w, err);
  }
  if (sqerr_sum) block_add_double(static_cast<double>(err), sqerr_sum);
}

static unsigned stream_grid(int64_t items) {
  int64_t blocks = (items + ROW_THREADS - 1) / ROW_THREADS;
  const int64_t cap = static_cast<int64_t>(kNumSMs) * 8 * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<unsigned>(blocks);
}

int launch_st_loss(const float* z, const float* zq, int64_t n_elems, float* st, double* sqerr_sum, cudaStream_t s) {
  if (n_elems == 0) return VQB200_OK;
  const int64_t n4 = n_elems >> 2;
  st_loss_kernel<<<stream_grid(n4), ROW_THREADS, 0, s>>>(reinterpret_cast<const float4*>(z),
                                                        reinterpret_cast<const float4*>(zq), n4,
                                                        reinterpret_cast<float4*>(st), sqerr_sum);
  return status_of(cudaGetLastError());
}

// --------------------------------------------------------------------------------------------
// residual-VQ tail in ONE pass (eval mode): z_q = ((E[i_0] + E[i_1]) + ...) summed in level order
// (models/vq_vae.py:261), z_q_st = fl(z + fl(z_q - z)) (:263), sum (z_q - z)^2 (:1293) and the usage histogram
// of every level (:266), straight from the level-major indices.  The per-level path keeps z_q in HBM and
// re-reads / re-writes it on every level (3 x 8 D bytes per latent more at four levels) and reads it once more
// for the straight-through pass; here the codebook rows come from L2 and each output is written once.
// Not usable while the EMA update mutates the codebook between levels (:251 then :248 of the next level).
// --------------------------------------------------------------------------------------------
constexpr int RVQ_MAX_LEVELS = 8;

__global__ void __launch_bounds__(ROW_THREADS)
rvq_finalize_kernel(const float4* __restrict__ z, const int64_t* __restrict__ idx, int64_t lstride, int64_t N, int D4,
                    int d4_shift, int L, const float4* __restrict__ E, int K_total, float4* __restrict__ zq_out,
                    float4* __restrict__ zq_st_out, double* sqerr_sum, int32_t* __restrict__ hist) {
  const int64_t total = N * D4;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  float err = 0.f;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    int64_t row;
    int c;
    if (d4_shift >= 0) { row = i >> d4_shift; c = static_cast<int>(i & ((1 << d4_shift) - 1)); }
    else { row = i / D4; c = static_cast<int>(i - row * D4); }
    int k[RVQ_MAX_LEVELS];
#pragma unroll
    for (int l = 0; l < RVQ_MAX_LEVELS; ++l) {
      if (l < L) {
        const int64_t kk = idx[static_cast<int64_t>(l) * lstride + row];
        k[l] = (kk >= 0 && kk < K_total) ? static_cast<int>(kk) : -1;
      }
    }
    const float4 v = ld_stream(z + i);
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int l = 0; l < RVQ_MAX_LEVELS; ++l) {
      if (l < L && k[l] >= 0) {
        const float4 e = __ldg(E + static_cast<int64_t>(k[l]) * D4 + c);
        if (l == 0) q = e;
        else { q.x = __fadd_rn(q.x, e.x); q.y = __fadd_rn(q.y, e.y); q.z = __fadd_rn(q.z, e.z); q.w = __fadd_rn(q.w, e.w); }
        if (hist && c == 0) atomicAdd(hist + k[l], 1);
      }
    }
    float4 df;
    df.x = __fsub_rn(q.x, v.x); df.y = __fsub_rn(q.y, v.y); df.z = __fsub_rn(q.z, v.z); df.w = __fsub_rn(q.w, v.w);
    if (zq_out) st_stream(zq_out + i, q);
    if (zq_st_out)
      st_stream(zq_st_out + i, make_float4(__fadd_rn(v.x, df.x), __fadd_rn(v.y, df.y), __fadd_rn(v.z, df.z),
                                           __fadd_rn(v.w, df.w)));
    err = fmaf(df.x, df.x, err); err = fmaf(df.y, df.y, err);
    err = fmaf(df.z, df.z, err); err = fmaf(df.w, df.w, err);
  }
  if (sqerr_sum) block_add_double(static_cast<double>(err), sqerr_sum);
}

int launch_rvq_finalize(const float* z, const int64_t* idx, int64_t lstride, int64_t N, int D, int L, const float* E,
                        int K_total,
                        float* zq_out, float* zq_st_out, double* sqerr_sum, int32_t* hist, cudaStream_t s) {
  if (N == 0) return VQB200_OK;
  if (L < 1 || L > RVQ_MAX_LEVELS) return VQB200_ESHAPE;
  const int D4 = D >> 2;
  int shift = -1;
  if ((D4 & (D4 - 1)) == 0) { shift = 0; while ((1 << shift) < D4) ++shift; }
  rvq_finalize_kernel<<<stream_grid(N * D4), ROW_THREADS, 0, s>>>(
      reinterpret_cast<const float4*>(z), idx, lstride, N, D4, shift, L, reinterpret_cast<const float4*>(E), K_total,
      reinterpret_cast<float4*>(zq_out), reinterpret_cast<float4*>(zq_st_out), sqerr_sum, hist);
  return status_of(cudaGetLastError());
}

// --------------------------------------------------------------------------------------------
// statistics: one CTA over the K_total-bin histogram
// --------------------------------------------------------------------------------------------
// PACKED: the statistics come from the all-reduced float64 pack [sum sq err | element count | histogram]
// (vqb200_stats_pack) instead of this rank's int32 histogram and double sum.
template <bool PACKED>
__global__ void __launch_bounds__(1024)
stats_finalize_kernel(const int32_t* __restrict__ hist_i, const double* __restrict__ packed, int K_total,
                      float count_add, const double* __restrict__ sqerr_sum, double inv_elems, float* ep_usage,
                      float* ep_cnt, float* __restrict__ stats_out) {
  __shared__ double red[32];
  __shared__ double s_total;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  auto hist = [&](int k) -> long long {
    return PACKED ? __double2ll_rn(packed[2 + k]) : static_cast<long long>(hist_i[k]);
  };
  double t = 0.0;
  for (int k = tid; k < K_total; k += blockDim.x) t += static_cast<double>(hist(k));
  t = warp_sum(t);
  if (lane == 0) red[warp] = t;
  __syncthreads();
  if (warp == 0) {
    double v = (lane < (blockDim.x >> 5)) ? red[lane] : 0.0;
    v = warp_sum(v);
    if (lane == 0) s_total = v < 1.0 ? 1.0 : v;             // total.clamp_min(1.0)
  }
  __syncthreads();
  const double total = s_total;
  double h = 0.0, dead = 0.0;
  for (int k = tid; k < K_total; k += blockDim.x) {
    const long long c = hist(k);
    if (c > 0) { const double p = static_cast<double>(c) / total; h += p * log(p); }
    else dead += 1.0;
    if (ep_usage) ep_usage[k] += static_cast<float>(c);
  }
  __syncthreads();
  h = warp_sum(h);
  dead = warp_sum(dead);
  __shared__ double red2[32];
  if (lane == 0) { red[warp] = h; red2[warp] = dead; }
  __syncthreads();
  if (warp == 0) {
    double a = (lane < (blockDim.x >> 5)) ? red[lane] : 0.0;
    double b = (lane < (blockDim.x >> 5)) ? red2[lane] : 0.0;
    a = warp_sum(a);
    b = warp_sum(b);
    if (lane == 0) {
      const bool any = b < static_cast<double>(K_total);
      stats_out[0] = any ? static_cast<float>(exp(-a)) : 0.f;
      stats_out[1] = static_cast<float>(b / static_cast<double>(K_total));
      if (PACKED) stats_out[2] = static_cast<float>(packed[0] / (packed[1] < 1.0 ? 1.0 : packed[1]));
      else stats_out[2] = sqerr_sum ? static_cast<float>(*sqerr_sum * inv_elems) : 0.f;
      if (ep_cnt) ep_cnt[0] += PACKED ? static_cast<float>(rint(packed[1] * inv_elems)) : count_add;   // PACKED: global count
    }
  }
}

// --------------------------------------------------------------------------------------------
// Global statistics over the ranks of one NVLink / NVSwitch domain in ONE kernel (no NCCL call on the step path):
// pack -> exchange over peer memory -> reduce -> finalize.  Every rank owns a symmetric buffer (the same layout on every
// GPU, mapped into every process: torch.distributed._symmetric_memory):
//     [0, 256)            flags: uint32 flag[r] = last epoch rank r's payload arrived for
//     [256, 264)          this rank's epoch counter (device-resident so that a CUDA graph can replay the kernel)
//     [512, ...)          two payload areas (epoch parity) of world x (K_total + 2) doubles, then the reduced pack
// Push model: a rank writes its payload [sum sq err | element count | histogram] into slot[rank] of EVERY rank's buffer
// (NVLink stores), fences at system scope, raises flag[rank] on every rank, waits until all of its own flags show the
// epoch, and sums the slots in rank order -- so all ranks obtain bit-identical statistics.  Two areas are enough: a
// rank can only finish epoch e after every rank has pushed epoch e, i.e. after every rank finished reading e - 1.
// --------------------------------------------------------------------------------------------
constexpr int XCH_FLAG_BYTES = 256, XCH_HDR_BYTES = 512, XCH_MAX_WORLD = 64;

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(1024)
stats_exchange_kernel(const int32_t* __restrict__ hist_i, int K_total, const double* __restrict__ sqerr_sum,
                      double n_elems, double positions_per_elem, const uint64_t* __restrict__ peer_bufs, int rank,
                      int world, unsigned long long spin_limit, float* ep_usage, float* ep_cnt,
                      float* __restrict__ stats_out) {
  __shared__ uint32_t s_epoch;
  __shared__ double red[32], red2[32];
  __shared__ double s_total;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint8_t* mine = reinterpret_cast<uint8_t*>(peer_bufs[rank]);
  if (tid == 0) {
    uint32_t* ctr = reinterpret_cast<uint32_t*>(mine + XCH_FLAG_BYTES);
    s_epoch = ++(*ctr);                                     // every rank calls in the same order: the counters agree
  }
  __syncthreads();
  const uint32_t epoch = s_epoch;
  const int slot = K_total + 2;
  const size_t area_off = XCH_HDR_BYTES + static_cast<size_t>(epoch & 1u) * world * slot * sizeof(double);
  // 1. push this rank's payload into slot[rank] of every rank's buffer
  for (int p = 0; p < world; ++p) {
    double* dst = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(peer_bufs[p]) + area_off) + static_cast<size_t>(rank) * slot;
    for (int i = tid; i < slot; i += blockDim.x)
      dst[i] = i == 0 ? (sqerr_sum ? *sqerr_sum : 0.0) : (i == 1 ? n_elems : static_cast<double>(hist_i[i - 2]));
  }
  __threadfence_system();
  __syncthreads();
  // 2. raise this rank's flag everywhere; 3. wait for every rank's payload
  if (tid < world) st_release_sys(reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(peer_bufs[tid])) + rank, epoch);
  if (tid < world) {
    const uint32_t* f = reinterpret_cast<const uint32_t*>(mine) + tid;
    unsigned long long spins = 0;
    while (static_cast<int32_t>(ld_acquire_sys(f) - epoch) < 0)
      if (spin_limit && ++spins > spin_limit) __trap();     // a rank that never arrives must fault, not hang the GPU
  }
  __syncthreads();
  // 4. reduce the slots in rank order (bit-identical on every rank)
  const double* src = reinterpret_cast<const double*>(mine + area_off);
  double* pack = reinterpret_cast<double*>(mine + XCH_HDR_BYTES + 2 * static_cast<size_t>(world) * slot * sizeof(double));
  for (int i = tid; i < slot; i += blockDim.x) {
    double s = 0.0;
    for (int p = 0; p < world; ++p) s += __ldcv(src + static_cast<size_t>(p) * slot + i);
    pack[i] = s;
  }
  __syncthreads();
  // 5. finalize (the arithmetic of stats_finalize_kernel<true>)
  double t = 0.0;
  for (int k = tid; k < K_total; k += blockDim.x) t += static_cast<double>(__double2ll_rn(pack[2 + k]));
  t = warp_sum(t);
  if (lane == 0) red[warp] = t;
  __syncthreads();
  if (warp == 0) {
    double v = (lane < (blockDim.x >> 5)) ? red[lane] : 0.0;
    v = warp_sum(v);
    if (lane == 0) s_total = v < 1.0 ? 1.0 : v;
  }
  __syncthreads();
  const double total = s_total;
  double h = 0.0, dead = 0.0;
  for (int k = tid; k < K_total; k += blockDim.x) {
    const long long c = __double2ll_rn(pack[2 + k]);
    if (c > 0) { const double pr = static_cast<double>(c) / total; h += pr * log(pr); }
    else dead += 1.0;
    if (ep_usage) ep_usage[k] += static_cast<float>(c);
  }
  __syncthreads();
  h = warp_sum(h);
  dead = warp_sum(dead);
  if (lane == 0) { red[warp] = h; red2[warp] = dead; }
  __syncthreads();
  if (warp == 0) {
    double a = (lane < (blockDim.x >> 5)) ? red[lane] : 0.0;
    double b = (lane < (blockDim.x >> 5)) ? red2[lane] : 0.0;
    a = warp_sum(a);
    b = warp_sum(b);
    if (lane == 0) {
      const bool any = b < static_cast<double>(K_total);
      stats_out[0] = any ? static_cast<float>(exp(-a)) : 0.f;
      stats_out[1] = static_cast<float>(b / static_cast<double>(K_total));
      stats_out[2] = static_cast<float>(pack[0] / (pack[1] < 1.0 ? 1.0 : pack[1]));
      if (ep_cnt) ep_cnt[0] += static_cast<float>(rint(pack[1] * positions_per_elem));
    }
  }
}

size_t stats_exchange_buffer_bytes(int K_total, int world) {
  return XCH_HDR_BYTES + (2 * static_cast<size_t>(world) + 1) * (K_total + 2) * sizeof(double);
}

int launch_stats_exchange(const int32_t* hist, int K_total, const double* sqerr_sum, double n_elems, int levels, int D,
                          const uint64_t* peer_bufs, int rank, int world, unsigned long long spin_limit, float* ep_usage,
                          float* ep_cnt, float* stats_out, cudaStream_t s) {
  if (world < 1 || world > XCH_MAX_WORLD || rank < 0 || rank >= world) return VQB200_EINVAL;
  stats_exchange_kernel<<<1, 1024, 0, s>>>(hist, K_total, sqerr_sum, n_elems,
                                           static_cast<double>(levels) / static_cast<double>(D), peer_bufs, rank, world,
                                           spin_limit, ep_usage, ep_cnt, stats_out);
  return status_of(cudaGetLastError());
}

int launch_stats_finalize(const int32_t* hist, int K_total, float count_add, const double* sqerr_sum,
                          double inv_elems, float* ep_usage, float* ep_cnt, float* stats_out, cudaStream_t s) {
  stats_finalize_kernel<false><<<1, 1024, 0, s>>>(hist, nullptr, K_total, count_add, sqerr_sum, inv_elems, ep_usage,
                                                  ep_cnt, stats_out);
  return status_of(cudaGetLastError());
}

__global__ void __launch_bounds__(256)
stats_pack_kernel(const int32_t* __restrict__ hist, int K_total, const double* __restrict__ sqerr_sum, double n_elems,
                  double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) { out[0] = sqerr_sum ? *sqerr_sum : 0.0; out[1] = n_elems; }
  if (i < K_total) out[2 + i] = static_cast<double>(hist[i]);
}

int launch_stats_pack(const int32_t* hist, int K_total, const double* sqerr_sum, double n_elems, double* out,
                      cudaStream_t s) {
  stats_pack_kernel<<<(K_total + 255) / 256, 256, 0, s>>>(hist, K_total, sqerr_sum, n_elems, out);
  return status_of(cudaGetLastError());
}

int launch_stats_finalize_packed(const double* packed, int K_total, int levels, int D, float* ep_usage, float* ep_cnt,
                                 float* stats_out, cudaStream_t s) {
  // PACKED: inv_elems carries levels / D -- the positions every reduced element stands for
  stats_finalize_kernel<true><<<1, 1024, 0, s>>>(nullptr, packed, K_total, 0.f, nullptr,
                                                 static_cast<double>(levels) / static_cast<double>(D), ep_usage, ep_cnt,
                                                 stats_out);
  return status_of(cudaGetLastError());
}

// --------------------------------------------------------------------------------------------
// EMA segment sums: warp per row, warp-aggregated counts, vector reductions into [K_total, D]
// --------------------------------------------------------------------------------------------

constexpr int SCATTER_SLICES = 4;   // float4 slices per lane held in registers: D <= 512

__global__ void __launch_bounds__(256)
scatter_add_kernel(const float4* __restrict__ z, const int64_t* __restrict__ idx,
                   const uint8_t* __restrict__ row_mask, int64_t N, int D4, int K_total, int rpw,
                   float* __restrict__ seg_sum, float* __restrict__ seg_cnt) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  // each warp owns rpw consecutive rows per step (32 for large N; fewer for small N so that every SM has
  // warps -- 8192 rows at 32 per warp would occupy 32 of the 148 SMs): lane l < rpw looks up row base+l, peers
  // with the same code are merged so the count costs one atomic per distinct code (warp-aggregated atomics).
  for (int64_t base = warp0 * rpw; base < N; base += nwarps * rpw) {
    const int64_t my_row = base + lane;
    int64_t my_k = -1;
    if (lane < rpw && my_row < N && (!row_mask || row_mask[my_row])) {
      my_k = idx[my_row];
      if (my_k < 0 || my_k >= K_total) my_k = -1;
    }
    const unsigned peers = __match_any_sync(0xffffffffu, my_k);
    if (my_k >= 0 && lane == (__ffs(peers) - 1)) atomicAdd(seg_cnt + my_k, static_cast<float>(__popc(peers)));
    if (D4 <= 32 * SCATTER_SLICES) {
      // runs of consecutive rows with the same code are summed in registers and cost ONE set of reductions:
      // early in training, and on every level of a collapsed residual codebook, most rows share a code, and
      // thousands of atomics on the same 128 addresses serialise in L2 (measured 68 us -> 36 us -> see profiles/)
      float4 run[SCATTER_SLICES];
      int64_t run_k = -1;
      auto flush = [&]() {
        if (run_k < 0) return;
        float* dst = seg_sum + run_k * (static_cast<int64_t>(D4) * 4);
#pragma unroll
        for (int j = 0; j < SCATTER_SLICES; ++j) {
          const int c = lane + 32 * j;
          if (c < D4) red_add_v4(dst + c * 4, run[j]);
        }
      };
      for (int r = 0; r < rpw; ++r) {
        const int64_t k = __shfl_sync(0xffffffffu, my_k, r);
        if (k < 0) continue;
        if (k != run_k) {
          flush();
          run_k = k;
#pragma unroll
          for (int j = 0; j < SCATTER_SLICES; ++j) run[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const float4* src = z + (base + r) * D4;
#pragma unroll
        for (int j = 0; j < SCATTER_SLICES; ++j) {
          const int c = lane + 32 * j;
          if (c < D4) {
            const float4 v = ld_stream(src + c);
            run[j].x += v.x; run[j].y += v.y; run[j].z += v.z; run[j].w += v.w;
          }
        }
      }
      flush();
    } else {
      for (int r = 0; r < rpw; ++r) {
        const int64_t k = __shfl_sync(0xffffffffu, my_k, r);
        if (k < 0) continue;
        const float4* src = z + (base + r) * D4;
        float* dst = seg_sum + k * (static_cast<int64_t>(D4) * 4);
        for (int c = lane; c < D4; c += 32) red_add_v4(dst + c * 4, ld_stream(src + c));
      }
    }
  }
}

int launch_scatter_add(const float* z, const int64_t* idx, const uint8_t* row_mask, int64_t N, int D, int K_total,
                       float* seg_sum, float* seg_cnt, cudaStream_t s) {
  if (N == 0) return VQB200_OK;
  int rpw = 32;                                            // rows per warp step: ~4K warps in flight when N allows,
  while (rpw > 8 && N / rpw < 4096) rpw >>= 1;             // but >= 8 rows so that runs of one code aggregate
  int64_t warps = (N + rpw - 1) / rpw;
  int64_t blocks = (warps + 7) / 8;
  const int64_t cap = static_cast<int64_t>(kNumSMs) * 8;
  if (blocks > cap) blocks = cap;
  scatter_add_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(reinterpret_cast<const float4*>(z), idx, row_mask,
                                                                  N, D >> 2, K_total, rpw, seg_sum, seg_cnt);
  return status_of(cudaGetLastError());
}

// --------------------------------------------------------------------------------------------
// backward
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ROW_THREADS)
commit_backward_kernel(const float4* __restrict__ g, const float* __restrict__ gc, const float4* __restrict__ z,
                       const float4* __restrict__ zq, int64_t n4, float scale, float4* __restrict__ out) {
  const float a = gc ? (*gc) * scale : 0.f;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = ld_stream(z + i), q = ld_stream(zq + i);
    float4 o = g ? ld_stream(g + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (gc) { o.x += a * (v.x - q.x); o.y += a * (v.y - q.y); o.z += a * (v.z - q.z); o.w += a * (v.w - q.w); }   // no commitment term: no 0 * inf
    st_stream(out + i, o);
  }
}

int launch_commit_backward(const float* g, const float* gc, const float* z, const float* zq, int64_t n_elems,
                           float scale, float* out, cudaStream_t s) {
  if (n_elems == 0) return VQB200_OK;
  const int64_t n4 = n_elems >> 2;
  commit_backward_kernel<<<stream_grid(n4), ROW_THREADS, 0, s>>>(
      reinterpret_cast<const float4*>(g), gc, reinterpret_cast<const float4*>(z),
      reinterpret_cast<const float4*>(zq), n4, scale, reinterpret_cast<float4*>(out));
  return status_of(cudaGetLastError());
}

// --------------------------------------------------------------------------------------------
// wire formats
// --------------------------------------------------------------------------------------------
template <typename T>
__global__ void relayout_kernel(const int64_t* __restrict__ in, int Q, int64_t B, int64_t M, T* __restrict__ out) {
  // out[b, m*Q + q] = in[q, b, m]; consecutive threads walk the OUTPUT so stores coalesce
  const int64_t total = static_cast<int64_t>(Q) * B * M;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t o = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; o < total; o += stride) {
    const int q = static_cast<int>(o % Q);
    const int64_t bm = o / Q;
    out[o] = static_cast<T>(in[static_cast<int64_t>(q) * B * M + bm]);
  }
}

int launch_relayout(const int64_t* in, int Q, int64_t B, int64_t M, void* out, int bytes, cudaStream_t s) {
  const int64_t total = static_cast<int64_t>(Q) * B * M;
  if (total == 0) return VQB200_OK;
  const unsigned grid = stream_grid(total);
  if (bytes == VQB200_IDX_I16) relayout_kernel<int16_t><<<grid, ROW_THREADS, 0, s>>>(in, Q, B, M, static_cast<int16_t*>(out));
  else if (bytes == VQB200_IDX_I32) relayout_kernel<int32_t><<<grid, ROW_THREADS, 0, s>>>(in, Q, B, M, static_cast<int32_t*>(out));
  else if (bytes == VQB200_IDX_I64) relayout_kernel<int64_t><<<grid, ROW_THREADS, 0, s>>>(in, Q, B, M, static_cast<int64_t*>(out));
  else return VQB200_EINVAL;
  return status_of(cudaGetLastError());
}

template <typename T>
__global__ void __launch_bounds__(ROW_THREADS)
indices_to_latent_kernel(const T* __restrict__ idx, int64_t n_tok, int Q, const float4* __restrict__ E, int K_total,
                         int D4, float4* __restrict__ out) {
  const int64_t total = n_tok * D4;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t tok = i / D4;
    const int c = static_cast<int>(i - tok * D4);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = 0; q < Q; ++q) {                           // level order 0..Q-1, as torch's sum(dim=2)
      const int64_t k = static_cast<int64_t>(idx[tok * Q + q]);
      if (k < 0 || k >= K_total) continue;
      const float4 e = __ldg(E + k * D4 + c);
      if (q == 0) acc = e;
      else { acc.x = __fadd_rn(acc.x, e.x); acc.y = __fadd_rn(acc.y, e.y); acc.z = __fadd_rn(acc.z, e.z); acc.w = __fadd_rn(acc.w, e.w); }
    }
    st_stream(out + i, acc);
  }
}

int launch_indices_to_latent(const void* idx, int bytes, int64_t n_tok, int Q, const float* E, int K_total, int D,
                             float* out, cudaStream_t s) {
  if (n_tok == 0) return VQB200_OK;
  const int D4 = D >> 2;
  const unsigned grid = stream_grid(n_tok * D4);
  auto Ev = reinterpret_cast<const float4*>(E);
  auto O = reinterpret_cast<float4*>(out);
  if (bytes == VQB200_IDX_I16) indices_to_latent_kernel<int16_t><<<grid, ROW_THREADS, 0, s>>>(static_cast<const int16_t*>(idx), n_tok, Q, Ev, K_total, D4, O);
  else if (bytes == VQB200_IDX_I32) indices_to_latent_kernel<int32_t><<<grid, ROW_THREADS, 0, s>>>(static_cast<const int32_t*>(idx), n_tok, Q, Ev, K_total, D4, O);
  else if (bytes == VQB200_IDX_I64) indices_to_latent_kernel<int64_t><<<grid, ROW_THREADS, 0, s>>>(static_cast<const int64_t*>(idx), n_tok, Q, Ev, K_total, D4, O);
  else return VQB200_EINVAL;
  return status_of(cudaGetLastError());
}

// --------------------------------------------------------------------------------------------
// indices -> decoder memory (models/vq_vae.py:749: memory = mem_ln(from_code(z_q)), scripts/decode_with_vqvae.py:110-130)
// from_code is linear, so from_code(sum_q E[i_q]) = sum_q (E W^T)[i_q] + b: with the projected table P = E W^T
// ([K_total, H], rebuilt only when the codebook or the weight changes) the Linear over N tokens becomes a gather of Q
// rows per token -- no GEMM, no z_q round trip -- and the LayerNorm runs on the row while it is in registers.
// One warp per token, H <= 1024 (8 float4 slices per lane).
// --------------------------------------------------------------------------------------------
constexpr int MEM_SLICES = 8;

template <typename T>
__global__ void __launch_bounds__(256)
indices_to_memory_kernel(const T* __restrict__ idx, int64_t n_tok, int Q, const float4* __restrict__ P, int K_total,
                         int H4, const float4* __restrict__ bias, const float4* __restrict__ ln_w,
                         const float4* __restrict__ ln_b, float ln_eps, float4* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const float inv_h = 1.f / static_cast<float>(H4 * 4);
  for (int64_t tok = warp0; tok < n_tok; tok += nwarps) {
    float4 acc[MEM_SLICES];
#pragma unroll
    for (int j = 0; j < MEM_SLICES; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = 0; q < Q; ++q) {                           // level order 0..Q-1
      const int64_t k = static_cast<int64_t>(idx[tok * Q + q]);
      if (k < 0 || k >= K_total) continue;
#pragma unroll
      for (int j = 0; j < MEM_SLICES; ++j) {
        const int c = lane + 32 * j;
        if (c < H4) {
          const float4 e = __ldg(P + k * H4 + c);
          acc[j].x += e.x; acc[j].y += e.y; acc[j].z += e.z; acc[j].w += e.w;
        }
      }
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < MEM_SLICES; ++j) {
      const int c = lane + 32 * j;
      if (c < H4) {
        if (bias) { const float4 b = __ldg(bias + c); acc[j].x += b.x; acc[j].y += b.y; acc[j].z += b.z; acc[j].w += b.w; }
        sum += (acc[j].x + acc[j].y) + (acc[j].z + acc[j].w);
      }
    }
    const float mean = warp_sum(sum) * inv_h;
    float var = 0.f;
#pragma unroll
    for (int j = 0; j < MEM_SLICES; ++j) {
      const int c = lane + 32 * j;
      if (c < H4) {
        const float dx = acc[j].x - mean, dy = acc[j].y - mean, dz = acc[j].z - mean, dw = acc[j].w - mean;
        var += (dx * dx + dy * dy) + (dz * dz + dw * dw);
      }
    }
    const float rstd = rsqrtf(warp_sum(var) * inv_h + ln_eps);      // biased variance, as nn.LayerNorm
#pragma unroll
    for (int j = 0; j < MEM_SLICES; ++j) {
      const int c = lane + 32 * j;
      if (c < H4) {
        float4 o = make_float4((acc[j].x - mean) * rstd, (acc[j].y - mean) * rstd, (acc[j].z - mean) * rstd, (acc[j].w - mean) * rstd);
        if (ln_w) { const float4 g = __ldg(ln_w + c); o.x *= g.x; o.y *= g.y; o.z *= g.z; o.w *= g.w; }
        if (ln_b) { const float4 b = __ldg(ln_b + c); o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w; }
        st_stream(out + tok * H4 + c, o);
      }
    }
  }
}

int launch_indices_to_memory(const void* idx, int bytes, int64_t n_tok, int Q, const float* P, int K_total, int H,
                             const float* bias, const float* ln_w, const float* ln_b, float ln_eps, float* out,
                             cudaStream_t s) {
  if (n_tok == 0) return VQB200_OK;
  if (H % 4 != 0 || H > MEM_SLICES * 128) return VQB200_ESHAPE;
  const int H4 = H >> 2;
  int64_t blocks = (n_tok + 7) / 8;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  const unsigned g = static_cast<unsigned>(blocks);
  auto Pv = reinterpret_cast<const float4*>(P);
  auto B = reinterpret_cast<const float4*>(bias);
  auto W = reinterpret_cast<const float4*>(ln_w);
  auto Lb = reinterpret_cast<const float4*>(ln_b);
  auto O = reinterpret_cast<float4*>(out);
  if (bytes == VQB200_IDX_I16) indices_to_memory_kernel<int16_t><<<g, 256, 0, s>>>(static_cast<const int16_t*>(idx), n_tok, Q, Pv, K_total, H4, B, W, Lb, ln_eps, O);
  else if (bytes == VQB200_IDX_I32) indices_to_memory_kernel<int32_t><<<g, 256, 0, s>>>(static_cast<const int32_t*>(idx), n_tok, Q, Pv, K_total, H4, B, W, Lb, ln_eps, O);
  else if (bytes == VQB200_IDX_I64) indices_to_memory_kernel<int64_t><<<g, 256, 0, s>>>(static_cast<const int64_t*>(idx), n_tok, Q, Pv, K_total, H4, B, W, Lb, ln_eps, O);
  else return VQB200_EINVAL;
  return status_of(cudaGetLastError());
}

__global__ void minloc_unpack_kernel(const uint64_t* __restrict__ p, int64_t N, int64_t* __restrict__ out) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < N; i += stride)
    out[i] = static_cast<int64_t>(p[i] & 0xffffffffull);
}

// Codebook-sharded search on the tensor path: every rank has found the EXACT winner inside its slice of the codes
// (vqb200_search on the slice: tcgen05 candidates + fp64 re-rank); to combine the winners of the ranks with one
// all-reduce(MIN) each rank packs  (orderable(fp64 score) >> 24 << 24) | global id  -- 40 bits of the exact score
// (sign, exponent, 28 mantissa bits: ties only below 4e-9 relative, then the lower id wins) and 24 bits of id.
// The score is  |e|^2 / 2 - z.e  (the distance minus the row constant |z|^2, halved), fp64 accumulation of the
// fp32 inputs, one warp per row.
__global__ void __launch_bounds__(256)
pack_exact_kernel(const float* __restrict__ z, int64_t N, int D, const float* __restrict__ E, int K_total,
                  const int64_t* __restrict__ idx, uint64_t* __restrict__ packed) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t row = warp0; row < N; row += nwarps) {
    const int64_t k = idx[row];
    double dot = 0.0, ee = 0.0;
    const bool ok = k >= 0 && k < K_total;
    if (ok) {
      for (int d = lane * 4; d < D; d += 128) {
        const float4 a = ld_stream(reinterpret_cast<const float4*>(z + row * D + d));
        const float4 b = __ldg(reinterpret_cast<const float4*>(E + k * D + d));
        dot = fma(static_cast<double>(a.x), static_cast<double>(b.x), dot); ee = fma(static_cast<double>(b.x), static_cast<double>(b.x), ee);
        dot = fma(static_cast<double>(a.y), static_cast<double>(b.y), dot); ee = fma(static_cast<double>(b.y), static_cast<double>(b.y), ee);
        dot = fma(static_cast<double>(a.z), static_cast<double>(b.z), dot); ee = fma(static_cast<double>(b.z), static_cast<double>(b.z), ee);
        dot = fma(static_cast<double>(a.w), static_cast<double>(b.w), dot); ee = fma(static_cast<double>(b.w), static_cast<double>(b.w), ee);
      }
    }
    dot = warp_sum(dot);
    ee = warp_sum(ee);
    if (lane == 0) {
      const double d = 0.5 * ee - dot;                        // smaller is nearer
      uint64_t u = static_cast<uint64_t>(__double_as_longlong(d));
      u = (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);   // order-preserving
      if (d != d || d == -__longlong_as_double(0x7ff0000000000000ll)) u = 0ull;   // NaN (or -inf: see dist_key) wins, as in torch.argmin
      packed[row] = ok ? ((u & ~0xffffffull) | static_cast<uint64_t>(k)) : ~0ull;
    }
  }
}

int launch_pack_exact(const float* z, int64_t N, int D, const float* E, int K_total, const int64_t* idx, uint64_t* packed,
                      cudaStream_t s) {
  if (N == 0) return VQB200_OK;
  int64_t blocks = (N + 7) / 8;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  pack_exact_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(z, N, D, E, K_total, idx, packed);
  return status_of(cudaGetLastError());
}

__global__ void minloc_unpack24_kernel(const uint64_t* __restrict__ p, int64_t N, int64_t* __restrict__ out) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < N; i += stride)
    out[i] = static_cast<int64_t>(p[i] & 0xffffffull);
}

int launch_minloc_unpack24(const uint64_t* p, int64_t N, int64_t* out, cudaStream_t s) {
  if (N == 0) return VQB200_OK;
  minloc_unpack24_kernel<<<stream_grid(N), ROW_THREADS, 0, s>>>(p, N, out);
  return status_of(cudaGetLastError());
}

int launch_minloc_unpack(const uint64_t* p, int64_t N, int64_t* out, cudaStream_t s) {
  if (N == 0) return VQB200_OK;
  minloc_unpack_kernel<<<stream_grid(N), ROW_THREADS, 0, s>>>(p, N, out);
  return status_of(cudaGetLastError());
}

// --------------------------------------------------------------------------------------------
// soft assignment (models/vq_vae.py:838-843): z_soft = softmax_k(-|z - e_k|^2 / tau) @ E, one pass, online
// softmax -- the reference materialises the [N, K, D] difference tensor, the [N, K] logits and a GEMM.
// One warp per row; the row and its accumulator live in registers (float4 slices, lane-strided); codes are
// visited four at a time so that the four warp reductions overlap.  SIMT fp32: this is a training-time,
// single-level, small-N path (N = batch x tokens), not GEMM-sized work.
// --------------------------------------------------------------------------------------------
constexpr int SOFT_SLICES = 4;   // float4 slices per lane: D <= 512

__global__ void __launch_bounds__(256)
soft_assign_kernel(const float4* __restrict__ z, const float4* __restrict__ E, int64_t N, int K, int D4,
                   float inv_tau, float4* __restrict__ z_soft) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const float kNegInf = __int_as_float(0xff800000);
  for (int64_t row = warp0; row < N; row += nwarps) {
    float4 zr[SOFT_SLICES], acc[SOFT_SLICES];
#pragma unroll
    for (int j = 0; j < SOFT_SLICES; ++j) {
      const int c = lane + 32 * j;
      zr[j] = c < D4 ? z[row * D4 + c] : make_float4(0.f, 0.f, 0.f, 0.f);
      acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float m = kNegInf, ssum = 0.f;
    for (int k0 = 0; k0 < K; k0 += 4) {
      float4 ev[4][SOFT_SLICES];
      float part[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        part[u] = 0.f;
        const bool live = k0 + u < K;
#pragma unroll
        for (int j = 0; j < SOFT_SLICES; ++j) {
          const int c = lane + 32 * j;
          ev[u][j] = (live && c < D4) ? __ldg(E + static_cast<int64_t>(k0 + u) * D4 + c) : zr[j];   // zr: difference 0
          const float dx = zr[j].x - ev[u][j].x, dy = zr[j].y - ev[u][j].y, dz = zr[j].z - ev[u][j].z,
                      dw = zr[j].w - ev[u][j].w;
          part[u] += dx * dx + dy * dy + dz * dz + dw * dw;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int u = 0; u < 4; ++u) part[u] += __shfl_xor_sync(0xffffffffu, part[u], o);
      }
      float lg[4], mx = m;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        lg[u] = (k0 + u < K) ? -part[u] * inv_tau : kNegInf;
        mx = fmaxf(mx, lg[u]);
      }
      if (mx == kNegInf) continue;                        // only NaN-free -inf rows get here: nothing to add yet
      const float rescale = __expf(m - mx);                // m = -inf on the first group: exp(-inf) = 0
      ssum *= rescale;
      float w[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { w[u] = expf(lg[u] - mx); ssum += w[u]; }
#pragma unroll
      for (int j = 0; j < SOFT_SLICES; ++j) {
        float4 a = acc[j];
        a.x *= rescale; a.y *= rescale; a.z *= rescale; a.w *= rescale;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          a.x = fmaf(w[u], ev[u][j].x, a.x); a.y = fmaf(w[u], ev[u][j].y, a.y);
          a.z = fmaf(w[u], ev[u][j].z, a.z); a.w = fmaf(w[u], ev[u][j].w, a.w);
        }
        acc[j] = a;
      }
      m = mx;
    }
    const float inv = 1.f / ssum;
#pragma unroll
    for (int j = 0; j < SOFT_SLICES; ++j) {
      const int c = lane + 32 * j;
      if (c < D4) z_soft[row * D4 + c] = make_float4(acc[j].x * inv, acc[j].y * inv, acc[j].z * inv, acc[j].w * inv);
    }
  }
}

int launch_soft_assign(const float* z, int64_t N, int D, const float* E, int K, float tau, float* z_soft,
                       cudaStream_t s) {
  if (N == 0) return VQB200_OK;
  if (D > 128 * SOFT_SLICES) return VQB200_ESHAPE;
  const float t = tau > 1e-8f ? tau : 1e-8f;             // max(1e-8, tau), models/vq_vae.py:840
  int64_t blocks = (N + 7) / 8;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  soft_assign_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(reinterpret_cast<const float4*>(z),
                                                                  reinterpret_cast<const float4*>(E), N, K, D >> 2,
                                                                  1.0f / t, reinterpret_cast<float4*>(z_soft));
  return status_of(cudaGetLastError());
}

// --------------------------------------------------------------------------------------------
// usage-entropy regulariser (models/vq_vae.py:1298-1309): p_code = mean_n softmax_k(z_n . e_k).
// Forward: one warp per row, two sweeps over the codes (online max / sum, then the probabilities, added to
// p_sum[k] with one atomic per (row, code)); the per-row (max, sum) is kept for the backward.
// Backward: given g_k = dLoss/dp_code[k],  dLoss/dlogit_nj = P_nj (g_j - sum_k P_nk g_k) / N  and
// dLoss/dz_n = sum_j dLoss/dlogit_nj e_j: two more sweeps per row.  The reference materialises [N, K] logits
// and probabilities and runs two GEMMs.  SIMT fp32: a training-time, small-N path like soft_assign.
// --------------------------------------------------------------------------------------------
template <typename F>
__device__ __forceinline__ void usage_sweep(const float4 (&zr)[SOFT_SLICES], const float4* __restrict__ E, int K, int D4,
                                            int lane, F&& f) {
  for (int k0 = 0; k0 < K; k0 += 4) {
    float4 ev[4][SOFT_SLICES];
    float part[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      part[u] = 0.f;
      const bool live = k0 + u < K;
#pragma unroll
      for (int j = 0; j < SOFT_SLICES; ++j) {
        const int c = lane + 32 * j;
        ev[u][j] = (live && c < D4) ? __ldg(E + static_cast<int64_t>(k0 + u) * D4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        part[u] += zr[j].x * ev[u][j].x + zr[j].y * ev[u][j].y + zr[j].z * ev[u][j].z + zr[j].w * ev[u][j].w;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int u = 0; u < 4; ++u) part[u] += __shfl_xor_sync(0xffffffffu, part[u], o);
    }
    f(k0, part, ev);
  }
}

__global__ void __launch_bounds__(256)
usage_probs_kernel(const float4* __restrict__ z, const float4* __restrict__ E, int64_t N, int K, int D4,
                   float* __restrict__ p_sum, float2* __restrict__ row_stats) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const float kNegInf = __int_as_float(0xff800000);
  for (int64_t row = warp0; row < N; row += nwarps) {
    float4 zr[SOFT_SLICES];
#pragma unroll
    for (int j = 0; j < SOFT_SLICES; ++j) {
      const int c = lane + 32 * j;
      zr[j] = c < D4 ? z[row * D4 + c] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float m = kNegInf, ssum = 0.f;
    usage_sweep(zr, E, K, D4, lane, [&](int k0, const float (&lg)[4], const float4 (&)[4][SOFT_SLICES]) {
      float mx = m;
#pragma unroll
      for (int u = 0; u < 4; ++u) if (k0 + u < K) mx = fmaxf(mx, lg[u]);
      ssum *= expf(m - mx);
#pragma unroll
      for (int u = 0; u < 4; ++u) if (k0 + u < K) ssum += expf(lg[u] - mx);
      m = mx;
    });
    const float inv = 1.f / ssum;
    usage_sweep(zr, E, K, D4, lane, [&](int k0, const float (&lg)[4], const float4 (&)[4][SOFT_SLICES]) {
      if (lane < 4 && k0 + lane < K) atomicAdd(p_sum + k0 + lane, expf(lg[lane] - m) * inv);
    });
    if (lane == 0) row_stats[row] = make_float2(m, inv);
  }
}

__global__ void __launch_bounds__(256)
usage_probs_backward_kernel(const float4* __restrict__ z, const float4* __restrict__ E, int64_t N, int K, int D4,
                            const float2* __restrict__ row_stats, const float* __restrict__ g, float scale,
                            float4* __restrict__ grad_z) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t row = warp0; row < N; row += nwarps) {
    float4 zr[SOFT_SLICES], acc[SOFT_SLICES];
#pragma unroll
    for (int j = 0; j < SOFT_SLICES; ++j) {
      const int c = lane + 32 * j;
      zr[j] = c < D4 ? z[row * D4 + c] : make_float4(0.f, 0.f, 0.f, 0.f);
      acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float2 st = row_stats[row];
    float t = 0.f;                                        // sum_k P_nk g_k
    usage_sweep(zr, E, K, D4, lane, [&](int k0, const float (&lg)[4], const float4 (&)[4][SOFT_SLICES]) {
#pragma unroll
      for (int u = 0; u < 4; ++u) if (k0 + u < K) t = fmaf(expf(lg[u] - st.x) * st.y, g[k0 + u], t);
    });
    usage_sweep(zr, E, K, D4, lane, [&](int k0, const float (&lg)[4], const float4 (&ev)[4][SOFT_SLICES]) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (k0 + u >= K) continue;
        const float w = expf(lg[u] - st.x) * st.y * (g[k0 + u] - t);
#pragma unroll
        for (int j = 0; j < SOFT_SLICES; ++j) {
          acc[j].x = fmaf(w, ev[u][j].x, acc[j].x); acc[j].y = fmaf(w, ev[u][j].y, acc[j].y);
          acc[j].z = fmaf(w, ev[u][j].z, acc[j].z); acc[j].w = fmaf(w, ev[u][j].w, acc[j].w);
        }
      }
    });
#pragma unroll
    for (int j = 0; j < SOFT_SLICES; ++j) {
      const int c = lane + 32 * j;
      if (c < D4)
        grad_z[row * D4 + c] = make_float4(acc[j].x * scale, acc[j].y * scale, acc[j].z * scale, acc[j].w * scale);
    }
  }
}

int launch_usage_probs(const float* z, int64_t N, int D, const float* E, int K, float* p_sum, float* row_stats,
                       cudaStream_t s) {
  if (N == 0) return VQB200_OK;
  if (D > 128 * SOFT_SLICES) return VQB200_ESHAPE;
  int64_t blocks = (N + 7) / 8;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  usage_probs_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(reinterpret_cast<const float4*>(z),
                                                                  reinterpret_cast<const float4*>(E), N, K, D >> 2, p_sum,
                                                                  reinterpret_cast<float2*>(row_stats));
  return status_of(cudaGetLastError());
}

int launch_usage_probs_backward(const float* z, int64_t N, int D, const float* E, int K, const float* row_stats,
                                const float* g, float scale, float* grad_z, cudaStream_t s) {
  if (N == 0) return VQB200_OK;
  if (D > 128 * SOFT_SLICES) return VQB200_ESHAPE;
  int64_t blocks = (N + 7) / 8;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  usage_probs_backward_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(
      reinterpret_cast<const float4*>(z), reinterpret_cast<const float4*>(E), N, K, D >> 2,
      reinterpret_cast<const float2*>(row_stats), g, scale, reinterpret_cast<float4*>(grad_z));
  return status_of(cudaGetLastError());
}

}  // namespace vqb
