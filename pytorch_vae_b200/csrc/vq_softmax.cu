// Row softmax over the codes, fused with the contraction that produces its logits -- the shared first half of the
// soft-VQ assignment (models/vq_vae.py:838-843: softmax_k(-|z - e_k|^2 / tau)) and of the usage-entropy regulariser
// (:1298-1309: softmax_k(z . e_k)):
//     logit[n, k] = alpha * (z_n . e_k) + beta[k]
// (soft-VQ: alpha = 2 / tau, beta[k] = -|e_k|^2 / tau -- the row constant -|z_n|^2 / tau cancels in the softmax).
// The reference materialises [N, K, D] differences / [N, K] logits and softmaxes them with several ATen passes; the
// first version here (soft_assign_kernel / usage_probs_kernel, vq_rowops.cu: one warp per row, a warp reduction per
// code) ran at 2-17 % of the fp32 FMA peak.  This is a register-tiled fp32 contraction (128 rows x 128 codes per CTA,
// 8 x 8 per thread: the mainloop of search_simt_kernel) in two sweeps:
//   stats : every thread keeps a private online (max, sum exp) over the codes it sees; the 16 threads of a row merge
//           them once at the end; the codes may be split over CTAs (small N), partial results to [N, splits, 2]
//   emit  : one 128 x 128 tile per CTA: merges the row's partial statistics, p = exp(logit - max) / sum, writes the
//           probabilities P [N, K] and / or adds their column sums to p_sum [K]
// The second contraction of either path (P @ E, dS @ E) is a plain GEMM and is left to the library (ops.py).
// fp32 throughout: the results agree with the reference's fp32 to summation order.
#include "common.cuh"

namespace vqb {

constexpr int SX_BM = 128, SX_BN = 128, SX_BK = 16, SX_THREADS = 256, SX_PAD = 4;

struct SxTiles {
  float zs[SX_BK][SX_BM + SX_PAD];
  float es[SX_BK][SX_BN + SX_PAD];
};

// acc[i][j] = z[row0 + ty * 8 + i] . E[n0 + tx * 8 + j]  (rows / codes out of range contribute zeros)
__device__ __forceinline__ void sx_tile_dots(const float* __restrict__ z, int64_t n_rows, int D, const float* __restrict__ E,
                                             int K, int64_t row0, int n0, SxTiles& sh, float (&acc)[8][8]) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int ld_r = tid >> 2, ld_c = (tid & 3) * 4;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < D; k0 += SX_BK) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = ld_r + 64 * i;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f), w = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row0 + r < n_rows && k0 + ld_c < D) v = *reinterpret_cast<const float4*>(z + (row0 + r) * D + k0 + ld_c);
      if (n0 + r < K && k0 + ld_c < D) w = __ldg(reinterpret_cast<const float4*>(E + static_cast<int64_t>(n0 + r) * D + k0 + ld_c));
      sh.zs[ld_c + 0][r] = v.x; sh.zs[ld_c + 1][r] = v.y; sh.zs[ld_c + 2][r] = v.z; sh.zs[ld_c + 3][r] = v.w;
      sh.es[ld_c + 0][r] = w.x; sh.es[ld_c + 1][r] = w.y; sh.es[ld_c + 2][r] = w.z; sh.es[ld_c + 3][r] = w.w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SX_BK; ++k) {
      float a[8], b[8];
      *reinterpret_cast<float4*>(&a[0]) = *reinterpret_cast<const float4*>(&sh.zs[k][ty * 8]);
      *reinterpret_cast<float4*>(&a[4]) = *reinterpret_cast<const float4*>(&sh.zs[k][ty * 8 + 4]);
      *reinterpret_cast<float4*>(&b[0]) = *reinterpret_cast<const float4*>(&sh.es[k][tx * 8]);
      *reinterpret_cast<float4*>(&b[4]) = *reinterpret_cast<const float4*>(&sh.es[k][tx * 8 + 4]);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
}

// merge of two online-softmax states (m, l): l counts exp(logit - m)
__device__ __forceinline__ void sx_merge(float& m, float& l, float m2, float l2) {
  const float mm = fmaxf(m, m2);
  // exp(-inf - (-inf)) would be NaN: an empty state contributes nothing
  const float a = (m == mm) ? l : (l == 0.f ? 0.f : l * expf(m - mm));
  const float b = (m2 == mm) ? l2 : (l2 == 0.f ? 0.f : l2 * expf(m2 - mm));
  m = mm;
  l = a + b;
}

// ---- sweep 1: per-row (max, sum exp) over the codes [blockIdx.y * codes_per_cta, ...) -> part [N][splits][2]
__global__ void __launch_bounds__(SX_THREADS, 2)
softmax_stats_kernel(const float* __restrict__ z, int64_t n_rows, int D, const float* __restrict__ E,
                     const float* __restrict__ beta, int K, float alpha, int codes_per_cta, int splits,
                     float* __restrict__ part) {
  __shared__ SxTiles sh;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t row0 = static_cast<int64_t>(blockIdx.x) * SX_BM;
  const int k_begin = static_cast<int>(blockIdx.y) * codes_per_cta;
  const int k_end = min(K, k_begin + codes_per_cta);
  const float kNegInf = __int_as_float(0xff800000);
  float m[8], l[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { m[i] = kNegInf; l[i] = 0.f; }
  for (int n0 = k_begin; n0 < k_end; n0 += SX_BN) {
    float acc[8][8];
    sx_tile_dots(z, n_rows, D, E, K, row0, n0, sh, acc);
    float bj[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = n0 + tx * 8 + j;
      bj[j] = c < k_end ? (beta ? __ldg(beta + c) : 0.f) : kNegInf;     // codes out of range: logit -inf
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float s[8], tm = kNegInf;
#pragma unroll
      for (int j = 0; j < 8; ++j) { s[j] = fmaf(alpha, acc[i][j], bj[j]); tm = fmaxf(tm, s[j]); }
      if (tm > m[i]) {                                      // rescale the running sum to the new maximum
        l[i] = l[i] == 0.f ? 0.f : l[i] * expf(m[i] - tm);
        m[i] = tm;
      }
      if (m[i] > kNegInf) {
#pragma unroll
        for (int j = 0; j < 8; ++j) l[i] += expf(s[j] - m[i]);
      }
    }
  }
  // the 16 threads of a row (tx = 0..15: lanes differing in the low four bits) merge once
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m[i], o), l2 = __shfl_xor_sync(0xffffffffu, l[i], o);
      sx_merge(m[i], l[i], m2, l2);
    }
    const int64_t row = row0 + ty * 8 + i;
    if (tx == 0 && row < n_rows) {
      float* dst = part + (row * splits + blockIdx.y) * 2;
      dst[0] = m[i];
      dst[1] = l[i];
    }
  }
}

// ---- sweep 2: one tile per CTA: probabilities out and / or their column sums
__global__ void __launch_bounds__(SX_THREADS, 2)
softmax_emit_kernel(const float* __restrict__ z, int64_t n_rows, int D, const float* __restrict__ E,
                    const float* __restrict__ beta, int K, float alpha, const float* __restrict__ part, int splits,
                    float* __restrict__ row_stats, float* __restrict__ P_out, float* __restrict__ p_sum) {
  __shared__ SxTiles sh;
  __shared__ float colsum[SX_BN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t row0 = static_cast<int64_t>(blockIdx.x) * SX_BM;
  const int n0 = static_cast<int>(blockIdx.y) * SX_BN;
  const float kNegInf = __int_as_float(0xff800000);
  if (p_sum && tid < SX_BN) colsum[tid] = 0.f;
  float acc[8][8];
  sx_tile_dots(z, n_rows, D, E, K, row0, n0, sh, acc);     // (its barriers also order the colsum initialisation)
  float bj[8], cs[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = n0 + tx * 8 + j;
    bj[j] = c < K ? (beta ? __ldg(beta + c) : 0.f) : kNegInf;
    cs[j] = 0.f;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t row = row0 + ty * 8 + i;
    if (row >= n_rows) continue;
    float m = kNegInf, l = 0.f;
    for (int sp = 0; sp < splits; ++sp) {
      const float2 q = *reinterpret_cast<const float2*>(part + (row * splits + sp) * 2);
      sx_merge(m, l, q.x, q.y);
    }
    const float inv = 1.f / l;
    if (row_stats && blockIdx.y == 0 && tx == 0) { row_stats[row * 2] = m; row_stats[row * 2 + 1] = inv; }
    float pr[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      pr[j] = expf(fmaf(alpha, acc[i][j], bj[j]) - m) * inv;
      cs[j] += pr[j];
    }
    if (P_out) {
      float* dst = P_out + row * K + n0 + tx * 8;
      if ((K & 3) == 0 && n0 + tx * 8 + 8 <= K) {
        *reinterpret_cast<float4*>(dst) = make_float4(pr[0], pr[1], pr[2], pr[3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(pr[4], pr[5], pr[6], pr[7]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (n0 + tx * 8 + j < K) dst[j] = pr[j];
      }
    }
  }
  if (p_sum) {
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&colsum[tx * 8 + j], cs[j]);
    __syncthreads();
    if (tid < SX_BN && n0 + tid < K) atomicAdd(p_sum + n0 + tid, colsum[tid]);
  }
}

int softmax_rows_splits(int64_t N, int K) {
  const int64_t row_tiles = (N + SX_BM - 1) / SX_BM;
  const int code_tiles = (K + SX_BN - 1) / SX_BN;
  int64_t want = (2 * kNumSMs * 2 + row_tiles - 1) / (row_tiles > 0 ? row_tiles : 1);   // ~2 waves of 2 CTAs per SM
  if (want < 1) want = 1;
  int splits = static_cast<int>(want < code_tiles ? want : code_tiles);
  return splits < 1 ? 1 : splits;
}

size_t softmax_rows_workspace_bytes(int64_t N, int K) {
  return static_cast<size_t>(N > 0 ? N : 0) * softmax_rows_splits(N, K) * 2 * sizeof(float);
}

int launch_softmax_rows(const float* z, int64_t N, int D, const float* E, const float* beta, int K, float alpha,
                        float* row_stats, float* P_out, float* p_sum, void* workspace, size_t workspace_bytes,
                        cudaStream_t s) {
  if (N == 0) return VQB200_OK;
  if (workspace_bytes < softmax_rows_workspace_bytes(N, K)) return VQB200_EWORKSPACE;
  const int64_t row_tiles = (N + SX_BM - 1) / SX_BM;
  const int code_tiles = (K + SX_BN - 1) / SX_BN;
  if (row_tiles > 0x7fffffff) return VQB200_ESHAPE;
  const int splits = softmax_rows_splits(N, K);
  const int tiles_per = (code_tiles + splits - 1) / splits;
  const int used = (code_tiles + tiles_per - 1) / tiles_per;         // splits that own at least one tile
  float* part = static_cast<float*>(workspace);
  softmax_stats_kernel<<<dim3(static_cast<unsigned>(row_tiles), used), SX_THREADS, 0, s>>>(z, N, D, E, beta, K, alpha,
                                                                                          tiles_per * SX_BN, used, part);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return status_of(e);
  softmax_emit_kernel<<<dim3(static_cast<unsigned>(row_tiles), code_tiles), SX_THREADS, 0, s>>>(z, N, D, E, beta, K, alpha, part,
                                                                                              used, row_stats, P_out, p_sum);
  return status_of(cudaGetLastError());
}

}  // namespace vqb
