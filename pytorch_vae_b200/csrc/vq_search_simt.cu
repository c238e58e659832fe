// Exact nearest-code search on the fp32 CUDA cores, fused with the row argmin.
//
// Role: (1) the arbiter path for rows / codebooks the tensor-core kernel hands back (non-finite
// values, candidate-list overflow), (2) every shape the tensor-core kernel does not take,
// (3) the codebook-sharded packed min-loc search.  It is a register-tiled SGEMM
// (128 rows x 128 codes per CTA, 8x8 per thread) whose epilogue keeps a running
// (distance, index) minimum per row, so the N x K matrix never leaves the SM.
//
// Distance used: d' = |e|^2/2 - z.e  (the row-constant |z|^2 is dropped: it does not change the
// argmin and only adds rounding noise).  z.e is one fp32 FMA chain over
// D in ascending order, identical for every tile position, so bit-identical code rows give
// bit-identical distances and the lowest index wins, as in torch.argmin.
#include "common.cuh"

namespace vqb {

constexpr int BM = 128, BN = 128, BK = 16, TM = 8, TN = 8;
constexpr int SIMT_THREADS = 256;
constexpr int PAD = 4;

template <bool ROUND_BF16>
__global__ void __launch_bounds__(SIMT_THREADS, 2)
search_simt_kernel(const float* __restrict__ z, const int32_t* __restrict__ row_list,
                   const int* __restrict__ n_rows_dev, int64_t n_rows, int D, const float* __restrict__ E,
                   const float* __restrict__ ee_half, int K, int codes_per_cta, int64_t idx_offset, int64_t* __restrict__ idx_out, uint64_t* __restrict__ packed_out) {
  __shared__ float zs[BK][BM + PAD];
  __shared__ float es[BK][BN + PAD];
  __shared__ float ees[BN];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  if (n_rows_dev) n_rows = *n_rows_dev;      // row-list length produced on the device (no host sync)
  // grid.x may be smaller than the number of row tiles (the hand-back path launches a bounded grid)
  for (int64_t row0 = static_cast<int64_t>(blockIdx.x) * BM; row0 < n_rows; row0 += static_cast<int64_t>(gridDim.x) * BM) {
  __syncthreads();

  // rows this thread stages into shared memory (2 float4 per tile step)
  const int ld_r = tid >> 2, ld_c = (tid & 3) * 4;
  int64_t src_row[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int64_t r = row0 + ld_r + 64 * i;
    src_row[i] = (r < n_rows) ? (row_list ? static_cast<int64_t>(row_list[r]) : r) : -1;
  }

  uint64_t best[TM];
#pragma unroll
  for (int i = 0; i < TM; ++i) best[i] = ~0ull;

  // blockIdx.y selects a code range (the hand-back path splits K over CTAs so that a handful of rows
  // does not serialise a whole codebook sweep in one CTA); results then merge through atomicMin.
  const int k_begin = codes_per_cta > 0 ? static_cast<int>(blockIdx.y) * codes_per_cta : 0;
  const int k_end = codes_per_cta > 0 ? min(K, k_begin + codes_per_cta) : K;
  for (int n0 = k_begin; n0 < k_end; n0 += BN) {
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    __syncthreads();   // the previous tile's epilogue may still be reading ees
    if (tid < BN) ees[tid] = (n0 + tid < k_end) ? ee_half[n0 + tid] : __int_as_float(0x7f800000);

    for (int k0 = 0; k0 < D; k0 += BK) {
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (src_row[i] >= 0 && k0 + ld_c < D)
          v = *reinterpret_cast<const float4*>(z + src_row[i] * D + k0 + ld_c);
        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
        int code = n0 + ld_r + 64 * i;
        if (code < k_end && k0 + ld_c < D)
          w = *reinterpret_cast<const float4*>(E + static_cast<int64_t>(code) * D + k0 + ld_c);
        if (ROUND_BF16) {
          v.x = bf16_round(v.x); v.y = bf16_round(v.y); v.z = bf16_round(v.z); v.w = bf16_round(v.w);
          w.x = bf16_round(w.x); w.y = bf16_round(w.y); w.z = bf16_round(w.z); w.w = bf16_round(w.w);
        }
        const int r = ld_r + 64 * i;
        zs[ld_c + 0][r] = v.x; zs[ld_c + 1][r] = v.y; zs[ld_c + 2][r] = v.z; zs[ld_c + 3][r] = v.w;
        es[ld_c + 0][r] = w.x; es[ld_c + 1][r] = w.y; es[ld_c + 2][r] = w.z; es[ld_c + 3][r] = w.w;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        float a[TM], b[TN];
        *reinterpret_cast<float4*>(&a[0]) = *reinterpret_cast<const float4*>(&zs[k][ty * TM]);
        *reinterpret_cast<float4*>(&a[4]) = *reinterpret_cast<const float4*>(&zs[k][ty * TM + 4]);
        *reinterpret_cast<float4*>(&b[0]) = *reinterpret_cast<const float4*>(&es[k][tx * TN]);
        *reinterpret_cast<float4*>(&b[4]) = *reinterpret_cast<const float4*>(&es[k][tx * TN + 4]);
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }

    // fused argmin epilogue: packed (key(d'), code) minima, reduced over the 16 threads of a row
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      uint64_t m = ~0ull;
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int code = n0 + tx * TN + j;
        if (code < k_end) m = umin64(m, pack_minloc(ees[tx * TN + j] - acc[i][j], static_cast<uint32_t>(code)));
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) m = umin64(m, __shfl_xor_sync(0xffffffffu, m, o));
      best[i] = umin64(best[i], m);
    }
  }

  if (tx == 0) {
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int64_t r = row0 + ty * TM + i;
      if (r >= n_rows) continue;
      const int64_t dst = row_list ? static_cast<int64_t>(row_list[r]) : r;
      if (packed_out) {
        const uint64_t v = (best[i] & 0xffffffff00000000ull) |
                           static_cast<uint64_t>((best[i] & 0xffffffffull) + idx_offset);
        if (codes_per_cta > 0) atomicMin(reinterpret_cast<unsigned long long*>(packed_out + dst), v);
        else packed_out[dst] = v;
      }
      if (idx_out) idx_out[dst] = static_cast<int64_t>(best[i] & 0xffffffffull) + idx_offset;
    }
  }
  }  // row-tile loop
}

static int launch_impl(const float* z, const int32_t* row_list, const int* n_rows_dev, int64_t n_rows, int D,
                       const float* E, const float* ee_half, int K, int codes_per_cta, int round_bf16,
                       int64_t idx_offset, int64_t* idx_out, uint64_t* packed_out, cudaStream_t s) {
  if (n_rows == 0) return VQB200_OK;
  int64_t blocks = (n_rows + BM - 1) / BM;
  // device-side row lists are short (hand-back rows): a bounded grid avoids scheduling thousands of CTAs
  // that would exit at once; CTAs loop over row tiles.
  if (n_rows_dev && blocks > 64) blocks = 64;
  if (blocks > 0x7fffffff) return VQB200_ESHAPE;
  dim3 grid(static_cast<unsigned>(blocks), codes_per_cta > 0 ? (K + codes_per_cta - 1) / codes_per_cta : 1);
  if (round_bf16)
    search_simt_kernel<true><<<grid, SIMT_THREADS, 0, s>>>(z, row_list, n_rows_dev, n_rows, D, E, ee_half, K, codes_per_cta, idx_offset,
                                                          idx_out, packed_out);
  else
    search_simt_kernel<false><<<grid, SIMT_THREADS, 0, s>>>(z, row_list, n_rows_dev, n_rows, D, E, ee_half, K, codes_per_cta, idx_offset,
                                                           idx_out, packed_out);
  return status_of(cudaGetLastError());
}

int launch_search_simt(const float* z, const int32_t* row_list, int64_t n_rows, int D, const float* E,
                       const float* ee_half, int K, int round_bf16, int64_t idx_offset, int64_t* idx_out,
                       uint64_t* packed_out, cudaStream_t s) {
  return launch_impl(z, row_list, nullptr, n_rows, D, E, ee_half, K, 0, round_bf16, idx_offset, idx_out, packed_out, s);
}

// Row list whose length lives on the device: the grid covers max_rows x code ranges, surplus CTAs exit
// at once.  packed[row] must hold ~0 on entry; it receives key(d) << 32 | (idx_offset + argmin).
int launch_search_simt_list(const float* z, const int32_t* row_list, const int* n_rows_dev, int64_t max_rows, int D,
                            const float* E, const float* ee_half, int K, int round_bf16, int64_t idx_offset,
                            uint64_t* packed, cudaStream_t s) {
  int tiles = (K + BN - 1) / BN;
  int per = (tiles + 63) / 64;                     // at most 64 code ranges
  return launch_impl(z, row_list, n_rows_dev, max_rows, D, E, ee_half, K, per * BN, round_bf16, idx_offset, nullptr,
                     packed, s);
}

}  // namespace vqb
