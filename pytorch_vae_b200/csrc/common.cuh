// Shared device helpers for libvqb200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vq_b200.h"

#ifndef __CUDA_ARCH_FEAT_SM100_ALL
#if defined(__CUDA_ARCH__)
#error "libvqb200 is written for sm_100a only: compile with -gencode arch=compute_100a,code=sm_100a"
#endif
#endif

namespace vqb {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// Order-preserving float -> uint32 key with torch.argmin's NaN rule folded in:
// NaN maps to 0, so a packed min picks the first NaN.  d' = |e|^2/2 - z.e = -inf maps to 0 as well: the row-constant
// |z|^2 is dropped here, and a row holding +-inf has |z|^2 = +inf in the reference, whose distance
// (inf - 2 (+inf)) + |e|^2 is NaN exactly where d' is -inf -- so "first NaN or -inf" IS torch.argmin's answer for
// such rows, also when the codebook holds a NaN code further up (tests/test_gpu_rvq_fused.py::hard_rows).
__device__ __forceinline__ uint32_t dist_key(float d) {
  uint32_t u = __float_as_uint(d);
  uint32_t k = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return (d != d || u == 0xff800000u) ? 0u : k;
}
__device__ __forceinline__ uint64_t pack_minloc(float d, uint32_t idx) {
  return (static_cast<uint64_t>(dist_key(d)) << 32) | idx;
}
__device__ __forceinline__ uint64_t umin64(uint64_t a, uint64_t b) { return a < b ? a : b; }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// vector reduction into global memory (sm_90+): one L2 atomic transaction per 16 bytes
__device__ __forceinline__ void red_add_v4(float* p, const float4& v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// Streaming 128-bit accesses: rows are touched once, keep them out of L1.
__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

// cudaFuncSetAttribute is per device: one-time flags are kept per device ordinal
inline int current_device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  return dev;
}
inline int status_of(cudaError_t e) { return e == cudaSuccess ? VQB200_OK : static_cast<int>(e); }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace vqb

// Internal launchers (defined in the .cu files, called from cabi.cu).
namespace vqb {
// measurement hook, see vqb200_timing_enable
// Admission margin of one row in fp32 mode.  The tensor core scores s~ = z~.e~ - |e|^2/2 with z~ = f16(z),
// e~ = f16(e) (round to nearest, subnormals flushed); the exact score is s = z.e - |e|^2/2, and
//     s~ - s = (z~ - z).e~ + z.(e~ - e)   =>   |s~ - s| <= |z - z~| |e~| + |z| |e - e~|      (Cauchy-Schwarz)
// with the ACTUAL rounding-error norms |z - z~| (this row, summed while converting) and max_k |e_k - e~_k|
// (codebook cache, level_meta[5]) -- tighter than the worst case (2u + u^2)|z||e|, u = 2^-11, by about 2.5x,
// rigorous all the same, and valid whatever the conversion did (underflow, flushed subnormals: the error is
// measured, not assumed; an overflow makes it infinite and sends the row to the exact kernel).  For the true
// arg max a and the approximate one b:  s~_a >= s~_b - (err_a + err_b), so the margin is twice the bound, plus
// the fp32 accumulation inside the tensor core (<= (D + 32) 2^-23 |z~||e~| per score, as in bf16_input mode)
// and an absolute term for the rounding of the fp32 bias.  ss = |z|^2, sse = |z - z~|^2.
__device__ __forceinline__ float admission_margin_fp32(float ss, float sse, float emax, float emax_lp, float rho_e,
                                                       int D) {
  const float nz = sqrtf(ss) * 1.0001f, ne = sqrtf(sse) * 1.0001f;
  return 2.001f * (ne * emax_lp + nz * rho_e) + 2.f * static_cast<float>(D + 32) * 1.1920929e-7f * nz * emax_lp +
         1e-6f * emax * (emax + nz) + 1e-30f;
}

// fp32 -> fp16 bits, round to nearest even, results below the normal range flushed to zero (the tensor core is
// then never fed a subnormal; the flush is part of the measured conversion error)
__device__ __forceinline__ uint16_t f16_bits_flush(float x) {
  const uint16_t h = __half_as_ushort(__float2half_rn(x));
  return (h & 0x7c00u) ? h : static_cast<uint16_t>(h & 0x8000u);
}
__device__ __forceinline__ float f16_bits_to_float(uint16_t h) { return __half2float(__ushort_as_half(h)); }
// two at a time: packed convert, then clear the halves whose exponent field is zero (subnormal or zero)
__device__ __forceinline__ uint32_t f16x2_bits_flush(float lo, float hi) {
  const __half2 h = __floats2half2_rn(lo, hi);
  const uint32_t b = *reinterpret_cast<const uint32_t*>(&h);
  const uint32_t e = b & 0x7c007c00u;
  const uint32_t keep = ((e & 0x0000ffffu) ? 0x0000ffffu : 0x00008000u) | ((e & 0xffff0000u) ? 0xffff0000u : 0x80000000u);
  return b & keep;
}
__device__ __forceinline__ float2 f16x2_bits_to_float2(uint32_t b) {
  return __half22float2(*reinterpret_cast<const __half2*>(&b));
}

// perplexity / dead ratio / mean squared error from an int32 histogram, by ONE thread block of any size that is a
// multiple of 32 (models/vq_vae.py:209-222,267-280); ep_usage [K_total] += usage, ep_cnt [1] += count_add (either NULL).
__device__ __forceinline__ void stats_finalize_block(const int32_t* hist, int K_total, float count_add, const double* sqerr_sum,
                                                     double inv_elems, float* ep_usage, float* ep_cnt, float* stats_out) {
  __shared__ double sf_red[32], sf_red2[32];
  __shared__ double sf_total;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  // up to SF_REG bins per thread stay in registers for both passes (their loads in flight together); larger codebooks
  // re-read the tail from L2
  constexpr int SF_REG = 8;
  int c_reg[SF_REG];
#pragma unroll
  for (int u = 0; u < SF_REG; ++u) {
    const int k = tid + u * static_cast<int>(blockDim.x);
    c_reg[u] = k < K_total ? __ldcg(hist + k) : 0;
  }
  double t = 0.0;
#pragma unroll
  for (int u = 0; u < SF_REG; ++u) t += static_cast<double>(c_reg[u]);
  for (int k = tid + SF_REG * static_cast<int>(blockDim.x); k < K_total; k += blockDim.x) t += static_cast<double>(__ldcg(hist + k));
  t = warp_sum(t);
  if (lane == 0) sf_red[warp] = t;
  __syncthreads();
  if (warp == 0) {
    double v = lane < nwarps ? sf_red[lane] : 0.0;
    v = warp_sum(v);
    if (lane == 0) sf_total = v < 1.0 ? 1.0 : v;             // total.clamp_min(1.0)
  }
  __syncthreads();
  const double total = sf_total;
  double h = 0.0, dead = 0.0;
  auto bin = [&](int k, int ci) {
    const long long c = ci;
    if (c > 0) { const double pr = static_cast<double>(c) / total; h += pr * log(pr); }
    else dead += 1.0;
    if (ep_usage) ep_usage[k] += static_cast<float>(c);
  };
#pragma unroll
  for (int u = 0; u < SF_REG; ++u) {
    const int k = tid + u * static_cast<int>(blockDim.x);
    if (k < K_total) bin(k, c_reg[u]);
  }
  for (int k = tid + SF_REG * static_cast<int>(blockDim.x); k < K_total; k += blockDim.x) bin(k, __ldcg(hist + k));
  __syncthreads();
  h = warp_sum(h);
  dead = warp_sum(dead);
  if (lane == 0) { sf_red[warp] = h; sf_red2[warp] = dead; }
  __syncthreads();
  if (warp == 0) {
    double a = lane < nwarps ? sf_red[lane] : 0.0;
    double b = lane < nwarps ? sf_red2[lane] : 0.0;
    a = warp_sum(a);
    b = warp_sum(b);
    if (lane == 0) {
      const bool any = b < static_cast<double>(K_total);
      stats_out[0] = any ? static_cast<float>(exp(-a)) : 0.f;
      stats_out[1] = static_cast<float>(b / static_cast<double>(K_total));
      stats_out[2] = sqerr_sum ? static_cast<float>(__ldcg(sqerr_sum) * inv_elems) : 0.f;
      if (ep_cnt) ep_cnt[0] += count_add;
    }
  }
}

void timing_mark_begin(cudaStream_t s);
void timing_mark_end(cudaStream_t s);
int launch_search_simt(const float* z, const int32_t* row_list, int64_t n_rows, int D, const float* E,
                       const float* ee_half, int K, int round_bf16, int64_t idx_offset, int64_t* idx_out,
                       uint64_t* packed_out, cudaStream_t s);
int launch_search_simt_list(const float* z, const int32_t* row_list, const int* n_rows_dev, int64_t max_rows, int D,
                            const float* E, const float* ee_half, int K, int round_bf16, int64_t idx_offset,
                            uint64_t* packed, cudaStream_t s);
bool tc_supported(int64_t N, int K, int D);
size_t tc_workspace_bytes(int64_t N, int K, int D);
int tc_launches(int64_t N, int K, int D);
bool tc_side_pipeline(int64_t N, int K, int D);   // several chunks on the CTA-pair kernel with side jobs (vq_search_tc.cu)
// optional gather stage appended to each chunk of the tensor-core search pipeline
struct GatherArgs {
  const float* E_full;      // [K_total, D] (idx_out holds GLOBAL ids)
  int K_total;
  float* zq_out;
  float* zq_st_out;
  double* sqerr_sum;
  int32_t* hist;
  const uint8_t* row_mask;
};
// optional: the 16-bit copy and the admission margins of the rows were already produced (vqb200_residual_prep)
struct PrepArgs {
  const uint16_t* z16;      // [N, D]
  const float* margin;      // [N]
};
int launch_residual_prep(const float* z, const float* E_full, const int64_t* idx, int64_t N, int D, int K_total,
                         int mode, const float* next_level_meta, float* residual_out, uint16_t* z16_out,
                         float* margin_out, cudaStream_t s);
int launch_search_tc(const float* z, int64_t N, int D, const float* E, const uint16_t* E_bf16, const float* ee_half,
                     const float* ee_half_bf16, const float* level_meta, int K, int mode, int64_t idx_offset,
                     int64_t* idx_out, void* workspace, size_t workspace_bytes, cudaStream_t s,
                     const GatherArgs* ga = nullptr, const PrepArgs* prep = nullptr);
// optional statistics tail of the persistent kernel (the arguments of vqb200_stats_finalize)
struct RvqStatsTail {
  float* stats_out;
  float* ep_usage;
  float* ep_cnt;
  float count_add;
  double inv_elems;
};
// persistent residual-VQ forward, all levels in one kernel (vq_rvq_fused.cuh; host side vq_rvq_fused.cu)
bool rvq_fused_supported(int64_t N, int K_per, int D, int L);
size_t rvq_fused_workspace_bytes(int64_t N, int D);
int launch_rvq_fused(const float* z, int64_t N, int D, const float* E, const uint16_t* E_lp, const float* ee_half,
                     const float* level_meta, int K_per, int L, int mode, int64_t* idx_out, float* zq_out,
                     float* zq_st_out, double* sqerr_sum, int32_t* hist, void* workspace, size_t workspace_bytes,
                     cudaStream_t s, float* seg_sum = nullptr, float* seg_cnt = nullptr, const RvqStatsTail* tail = nullptr);
// training mode: refresh phase 1 -> the same kernel reducing the EMA segment sums -> refresh phase 2 (three launches)
bool rvq_fused_train_supported(int64_t N, int K_per, int D, int L);
size_t rvq_fused_train_workspace_bytes(int64_t N, int K_per, int D, int L);
int launch_rvq_train_begin(const float* z, int64_t N, int D, float* E, uint16_t* E_lp_planes, float* ee_half,
                           float* level_meta, int K_per, int L, int mode, float decay, float omd, float eps,
                           float* ema_cs, float* ema_emb, int64_t* idx_out, float* zq_out, float* zq_st_out,
                           double* sqerr_sum, int32_t* hist, float* seg_sum, float* seg_cnt, void* workspace,
                           size_t workspace_bytes, cudaStream_t s, const RvqStatsTail* tail = nullptr);
int launch_rvq_train_finish(const float* seg_sum, const float* seg_cnt, float decay, float omd, float eps, int K_per, int L,
                            int D, float* ema_cs, float* ema_emb, float* E, uint16_t* E_lp_planes, float* ee_half,
                            float* level_meta, cudaStream_t s);
int launch_rvq_fused_train(const float* z, int64_t N, int D, float* E, uint16_t* E_lp_planes, float* ee_half,
                           float* level_meta, int K_per, int L, int mode, float decay, float omd, float eps,
                           float* ema_cs, float* ema_emb, int64_t* idx_out, float* zq_out, float* zq_st_out,
                           double* sqerr_sum, int32_t* hist, void* workspace, size_t workspace_bytes, cudaStream_t s,
                           const RvqStatsTail* tail = nullptr);
bool fused_supported(int64_t N, int K, int D);
size_t fused_workspace_bytes(int64_t N);
int launch_quantize_fused(const float* z, int64_t N, int D, const float* E, const uint16_t* E_bf16, const float* ee_half,
                          const float* ee_half_bf16, const float* level_meta, int K, int mode, int64_t idx_offset,
                          int64_t* idx_out, float* zq_out, float* zq_st_out, double* sqerr_sum, int32_t* hist,
                          const uint8_t* row_mask, void* workspace, size_t workspace_bytes, cudaStream_t s);
// chain_phase (mode 1, residual codebooks): 0 = one EMA update of every code from seg_sum / seg_cnt; 1 = the decay-only
// updates a level's codes receive from the EARLIER levels of a training forward (level l: l of them; level 0 untouched);
// 2 = every level's own update from its segment sums followed by the decay-only updates of the LATER levels
int launch_codebook_refresh(int mode, const float* seg_sum, const float* seg_cnt, float decay, float omd,
                            float eps, int K_total, int D, int K_per, float* ema_cs, float* ema_emb, float* E,
                            uint16_t* E_bf16, float* ee_half, float* level_meta, cudaStream_t s, int chain_phase = 0);
int launch_gather(const float* z, const float* E, const int64_t* idx, int64_t N, int D, int K_total, float* zq_out,
                  int zq_accumulate, float* zq_st_out, float* residual_out, double* sqerr_sum, int32_t* hist,
                  const uint8_t* row_mask, cudaStream_t s);
int launch_st_loss(const float* z, const float* zq, int64_t n_elems, float* st, double* sqerr_sum, cudaStream_t s);
int launch_stats_finalize(const int32_t* hist, int K_total, float count_add, const double* sqerr_sum,
                          double inv_elems, float* ep_usage, float* ep_cnt, float* stats_out, cudaStream_t s);
size_t stats_exchange_buffer_bytes(int K_total, int world);
int launch_stats_exchange(const int32_t* hist, int K_total, const double* sqerr_sum, double n_elems, int levels, int D,
                          const uint64_t* peer_bufs, int rank, int world, unsigned long long spin_limit, float* ep_usage,
                          float* ep_cnt, float* stats_out, cudaStream_t s);
int launch_stats_pack(const int32_t* hist, int K_total, const double* sqerr_sum, double n_elems, double* out,
                      cudaStream_t s);
int launch_stats_finalize_packed(const double* packed, int K_total, int levels, int D, float* ep_usage, float* ep_cnt,
                                 float* stats_out, cudaStream_t s);
int launch_scatter_add(const float* z, const int64_t* idx, const uint8_t* row_mask, int64_t N, int D, int K_total,
                       float* seg_sum, float* seg_cnt, cudaStream_t s);
int launch_commit_backward(const float* g, const float* gc, const float* z, const float* zq, int64_t n_elems,
                           float scale, float* out, cudaStream_t s);
int launch_relayout(const int64_t* in, int Q, int64_t B, int64_t M, void* out, int bytes, cudaStream_t s);
int launch_rvq_finalize(const float* z, const int64_t* idx, int64_t lstride, int64_t N, int D, int L, const float* E,
                        int K_total,
                        float* zq_out, float* zq_st_out, double* sqerr_sum, int32_t* hist, cudaStream_t s);
int launch_usage_probs(const float* z, int64_t N, int D, const float* E, int K, float* p_sum, float* row_stats,
                       cudaStream_t s);
int launch_usage_probs_backward(const float* z, int64_t N, int D, const float* E, int K, const float* row_stats,
                                const float* g, float scale, float* grad_z, cudaStream_t s);
int launch_soft_assign(const float* z, int64_t N, int D, const float* E, int K, float tau, float* z_soft,
                       cudaStream_t s);
int launch_indices_to_latent(const void* idx, int bytes, int64_t n_tok, int Q, const float* E, int K_total, int D,
                             float* out, cudaStream_t s);
// tiled row softmax fused with its logits' contraction (vq_softmax.cu)
size_t softmax_rows_workspace_bytes(int64_t N, int K);
int launch_softmax_rows(const float* z, int64_t N, int D, const float* E, const float* beta, int K, float alpha,
                        float* row_stats, float* P_out, float* p_sum, void* workspace, size_t workspace_bytes,
                        cudaStream_t s);
int launch_indices_to_memory(const void* idx, int bytes, int64_t n_tok, int Q, const float* P, int K_total, int H,
                             const float* bias, const float* ln_w, const float* ln_b, float ln_eps, float* out,
                             cudaStream_t s);
int launch_minloc_unpack(const uint64_t* p, int64_t N, int64_t* out, cudaStream_t s);
int launch_pack_exact(const float* z, int64_t N, int D, const float* E, int K_total, const int64_t* idx, uint64_t* packed,
                      cudaStream_t s);
int launch_minloc_unpack24(const uint64_t* p, int64_t N, int64_t* out, cudaStream_t s);
}
