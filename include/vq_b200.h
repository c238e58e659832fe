/* vq_b200.h -- C ABI of libvqb200.so: the B200 (sm_100a) vector-quantizer hot path.
 *
 * The reference (jluuser/PyTorch-VAE) has no FFI: its "plugin API" for this path is the
 * Python class models/vq_vae.py:19 VectorQuantizerEMA, whose arithmetic is a sequence of
 * ATen calls.  Each entry point below replaces one group of those calls (cited per
 * function, paths relative to the reference root) and is what a ctypes/cffi/pybind binding
 * on the reference side would bind; see INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is DEVICE memory unless it says "host"
 *   - row-major, contiguous; float rows must be 16-byte aligned (D % 4 == 0)
 *   - caller allocates every output and workspace.  The library owns no data; the only process state is
 *     (a) per device, three helper streams + events created on first use by the chunk pipeline of
 *     vqb200_search / vqb200_quantize (never destroyed), (b) one-time cudaFuncSetAttribute flags per device and
 *     (c) the vqb200_timing_* measurement hooks, which are process-global, NOT thread-safe and meant for bench.py
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); nothing synchronises
 *   - return 0 on success, a negative VQB200_E* for argument errors, or a positive
 *     cudaError_t from the launch.  Nothing throws.  No CPU fallback exists.
 */
#ifndef VQ_B200_H_
#define VQ_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define VQB200_API __attribute__((visibility("default")))
#else
#define VQB200_API
#endif

#define VQB200_ABI_VERSION 18
#define VQB200_MAX_LEVELS 32
#define VQB200_LEVEL_META_FLOATS 8 /* per level: [0] max|e|, [1] non-finite flag, [2] max|bf16(e)|,
                                      [3] max|e - bf16(e)|, [4] max|f16(e)|, [5] max|e - f16(e)|, [6..7] internal (dead-code de-duplication) */

enum {
  VQB200_OK = 0,
  VQB200_EINVAL = -1,     /* null pointer / negative size */
  VQB200_ESHAPE = -2,     /* unsupported D, K or level layout */
  VQB200_EALIGN = -3,     /* pointer not 16-byte aligned */
  VQB200_EWORKSPACE = -4, /* workspace too small */
  VQB200_EDRIVER = -5     /* tensor-map encode / driver entry point failure */
};

enum {
  VQB200_MODE_FP32_EXACT = 0, /* indices of the fp32 inputs; tensor-core candidates are re-ranked exactly */
  VQB200_MODE_BF16_INPUT = 1  /* inputs rounded to bf16 (RN), products and sums in fp32 */
};

enum { VQB200_IDX_I16 = 2, VQB200_IDX_I32 = 4, VQB200_IDX_I64 = 8 };

VQB200_API int vqb200_abi_version(void);
VQB200_API const char* vqb200_status_string(int status);

/* Which search kernel `vqb200_search` picks for a shape: 0 = SIMT fp32, 1 = tcgen05 tensor core. */
VQB200_API int vqb200_search_path(int64_t N, int K, int D, int mode);

/* Per-code cache refresh.  Replaces `self.embedding.pow(2).sum(1)` (models/vq_vae.py:186,241),
 * recomputed by the reference on every forward, plus the operand conversion for the tensor path.
 *   E          [K_total, D] fp32 codebook (the module's `embedding` buffer, source of truth)
 *   K_per      codes per residual level (K_total % K_per == 0, at most VQB200_MAX_LEVELS levels)
 *   E_bf16     [2, K_total, D] 16-bit tensor-core operand copies: plane 0 = bf16 (RN), the operand of
 *              bf16_input mode; plane 1 = fp16 (RN, subnormals flushed), the operand of fp32 mode -- three more
 *              significand bits shrink the admission margin, and with it the exact re-rank, eightfold
 *   ee_half    [2, K_total]: plane 0 = |e|^2/2 of the fp32 rows, plane 1 = of the bf16-rounded rows
 *   level_meta [levels, 8] fp32, see VQB200_LEVEL_META_FLOATS
 */
VQB200_API int vqb200_codebook_prepare(const float* E, int K_total, int D, int K_per, uint16_t* E_bf16,
                            float* ee_half, float* level_meta, void* stream);

/* Nearest-code search for ONE level: replaces the distance assembly + argmin
 * (models/vq_vae.py:183-188 and :238-245).  The N x K distance matrix is never written.
 *   z          [N, D] fp32 latents (or the RVQ residual)
 *   E, ee_half(plane 0), ee_half_bf16(plane 1), level_meta: this level's slices of the cache
 *   E_bf16     this level's slice of the operand plane OF THE MODE (plane 1 / fp16 for fp32 mode, plane 0 /
 *              bf16 for bf16_input mode); same meaning in vqb200_quantize and vqb200_quantize_fused
 *   idx_out    [N] int64 = idx_offset + argmin_k |z - e_k|^2; first index on exact ties; a NaN
 *              distance counts as the minimum (torch.argmin semantics)
 */
VQB200_API size_t vqb200_search_workspace_bytes(int64_t N, int K, int D, int mode);
/* Number of kernels one vqb200_search call of this shape launches (for launch accounting). */
VQB200_API int vqb200_search_launches(int64_t N, int K, int D, int mode);
VQB200_API int vqb200_search(const float* z, int64_t N, int D, const float* E, const uint16_t* E_bf16,
                  const float* ee_half, const float* ee_half_bf16, const float* level_meta, int K,
                  int mode, int64_t idx_offset, int64_t* idx_out, void* workspace,
                  size_t workspace_bytes, void* stream);

/* Residual VQ, between two levels (models/vq_vae.py:258 then :238 of the next level), in ONE read of the rows:
 *   residual_out = fl(z - E_full[idx])   [N, D] fp32
 *   z16_out      = the 16-bit tensor-core operand copy of residual_out for `mode` (fp16 / bf16)   [N, D]
 *   margin_out   = the admission margins of the NEXT level's search for those rows (next_level_meta)   [N]
 * vqb200_search_prepped is vqb200_search given that copy and those margins (it skips its own pre-pass); it exists
 * for shapes vqb200_search_path() sends to the tensor kernel and returns VQB200_ESHAPE otherwise. */
VQB200_API int vqb200_residual_prep(const float* z, const float* E_full, const int64_t* idx, int64_t N, int D,
                         int K_total, int mode, const float* next_level_meta, float* residual_out,
                         uint16_t* z16_out, float* margin_out, void* stream);
VQB200_API int vqb200_search_prepped(const float* z, const uint16_t* z16, const float* margin, int64_t N, int D,
                          const float* E, const uint16_t* E_bf16, const float* ee_half,
                          const float* ee_half_bf16, const float* level_meta, int K, int mode,
                          int64_t idx_offset, int64_t* idx_out, void* workspace, size_t workspace_bytes,
                          void* stream);

/* The whole eval-mode residual forward (models/vq_vae.py:226-263 without the EMA branch) in ONE call: per level
 * search -> residual update fused with the next level's pre-pass, then vqb200_rvq_finalize.  Same results as the
 * per-level entry points; it exists because at the stage-2 shape (N = 8192) the ~20 separate calls cost more host
 * time than their kernels cost GPU time.  E, E_lp (operand plane of the mode), ee_half (plane 0), ee_half_bf16
 * (plane 1) and level_meta are the WHOLE cache arrays ([K_per * L, ...]); idx_out [L * N] level-major global ids. */
VQB200_API size_t vqb200_rvq_forward_workspace_bytes(int64_t N, int K_per, int D, int L, int mode);
VQB200_API int vqb200_rvq_forward_launches(int64_t N, int K_per, int D, int L, int mode);
/* 1 when vqb200_rvq_forward runs the shape as ONE persistent kernel (csrc/vq_rvq_fused.cuh: a CTA owns 64 or 128 rows and walks
 * all levels -- operand tile of the residual in shared memory, tcgen05 scores in TMEM, candidate records in shared
 * memory, exact re-rank, residual update, outputs -- without a launch between levels): D in {128, 256, 384, 512},
 * K_per >= 128, 2 <= L <= 8. */
VQB200_API int vqb200_rvq_fused_supported(int64_t N, int K_per, int D, int L, int mode);
VQB200_API int vqb200_rvq_forward(const float* z, int64_t N, int D, const float* E, const uint16_t* E_lp,
                       const float* ee_half, const float* ee_half_bf16, const float* level_meta, int K_per, int L,
                       int mode, int64_t* idx_out, float* zq_out, float* zq_st_out, double* sqerr_sum,
                       int32_t* hist, void* workspace, size_t workspace_bytes, void* stream);

/* The training-mode residual forward with the EMA update after every level (models/vq_vae.py:226-263 incl. :251,
 * each rank updating from its own rows, no mask) in ONE call: per level search -> gather (z_q accumulated BEFORE the
 * codebook moves, residual, histogram) -> scatter-add -> EMA finalize + cache refresh; then the straight-through /
 * loss pass.  E, E_lp_planes ([2, K_total, D]), ee_half ([2, K_total]), level_meta, ema_* are the whole arrays and
 * are UPDATED in place.  Same results as the per-level entry points. */
VQB200_API size_t vqb200_rvq_train_workspace_bytes(int64_t N, int K_per, int D, int L, int mode);
VQB200_API int vqb200_rvq_train_launches(int64_t N, int K_per, int D, int L, int mode);
VQB200_API int vqb200_rvq_train_forward(const float* z, int64_t N, int D, float* E, uint16_t* E_lp_planes,
                             float* ee_half, float* level_meta, int K_per, int L, int mode, float decay,
                             float one_minus_decay, float eps, float* ema_cluster_size, float* ema_embedding,
                             int64_t* idx_out, float* zq_out, float* zq_st_out, double* sqerr_sum, int32_t* hist,
                             void* workspace, size_t workspace_bytes, void* stream);

/* One level of the training-mode residual forward up to the exchange point of a multi-GPU EMA update:
 * search -> gather (zq_out accumulated when level > 0, residual_out for the next level or NULL, histogram) ->
 * segment sums of THIS rank (seg_sum / seg_cnt zeroed here).  The caller all-reduces the sums and calls
 * vqb200_ema_finalize.  Arrays are the whole cache arrays, as in vqb200_rvq_train_forward. */
VQB200_API int vqb200_rvq_train_level(const float* residual, int64_t N, int D, const float* E,
                           const uint16_t* E_lp_planes, const float* ee_half, const float* level_meta, int K_per,
                           int L, int level, int mode, int64_t* idx_out, float* zq_out, float* residual_out,
                           int32_t* hist, float* seg_sum, float* seg_cnt, void* workspace, size_t workspace_bytes,
                           void* stream);

/* Measurement hook (bench.py): while enabled, vqb200_search brackets every launch of its dominant kernel (the
 * tcgen05 search kernel; the SIMT kernel on shapes that take the SIMT path) with CUDA events on the launching
 * stream.  vqb200_timing_collect synchronises those events, returns their summed duration and launch count and
 * clears the list.  Not thread-safe; off by default (no events are created). */
VQB200_API int vqb200_timing_enable(int on);
VQB200_API int vqb200_timing_collect(float* total_ms, int* n_launches);

/* Gather + everything elementwise that follows it, in one pass over the rows.
 * Replaces F.embedding (:189,248), the straight-through expression (:199,263), the RVQ
 * residual update (:258) and level sum (:261), F.mse_loss partial sums (:1293) and
 * torch.bincount (:205-207,266).  Every output is optional (NULL = skip).
 *   idx           [N] int64 GLOBAL code ids (rows of E)
 *   zq_out        [N, D]: = E[idx], or += E[idx] when zq_accumulate != 0 (RVQ levels > 0)
 *   zq_st_out     [N, D]: fl(z + fl(E[idx] - z))
 *   residual_out  [N, D]: fl(z - E[idx])
 *   sqerr_sum     [1] double, += sum (E[idx] - z)^2   (caller zeroes)
 *   hist          [K_total] int32, += 1 per row with row_mask[row] != 0 (row_mask NULL = all rows)
 */
VQB200_API int vqb200_gather(const float* z, const float* E, const int64_t* idx, int64_t N, int D, int K_total,
                  float* zq_out, int zq_accumulate, float* zq_st_out, float* residual_out,
                  double* sqerr_sum, int32_t* hist, const uint8_t* row_mask, void* stream);

/* The whole single-level forward in ONE kernel (small code dimensions: D = 64 and D = 128): search, exact re-rank,
 * gather, straight-through, commitment partial sum and histogram, reading the fp32 latents once.  Same
 * outputs, bit for bit, as vqb200_search followed by vqb200_gather.  Every output except idx_out is optional.
 * vqb200_quantize_fused_supported returns 1 when the shape takes this path. */
VQB200_API int vqb200_quantize_fused_supported(int64_t N, int K, int D, int mode);
VQB200_API size_t vqb200_quantize_fused_workspace_bytes(int64_t N, int K, int D, int mode);
VQB200_API int vqb200_quantize_fused(const float* z, int64_t N, int D, const float* E, const uint16_t* E_bf16,
                          const float* ee_half, const float* ee_half_bf16, const float* level_meta, int K,
                          int mode, int64_t idx_offset, int64_t* idx_out, float* zq_out, float* zq_st_out,
                          double* sqerr_sum, int32_t* hist, const uint8_t* row_mask, void* workspace,
                          size_t workspace_bytes, void* stream);

/* vqb200_search + vqb200_gather for one single-level codebook in ONE call: on the tensor path the gather of each
 * chunk of rows is queued behind that chunk's re-rank, so it overlaps the tensor kernel of the next chunk.
 * Workspace as vqb200_search_workspace_bytes.  E_full / K_total: the whole codebook (idx are global ids). */
VQB200_API int vqb200_quantize(const float* z, int64_t N, int D, const float* E, const uint16_t* E_bf16,
                    const float* ee_half, const float* ee_half_bf16, const float* level_meta, int K, int mode,
                    int64_t idx_offset, int64_t* idx_out, const float* E_full, int K_total, float* zq_out,
                    float* zq_st_out, double* sqerr_sum, int32_t* hist, const uint8_t* row_mask, void* workspace,
                    size_t workspace_bytes, void* stream);

/* Straight-through value and commitment partial sum from z and an already-summed z_q
 * (the RVQ tail, models/vq_vae.py:263 and :1293). */
VQB200_API int vqb200_st_loss(const float* z, const float* zq, int64_t n_elems, float* zq_st_out,
                   double* sqerr_sum, void* stream);

/* Usage statistics: replaces the ~12 small ATen kernels of models/vq_vae.py:209-222, 267-280.
 *   stats_out [3] = perplexity, dead_ratio, commitment mse (= *sqerr_sum * inv_elems; 0 if NULL)
 *   ep_usage  [K_total] += usage;  ep_cnt [1] += count_add   (either may be NULL)
 */
VQB200_API int vqb200_stats_finalize(const int32_t* hist, int K_total, float count_add, const double* sqerr_sum,
                          double inv_elems, float* ep_usage, float* ep_cnt, float* stats_out,
                          void* stream);

/* Multi-GPU statistics: vqb200_stats_pack writes this rank's float64 pack [sum sq err | element count | histogram]
 * (K_total + 2 doubles: counts stay exact to 2^53), the caller all-reduces it (SUM) and
 * vqb200_stats_finalize_packed is vqb200_stats_finalize on the reduced pack (stats_out[2] = global mean sq err;
 * ep_cnt += levels * (reduced element count / D), the GLOBAL number of quantized positions). */
VQB200_API int vqb200_stats_pack(const int32_t* hist, int K_total, const double* sqerr_sum, double n_elems,
                      double* packed_out, void* stream);
VQB200_API int vqb200_stats_finalize_packed(const double* packed, int K_total, int levels, int D, float* ep_usage,
                                 float* ep_cnt, float* stats_out, void* stream);

/* The same exchange WITHOUT a collective library call on the step path: pack -> exchange over NVLink / NVSwitch peer
 * memory -> reduce -> finalize in ONE kernel (north star: "only the scalar loss and the code-usage histogram are
 * all-reduced").  Every rank passes peer_buffers, a DEVICE array of `world` pointers to the ranks' symmetric buffers
 * (the same allocation mapped into every process, e.g. torch.distributed._symmetric_memory; each
 * vqb200_stats_exchange_buffer_bytes(K_total, world) long and ZEROED once before the first call, with a barrier after
 * the zeroing).  A rank pushes its pack into its slot of every rank's buffer, raises a flag there, waits for all
 * flags of the epoch and sums the slots in rank order, so every rank obtains bit-identical statistics; the epoch counter
 * lives in the buffer (graph-replayable).  All ranks must call in the same order.  spin_limit: polls of a flag before
 * the kernel traps (0 = wait for ever). */
VQB200_API size_t vqb200_stats_exchange_buffer_bytes(int K_total, int world);
VQB200_API int vqb200_stats_exchange(const int32_t* hist, int K_total, const double* sqerr_sum, double n_elems, int levels,
                          int D, const uint64_t* peer_buffers, int rank, int world, uint64_t spin_limit,
                          float* ep_usage, float* ep_cnt, float* stats_out, void* stream);

/* EMA codebook update, part 1: segment sums.  Replaces the dense one-hot GEMM of
 * models/vq_vae.py:81-83.  seg_sum [K_total, D] and seg_cnt [K_total] are zeroed by the caller. */
VQB200_API int vqb200_scatter_add(const float* z, const int64_t* idx, const uint8_t* row_mask, int64_t N, int D,
                       int K_total, float* seg_sum, float* seg_cnt, void* stream);

/* EMA codebook update, part 2 (models/vq_vae.py:85-89) fused with the cache refresh:
 *   cs <- fl(fl(cs*decay) + fl(n*(1-decay)));  es likewise;  E <- es / (cs + eps)  for ALL codes.
 * one_minus_decay is passed separately because the reference forms it in double. */
VQB200_API int vqb200_ema_finalize(const float* seg_sum, const float* seg_cnt, float decay, float one_minus_decay,
                        float eps, int K_total, int D, int K_per, float* ema_cluster_size,
                        float* ema_embedding, float* E, uint16_t* E_bf16, float* ee_half,
                        float* level_meta, void* stream);

/* Lloyd step of the k-means codebook initialiser (the `--init_codebook` centroids of run.py:74-89 that
 * models/vq_vae.py:577-613 copies in): E[k] <- seg_sum[k] / seg_cnt[k] where seg_cnt[k] > 0, an empty cluster
 * keeps its centroid; the derived cache is refreshed in the same pass.  seg_* come from vqb200_scatter_add. */
VQB200_API int vqb200_kmeans_finalize(const float* seg_sum, const float* seg_cnt, int K_total, int D, int K_per,
                           float* E, uint16_t* E_bf16, float* ee_half, float* level_meta, void* stream);

/* Residual-VQ tail in one pass (eval mode, <= 8 levels): from the level-major GLOBAL ids
 * idx[l*level_stride + n] (level_stride = N for a packed [L*N] array),
 *   zq_out = ((E[i_0] + E[i_1]) + ...) in level order (models/vq_vae.py:261), zq_st_out = fl(z + fl(zq - z)) (:263),
 *   sqerr_sum += sum (zq - z)^2 (:1293), hist[i_l] += 1 for every level (:266).  Every output is optional.
 * Replaces the per-level z_q accumulation of vqb200_gather plus vqb200_st_loss when the codebook does not change
 * between levels (no EMA update): z_q is written once instead of being re-read and re-written per level. */
VQB200_API int vqb200_rvq_finalize(const float* z, const int64_t* idx_level_major, int64_t level_stride,
                        int64_t N, int D, int L, const float* E, int K_total, float* zq_out, float* zq_st_out, double* sqerr_sum,
                        int32_t* hist, void* stream);

/* Usage-entropy regulariser (models/vq_vae.py:1298-1309): the code-usage distribution
 *   p_code[k] = (1/N) sum_n softmax_k(z_n . e_k)
 * without the [N, K] logits / probabilities.  vqb200_usage_probs ADDS sum_n P_nk to p_sum [K] (caller zeroes; divide
 * by N) and stores the per-row (max logit, 1 / sum exp) in row_stats [N, 2] for the backward.
 * vqb200_usage_probs_backward: grad_z_n = scale * sum_j P_nj (g_j - sum_k P_nk g_k) e_j with g = dLoss/dp_code
 * (pass scale = 1 / N).  D <= 512. */
VQB200_API int vqb200_usage_probs(const float* z, int64_t N, int D, const float* E, int K, float* p_sum,
                       float* row_stats, void* stream);
VQB200_API int vqb200_usage_probs_backward(const float* z, int64_t N, int D, const float* E, int K,
                                const float* row_stats, const float* grad_p, float scale, float* grad_z_out,
                                void* stream);

/* Soft assignment of the soft-VQ training path (models/vq_vae.py:838-843, single-level codebooks):
 *   z_soft[n] = sum_k softmax_k(-|z_n - e_k|^2 / max(1e-8, tau)) e_k
 * in one pass with an online softmax; neither the [N, K, D] differences nor the [N, K] logits are stored.
 * The result is detached in the reference (:852), so there is no backward.  D <= 512. */
VQB200_API int vqb200_soft_assign(const float* z, int64_t N, int D, const float* E, int K, float tau,
                       float* z_soft_out, void* stream);

/* The shared first half of both, as a register-tiled fp32 contraction (128 rows x 128 codes per CTA) instead of one
 * warp reduction per (row, code) -- 10-20 x the throughput of the two entry points above at small D:
 *   probs[n, k] = softmax_k(alpha * (z_n . e_k) + beta[k])          (beta may be NULL)
 * soft-VQ: alpha = 2 / tau, beta[k] = -|e_k|^2 / tau (the row constant cancels); usage regulariser: alpha = 1.
 * Two sweeps (per-row max / sum exp, then the probabilities).  Any of the outputs may be NULL:
 *   row_stats [N, 2] (max logit, 1 / sum exp), probs_out [N, K], p_sum [K] += sum_n probs[n, k] (caller zeroes).
 * The second contraction of either path (probs @ E, dS @ E) is a plain GEMM and is left to the caller's library.
 * workspace: vqb200_softmax_rows_workspace_bytes(N, K), 8-byte aligned. */
VQB200_API size_t vqb200_softmax_rows_workspace_bytes(int64_t N, int K);
VQB200_API int vqb200_softmax_rows(const float* z, int64_t N, int D, const float* E, const float* beta, int K, float alpha,
                        float* row_stats, float* probs_out, float* p_sum, void* workspace, size_t workspace_bytes,
                        void* stream);

/* Backward of the two differentiable outputs (straight-through + commitment):
 *   grad_z = grad_st + (*grad_commit) * scale * (z - zq),  scale = 2 / (N D)
 * grad_st may be NULL (treated as 0); grad_commit is a DEVICE scalar (no host sync), NULL = 0. */
VQB200_API int vqb200_commit_backward(const float* grad_st, const float* grad_commit, const float* z,
                           const float* zq, int64_t n_elems, float scale, float* grad_z_out,
                           void* stream);

/* Wire formats either side of the path.
 * Level-major flat RVQ ids [Q, B, M] -> token-major [B, M*Q], optionally narrowed
 * (scripts/extract_code_indices.py:195-209, :494-549 narrows on the host). */
VQB200_API int vqb200_relayout_indices(const int64_t* idx_level_major, int Q, int64_t B, int64_t M, void* out,
                            int out_elem_bytes, void* stream);
/* Token-major ids [n_tok * Q] -> z_q [n_tok, D] = sum over the Q levels in level order
 * (scripts/decode_with_vqvae.py:110-130; models/vq_vae.py:1404-1418). */
VQB200_API int vqb200_indices_to_latent(const void* idx, int idx_elem_bytes, int64_t n_tok, int Q, const float* E,
                             int K_total, int D, float* zq_out, void* stream);

/* vqb200_rvq_forward / vqb200_rvq_train_forward followed by vqb200_stats_finalize(hist, K_per * L, count_add, sqerr_sum,
 * inv_elems, ep_usage, ep_cnt, stats_out) -- inside the persistent kernel where it runs (its last CTA to finish turns
 * the histogram into the statistics: one launch fewer on a 0.14 ms forward), as a separate launch elsewhere.
 * hist, sqerr_sum and stats_out are required. */
VQB200_API int vqb200_rvq_forward_stats(const float* z, int64_t N, int D, const float* E, const uint16_t* E_lp,
                             const float* ee_half, const float* ee_half_bf16, const float* level_meta, int K_per, int L,
                             int mode, int64_t* idx_out, float* zq_out, float* zq_st_out, double* sqerr_sum,
                             int32_t* hist, void* workspace, size_t workspace_bytes, float count_add, double inv_elems,
                             float* ep_usage, float* ep_cnt, float* stats_out, void* stream);
VQB200_API int vqb200_rvq_train_forward_stats(const float* z, int64_t N, int D, float* E, uint16_t* E_lp_planes,
                                   float* ee_half, float* level_meta, int K_per, int L, int mode, float decay,
                                   float one_minus_decay, float eps, float* ema_cluster_size, float* ema_embedding,
                                   int64_t* idx_out, float* zq_out, float* zq_st_out, double* sqerr_sum, int32_t* hist,
                                   void* workspace, size_t workspace_bytes, float count_add, double inv_elems,
                                   float* ep_usage, float* ep_cnt, float* stats_out, void* stream);

/* The training forward of a residual codebook in two halves, for data-parallel training with the EMA segment sums
 * all-reduced over ranks: the reference runs one EMA update per level and each touches all
 * K_total codes, but for the codes of ANOTHER level it is a decay-only step -- so level l is searched against codes that
 * depend on nothing computed in this step, and the whole forward can run before ONE exchange of the segment sums:
 *   begin : the decay-only steps of the earlier levels (level l: l of them), then every level in one persistent
 *           kernel (search, exact re-rank, residual, outputs as vqb200_rvq_forward) which also reduces the residual
 *           rows into seg_sum [K_total, D] / seg_cnt [K_total] (zeroed by the call);
 *   (the caller all-reduces seg_sum / seg_cnt over its ranks)
 *   finish: every level's own update from its segment sums and the decay-only steps of the later levels; codebook
 *           cache refreshed.  begin + finish with nothing in between = vqb200_rvq_train_forward, bit for bit.
 * Shapes: vqb200_rvq_train_fused_supported (D in {128, 256, 384, 512}, K_per >= 128, 2 <= L <= 8, N <= 65536 -- 49152 at D = 256 / 384). */
VQB200_API int vqb200_rvq_train_fused_supported(int64_t N, int K_per, int D, int L, int mode);
VQB200_API size_t vqb200_rvq_train_begin_workspace_bytes(int64_t N, int K_per, int D, int L, int mode);
VQB200_API int vqb200_rvq_train_begin(const float* z, int64_t N, int D, float* E, uint16_t* E_lp_planes, float* ee_half,
                           float* level_meta, int K_per, int L, int mode, float decay, float one_minus_decay, float eps,
                           float* ema_cluster_size, float* ema_embedding, int64_t* idx_out, float* zq_out,
                           float* zq_st_out, double* sqerr_sum, int32_t* hist, float* seg_sum, float* seg_cnt,
                           void* workspace, size_t workspace_bytes, void* stream);
VQB200_API int vqb200_rvq_train_finish(const float* seg_sum, const float* seg_cnt, float decay, float one_minus_decay,
                            float eps, int K_per, int L, int D, float* ema_cluster_size, float* ema_embedding, float* E,
                            uint16_t* E_lp_planes, float* ee_half, float* level_meta, void* stream);

/* Token-major ids [n_tok * Q] -> decoder memory [n_tok, H] = LayerNorm(from_code(z_q)) (models/vq_vae.py:749,
 * "from_code + mem_ln after K2").  from_code is linear, so with the projected table
 * P = E W^T ([K_total, H], the caller rebuilds it when the codebook or the weight changes) the Linear is a gather of Q
 * rows of P per token plus the bias; the LayerNorm (biased variance, eps, optional affine) runs on the row in registers.
 * No GEMM over the tokens and z_q never reaches memory.  H % 4 == 0, H <= 1024; bias / ln_weight / ln_bias may be NULL. */
VQB200_API int vqb200_indices_to_memory(const void* idx, int idx_elem_bytes, int64_t n_tok, int Q, const float* P,
                             int K_total, int H, const float* bias, const float* ln_weight, const float* ln_bias,
                             float ln_eps, float* memory_out, void* stream);

/* Codebook-sharded search support: a (distance, index) pair packed into one
 * orderable uint64 so that an all-reduce(MIN) performs a tie-stable min-loc.
 *   search_packed: like vqb200_search but writes packed[N] = key(d) << 32 | (idx_offset + argmin)
 *   unpack: packed -> int64 ids */
VQB200_API int vqb200_search_packed(const float* z, int64_t N, int D, const float* E, const float* ee_half, int K,
                         int64_t idx_offset, uint64_t* packed_out, void* stream);
VQB200_API int vqb200_minloc_unpack(const uint64_t* packed, int64_t N, int64_t* idx_out, void* stream);

/* The same on the TENSOR path (north star: "a codebook-sharded min-loc reduction is used when K.D exceeds the per-SM
 * staging budget"): each rank runs vqb200_search over its slice of the codes (exact winner of the slice, global id),
 *   pack_exact: packed[n] = (orderable(fp64 score of z_n against E_full[idx[n]]) with its low 24 bits cleared) | idx[n]
 * (score = |e|^2/2 - z.e, fp64 accumulation; 40 bits of score + 24 bits of id, K_total <= 2^24), the ranks all-reduce(MIN)
 * the packed words, and unpack24 extracts the ids.  Equal scores to 4e-9 relative resolve to the lower id. */
VQB200_API int vqb200_pack_exact(const float* z, int64_t N, int D, const float* E_full, int K_total, const int64_t* idx,
                      uint64_t* packed_out, void* stream);
VQB200_API int vqb200_minloc_unpack24(const uint64_t* packed, int64_t N, int64_t* idx_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VQ_B200_H_ */
