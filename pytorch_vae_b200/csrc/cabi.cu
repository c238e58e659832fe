// extern "C" entry points of libvqb200.so: argument validation + dispatch.  See include/vq_b200.h.
#include "common.cuh"

#include <utility>
#include <vector>

using namespace vqb;

namespace vqb {
static bool g_timing = false;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_events;
void timing_mark_begin(cudaStream_t s) {
  if (!g_timing) return;
  cudaEvent_t a, b;
  if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
  g_events.emplace_back(a, b);
  cudaEventRecord(a, s);
}
void timing_mark_end(cudaStream_t s) {
  if (!g_timing || g_events.empty()) return;
  cudaEventRecord(g_events.back().second, s);
}
}  // namespace vqb

#define VQ_REQUIRE(cond, code) \
  do {                         \
    if (!(cond)) return (code); \
  } while (0)

static bool shape_ok(int D) { return D >= 4 && D % 4 == 0 && D <= 4096; }

extern "C" {

int vqb200_abi_version(void) { return VQB200_ABI_VERSION; }

const char* vqb200_status_string(int status) {
  switch (status) {
    case VQB200_OK: return "ok";
    case VQB200_EINVAL: return "invalid argument (null pointer or negative size)";
    case VQB200_ESHAPE: return "unsupported shape (D must be a multiple of 4, K_total a multiple of K_per)";
    case VQB200_EALIGN: return "pointer not 16-byte aligned";
    case VQB200_EWORKSPACE: return "workspace too small";
    case VQB200_EDRIVER: return "CUDA driver entry point / tensor-map failure";
    default: return status > 0 ? cudaGetErrorString(static_cast<cudaError_t>(status)) : "unknown status";
  }
}

int vqb200_timing_enable(int on) {
  g_timing = on != 0;
  return VQB200_OK;
}

int vqb200_timing_collect(float* total_ms, int* n_launches) {
  VQ_REQUIRE(total_ms && n_launches, VQB200_EINVAL);
  float total = 0.f;
  int n = 0;
  for (auto& ev : g_events) {
    float ms = 0.f;
    if (cudaEventSynchronize(ev.second) == cudaSuccess && cudaEventElapsedTime(&ms, ev.first, ev.second) == cudaSuccess) {
      total += ms;
      ++n;
    }
    cudaEventDestroy(ev.first);
    cudaEventDestroy(ev.second);
  }
  g_events.clear();
  *total_ms = total;
  *n_launches = n;
  return VQB200_OK;
}

int vqb200_search_path(int64_t N, int K, int D, int mode) {
  (void)mode;
  return tc_supported(N, K, D) ? (tc_side_pipeline(N, K, D) ? 2 : 1) : 0;
}

int vqb200_codebook_prepare(const float* E, int K_total, int D, int K_per, uint16_t* E_bf16, float* ee_half,
                            float* level_meta, void* stream) {
  VQ_REQUIRE(E && E_bf16 && ee_half && level_meta, VQB200_EINVAL);
  VQ_REQUIRE(K_total > 0 && K_per > 0, VQB200_EINVAL);
  VQ_REQUIRE(shape_ok(D) && K_total % K_per == 0 && K_total / K_per <= VQB200_MAX_LEVELS, VQB200_ESHAPE);
  VQ_REQUIRE(aligned16(E) && aligned16(E_bf16), VQB200_EALIGN);
  return launch_codebook_refresh(0, nullptr, nullptr, 0.f, 0.f, 0.f, K_total, D, K_per, nullptr, nullptr,
                                 const_cast<float*>(E), E_bf16, ee_half, level_meta,
                                 static_cast<cudaStream_t>(stream));
}

size_t vqb200_search_workspace_bytes(int64_t N, int K, int D, int mode) {
  (void)mode;
  return tc_supported(N, K, D) ? tc_workspace_bytes(N, K, D) : 256;
}

int vqb200_search_launches(int64_t N, int K, int D, int mode) {
  (void)mode;
  if (N <= 0) return 0;
  return tc_supported(N, K, D) ? tc_launches(N, K, D) : 1;
}

// ---- the whole eval-mode residual forward in ONE call (host overhead: one ctypes call instead of ~20) ----
static size_t rvq_align(size_t v) { return (v + 255) / 256 * 256; }

size_t vqb200_rvq_forward_workspace_bytes(int64_t N, int K_per, int D, int L, int mode) {
  if (rvq_fused_supported(N, K_per, D, L)) return rvq_align(rvq_fused_workspace_bytes(N, D));
  const size_t rows = static_cast<size_t>(N > 0 ? N : 0);   // two residual buffers whatever the number of levels
  return rvq_align(vqb200_search_workspace_bytes(N, K_per, D, mode)) + 2 * rvq_align(rows * D * 4) +
         rvq_align(rows * D * 2) + rvq_align(rows * 4);
}

int vqb200_rvq_fused_supported(int64_t N, int K_per, int D, int L, int mode) {
  (void)mode;
  return rvq_fused_supported(N, K_per, D, L) ? 1 : 0;
}

int vqb200_rvq_forward_launches(int64_t N, int K_per, int D, int L, int mode) {
  if (N <= 0 || L < 1) return 0;
  if (rvq_fused_supported(N, K_per, D, L)) return 1;      // the persistent kernel
  const int per_search = vqb200_search_launches(N, K_per, D, mode);
  const bool tc = tc_supported(N, K_per, D);
  const int chunks = tc ? per_search / 5 : 0;
  // level 0: full search; levels > 0 on the tensor path skip their pre-pass; L-1 residual kernels; one finalize
  return per_search + (L - 1) * (per_search - chunks) + (L - 1) + 1;
}

static int rvq_forward_impl(const float* z, int64_t N, int D, const float* E, const uint16_t* E_lp, const float* ee_half,
                            const float* ee_half_bf16, const float* level_meta, int K_per, int L, int mode,
                            int64_t* idx_out, float* zq_out, float* zq_st_out, double* sqerr_sum, int32_t* hist,
                            void* workspace, size_t workspace_bytes, void* stream, const RvqStatsTail* tail) {
  VQ_REQUIRE(N >= 0 && K_per > 0 && L >= 1 && L <= 8, VQB200_EINVAL);
  if (N == 0) return VQB200_OK;
  VQ_REQUIRE(z && idx_out && E && E_lp && ee_half && ee_half_bf16 && level_meta && workspace, VQB200_EINVAL);
  VQ_REQUIRE(mode == VQB200_MODE_FP32_EXACT || mode == VQB200_MODE_BF16_INPUT, VQB200_EINVAL);
  VQ_REQUIRE(shape_ok(D), VQB200_ESHAPE);
  VQ_REQUIRE(aligned16(z) && aligned16(E) && aligned16(E_lp) && aligned16(zq_out) && aligned16(zq_st_out) &&
                 (reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, VQB200_EALIGN);
  VQ_REQUIRE(workspace_bytes >= vqb200_rvq_forward_workspace_bytes(N, K_per, D, L, mode), VQB200_EWORKSPACE);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool bf = mode == VQB200_MODE_BF16_INPUT;
  if (rvq_fused_supported(N, K_per, D, L))                // every level inside ONE persistent kernel
    return launch_rvq_fused(z, N, D, E, E_lp, bf ? ee_half_bf16 : ee_half, level_meta, K_per, L, mode, idx_out, zq_out,
                            zq_st_out, sqerr_sum, hist, workspace, workspace_bytes, s, nullptr, nullptr, tail);
  const bool tc = tc_supported(N, K_per, D);
  const int K_total = K_per * L;
  uint8_t* w = static_cast<uint8_t*>(workspace);
  void* ws_search = w;
  const size_t ws_search_bytes = vqb200_search_workspace_bytes(N, K_per, D, mode);
  w += rvq_align(ws_search_bytes);
  float* res[2];
  res[0] = reinterpret_cast<float*>(w); w += rvq_align(static_cast<size_t>(N) * D * 4);
  res[1] = reinterpret_cast<float*>(w); w += rvq_align(static_cast<size_t>(N) * D * 4);
  uint16_t* z16 = reinterpret_cast<uint16_t*>(w); w += rvq_align(static_cast<size_t>(N) * D * 2);
  float* margin = reinterpret_cast<float*>(w);

  const float* residual = z;
  for (int l = 0; l < L; ++l) {
    const int64_t s0 = static_cast<int64_t>(l) * K_per;
    const float* El = E + s0 * D;
    const uint16_t* Elp = E_lp + s0 * D;
    const float* meta = level_meta + l * VQB200_LEVEL_META_FLOATS;
    int64_t* idx_l = idx_out + static_cast<int64_t>(l) * N;
    int st;
    if (tc) {
      PrepArgs prep{z16, margin};
      st = launch_search_tc(residual, N, D, El, Elp, ee_half + s0, ee_half_bf16 + s0, meta, K_per, mode, s0, idx_l,
                            ws_search, ws_search_bytes, s, nullptr, l > 0 ? &prep : nullptr);
    } else {
      st = launch_search_simt(residual, nullptr, N, D, El, (bf ? ee_half_bf16 : ee_half) + s0, K_per, bf ? 1 : 0, s0,
                              idx_l, nullptr, s);
    }
    if (st != VQB200_OK) return st;
    if (l + 1 < L) {
      float* nxt = res[l & 1];
      st = tc ? launch_residual_prep(residual, E, idx_l, N, D, K_total, mode, meta + VQB200_LEVEL_META_FLOATS, nxt, z16,
                                     margin, s)
              : launch_gather(residual, E, idx_l, N, D, K_total, nullptr, 0, nullptr, nxt, nullptr, nullptr, nullptr, s);
      if (st != VQB200_OK) return st;
      residual = nxt;
    }
  }
  int st = launch_rvq_finalize(z, idx_out, N, N, D, L, E, K_total, zq_out, zq_st_out, sqerr_sum, hist, s);
  if (st == VQB200_OK && tail && tail->stats_out && hist)   // shapes the persistent kernel does not take: a separate launch
    st = launch_stats_finalize(hist, K_total, tail->count_add, sqerr_sum, tail->inv_elems, tail->ep_usage, tail->ep_cnt,
                               tail->stats_out, s);
  return st;
}

int vqb200_rvq_forward(const float* z, int64_t N, int D, const float* E, const uint16_t* E_lp, const float* ee_half,
                       const float* ee_half_bf16, const float* level_meta, int K_per, int L, int mode,
                       int64_t* idx_out, float* zq_out, float* zq_st_out, double* sqerr_sum, int32_t* hist,
                       void* workspace, size_t workspace_bytes, void* stream) {
  return rvq_forward_impl(z, N, D, E, E_lp, ee_half, ee_half_bf16, level_meta, K_per, L, mode, idx_out, zq_out, zq_st_out,
                          sqerr_sum, hist, workspace, workspace_bytes, stream, nullptr);
}

int vqb200_rvq_forward_stats(const float* z, int64_t N, int D, const float* E, const uint16_t* E_lp, const float* ee_half,
                             const float* ee_half_bf16, const float* level_meta, int K_per, int L, int mode,
                             int64_t* idx_out, float* zq_out, float* zq_st_out, double* sqerr_sum, int32_t* hist,
                             void* workspace, size_t workspace_bytes, float count_add, double inv_elems, float* ep_usage,
                             float* ep_cnt, float* stats_out, void* stream) {
  VQ_REQUIRE(N > 0 && hist && sqerr_sum && stats_out, VQB200_EINVAL);
  const RvqStatsTail tail{stats_out, ep_usage, ep_cnt, count_add, inv_elems};
  return rvq_forward_impl(z, N, D, E, E_lp, ee_half, ee_half_bf16, level_meta, K_per, L, mode, idx_out, zq_out, zq_st_out,
                          sqerr_sum, hist, workspace, workspace_bytes, stream, &tail);
}

// ---- the training-mode residual forward with a LOCAL EMA update after every level, in ONE call ----
size_t vqb200_rvq_train_workspace_bytes(int64_t N, int K_per, int D, int L, int mode) {
  if (rvq_fused_train_supported(N, K_per, D, L)) return rvq_align(rvq_fused_train_workspace_bytes(N, K_per, D, L));
  const size_t rows = static_cast<size_t>(N > 0 ? N : 0), Kt = static_cast<size_t>(K_per) * L;
  return rvq_align(vqb200_search_workspace_bytes(N, K_per, D, mode)) + 2 * rvq_align(rows * D * 4) +
         rvq_align((Kt * D + Kt) * 4);
}

int vqb200_rvq_train_launches(int64_t N, int K_per, int D, int L, int mode) {
  if (N <= 0 || L < 1) return 0;
  if (rvq_fused_train_supported(N, K_per, D, L)) return 3;               // refresh phase 1, the persistent kernel, refresh phase 2
  return L * (vqb200_search_launches(N, K_per, D, mode) + 3) + 1;      // + gather, scatter-add, EMA finalize; st_loss
}

static int rvq_train_forward_impl(const float* z, int64_t N, int D, float* E, uint16_t* E_lp_planes, float* ee_half,
                                  float* level_meta, int K_per, int L, int mode, float decay, float one_minus_decay,
                                  float eps, float* ema_cluster_size, float* ema_embedding, int64_t* idx_out,
                                  float* zq_out, float* zq_st_out, double* sqerr_sum, int32_t* hist, void* workspace,
                                  size_t workspace_bytes, void* stream, const RvqStatsTail* tail) {
  VQ_REQUIRE(N >= 0 && K_per > 0 && L >= 1 && L <= VQB200_MAX_LEVELS, VQB200_EINVAL);
  if (N == 0) return VQB200_OK;
  VQ_REQUIRE(z && idx_out && E && E_lp_planes && ee_half && level_meta && ema_cluster_size && ema_embedding &&
                 zq_out && workspace, VQB200_EINVAL);
  VQ_REQUIRE(mode == VQB200_MODE_FP32_EXACT || mode == VQB200_MODE_BF16_INPUT, VQB200_EINVAL);
  VQ_REQUIRE(shape_ok(D), VQB200_ESHAPE);
  VQ_REQUIRE(aligned16(z) && aligned16(E) && aligned16(E_lp_planes) && aligned16(zq_out) && aligned16(zq_st_out) &&
                 aligned16(ema_embedding) && (reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, VQB200_EALIGN);
  VQ_REQUIRE(workspace_bytes >= vqb200_rvq_train_workspace_bytes(N, K_per, D, L, mode), VQB200_EWORKSPACE);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (rvq_fused_train_supported(N, K_per, D, L))          // refresh phase 1 -> every level in ONE kernel -> refresh phase 2
    return launch_rvq_fused_train(z, N, D, E, E_lp_planes, ee_half, level_meta, K_per, L, mode, decay, one_minus_decay, eps,
                                  ema_cluster_size, ema_embedding, idx_out, zq_out, zq_st_out, sqerr_sum, hist, workspace,
                                  workspace_bytes, s, tail);
  const bool bf = mode == VQB200_MODE_BF16_INPUT;
  const bool tc = tc_supported(N, K_per, D);
  const int K_total = K_per * L;
  const uint16_t* plane = E_lp_planes + (bf ? 0 : static_cast<size_t>(K_total) * D);   // operand plane of the mode
  const float* ee_bf = ee_half + K_total;
  uint8_t* w = static_cast<uint8_t*>(workspace);
  void* ws_search = w;
  const size_t ws_search_bytes = vqb200_search_workspace_bytes(N, K_per, D, mode);
  w += rvq_align(ws_search_bytes);
  float* res[2];
  res[0] = reinterpret_cast<float*>(w); w += rvq_align(static_cast<size_t>(N) * D * 4);
  res[1] = reinterpret_cast<float*>(w); w += rvq_align(static_cast<size_t>(N) * D * 4);
  float* seg_sum = reinterpret_cast<float*>(w);
  float* seg_cnt = seg_sum + static_cast<size_t>(K_total) * D;
  const size_t seg_bytes = (static_cast<size_t>(K_total) * D + K_total) * 4;

  const float* residual = z;
  for (int l = 0; l < L; ++l) {
    const int64_t s0 = static_cast<int64_t>(l) * K_per;
    int64_t* idx_l = idx_out + static_cast<int64_t>(l) * N;
    const float* meta = level_meta + l * VQB200_LEVEL_META_FLOATS;
    int st = tc ? launch_search_tc(residual, N, D, E + s0 * D, plane + s0 * D, ee_half + s0, ee_bf + s0, meta, K_per, mode,
                                   s0, idx_l, ws_search, ws_search_bytes, s)
                : launch_search_simt(residual, nullptr, N, D, E + s0 * D, (bf ? ee_bf : ee_half) + s0, K_per, bf ? 1 : 0, s0,
                                     idx_l, nullptr, s);
    if (st != VQB200_OK) return st;
    float* nxt = l + 1 < L ? res[l & 1] : nullptr;
    // z_q is gathered BEFORE this level's EMA update moves the codebook (models/vq_vae.py:248 -> :251)
    st = launch_gather(residual, E, idx_l, N, D, K_total, zq_out, l > 0, nullptr, nxt, nullptr, hist, nullptr, s);
    if (st != VQB200_OK) return st;
    cudaError_t e = cudaMemsetAsync(seg_sum, 0, seg_bytes, s);
    if (e != cudaSuccess) return status_of(e);
    st = launch_scatter_add(residual, idx_l, nullptr, N, D, K_total, seg_sum, seg_cnt, s);
    if (st != VQB200_OK) return st;
    st = launch_codebook_refresh(1, seg_sum, seg_cnt, decay, one_minus_decay, eps, K_total, D, K_per, ema_cluster_size,
                                 ema_embedding, E, E_lp_planes, ee_half, level_meta, s);
    if (st != VQB200_OK) return st;
    if (nxt) residual = nxt;
  }
  int st = launch_st_loss(z, zq_out, N * D, zq_st_out, sqerr_sum, s);
  if (st == VQB200_OK && tail && tail->stats_out && hist)
    st = launch_stats_finalize(hist, K_total, tail->count_add, sqerr_sum, tail->inv_elems, tail->ep_usage, tail->ep_cnt,
                               tail->stats_out, s);
  return st;
}

int vqb200_rvq_train_forward(const float* z, int64_t N, int D, float* E, uint16_t* E_lp_planes, float* ee_half,
                             float* level_meta, int K_per, int L, int mode, float decay, float one_minus_decay,
                             float eps, float* ema_cluster_size, float* ema_embedding, int64_t* idx_out,
                             float* zq_out, float* zq_st_out, double* sqerr_sum, int32_t* hist, void* workspace,
                             size_t workspace_bytes, void* stream) {
  return rvq_train_forward_impl(z, N, D, E, E_lp_planes, ee_half, level_meta, K_per, L, mode, decay, one_minus_decay, eps,
                                ema_cluster_size, ema_embedding, idx_out, zq_out, zq_st_out, sqerr_sum, hist, workspace,
                                workspace_bytes, stream, nullptr);
}

int vqb200_rvq_train_forward_stats(const float* z, int64_t N, int D, float* E, uint16_t* E_lp_planes, float* ee_half,
                                   float* level_meta, int K_per, int L, int mode, float decay, float one_minus_decay,
                                   float eps, float* ema_cluster_size, float* ema_embedding, int64_t* idx_out,
                                   float* zq_out, float* zq_st_out, double* sqerr_sum, int32_t* hist, void* workspace,
                                   size_t workspace_bytes, float count_add, double inv_elems, float* ep_usage,
                                   float* ep_cnt, float* stats_out, void* stream) {
  VQ_REQUIRE(N > 0 && hist && sqerr_sum && stats_out, VQB200_EINVAL);
  const RvqStatsTail tail{stats_out, ep_usage, ep_cnt, count_add, inv_elems};
  return rvq_train_forward_impl(z, N, D, E, E_lp_planes, ee_half, level_meta, K_per, L, mode, decay, one_minus_decay, eps,
                                ema_cluster_size, ema_embedding, idx_out, zq_out, zq_st_out, sqerr_sum, hist, workspace,
                                workspace_bytes, stream, &tail);
}

// ---- the same forward in two halves around the point where ranks exchange their segment sums (all-reduced EMA):
// ONE exchange per step instead of one per level -- a level is searched against codes that have only seen decay-only
// updates from the earlier levels of the step (csrc/vq_rvq_fused.cuh), so no search waits for another rank's rows.
int vqb200_rvq_train_fused_supported(int64_t N, int K_per, int D, int L, int mode) {
  (void)mode;
  return rvq_fused_train_supported(N, K_per, D, L) ? 1 : 0;
}

size_t vqb200_rvq_train_begin_workspace_bytes(int64_t N, int K_per, int D, int L, int mode) {
  (void)K_per; (void)L; (void)mode;
  return rvq_align(rvq_fused_workspace_bytes(N, D));
}

int vqb200_rvq_train_begin(const float* z, int64_t N, int D, float* E, uint16_t* E_lp_planes, float* ee_half,
                           float* level_meta, int K_per, int L, int mode, float decay, float one_minus_decay, float eps,
                           float* ema_cluster_size, float* ema_embedding, int64_t* idx_out, float* zq_out,
                           float* zq_st_out, double* sqerr_sum, int32_t* hist, float* seg_sum, float* seg_cnt,
                           void* workspace, size_t workspace_bytes, void* stream) {
  VQ_REQUIRE(N > 0 && K_per > 0 && L >= 1 && L <= VQB200_MAX_LEVELS, VQB200_EINVAL);
  VQ_REQUIRE(z && idx_out && E && E_lp_planes && ee_half && level_meta && ema_cluster_size && ema_embedding && zq_out &&
                 seg_sum && seg_cnt && workspace, VQB200_EINVAL);
  VQ_REQUIRE(mode == VQB200_MODE_FP32_EXACT || mode == VQB200_MODE_BF16_INPUT, VQB200_EINVAL);
  VQ_REQUIRE(rvq_fused_train_supported(N, K_per, D, L), VQB200_ESHAPE);
  VQ_REQUIRE(aligned16(z) && aligned16(E) && aligned16(E_lp_planes) && aligned16(zq_out) && aligned16(zq_st_out) &&
                 aligned16(ema_embedding) && aligned16(seg_sum) && (reinterpret_cast<uintptr_t>(workspace) & 255u) == 0,
             VQB200_EALIGN);
  return launch_rvq_train_begin(z, N, D, E, E_lp_planes, ee_half, level_meta, K_per, L, mode, decay, one_minus_decay, eps,
                                ema_cluster_size, ema_embedding, idx_out, zq_out, zq_st_out, sqerr_sum, hist, seg_sum, seg_cnt,
                                workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int vqb200_rvq_train_finish(const float* seg_sum, const float* seg_cnt, float decay, float one_minus_decay, float eps,
                            int K_per, int L, int D, float* ema_cluster_size, float* ema_embedding, float* E,
                            uint16_t* E_lp_planes, float* ee_half, float* level_meta, void* stream) {
  VQ_REQUIRE(K_per > 0 && L >= 1 && L <= VQB200_MAX_LEVELS, VQB200_EINVAL);
  VQ_REQUIRE(seg_sum && seg_cnt && ema_cluster_size && ema_embedding && E && E_lp_planes && ee_half && level_meta,
             VQB200_EINVAL);
  VQ_REQUIRE(shape_ok(D), VQB200_ESHAPE);
  VQ_REQUIRE(aligned16(seg_sum) && aligned16(ema_embedding) && aligned16(E) && aligned16(E_lp_planes), VQB200_EALIGN);
  return launch_rvq_train_finish(seg_sum, seg_cnt, decay, one_minus_decay, eps, K_per, L, D, ema_cluster_size, ema_embedding,
                                 E, E_lp_planes, ee_half, level_meta, static_cast<cudaStream_t>(stream));
}

// One training level up to the point where ranks exchange their segment sums: search -> gather -> scatter-add.
int vqb200_rvq_train_level(const float* residual, int64_t N, int D, const float* E, const uint16_t* E_lp_planes,
                           const float* ee_half, const float* level_meta, int K_per, int L, int level, int mode,
                           int64_t* idx_out, float* zq_out, float* residual_out, int32_t* hist, float* seg_sum,
                           float* seg_cnt, void* workspace, size_t workspace_bytes, void* stream) {
  VQ_REQUIRE(N > 0 && K_per > 0 && L >= 1 && level >= 0 && level < L, VQB200_EINVAL);
  VQ_REQUIRE(residual && idx_out && E && E_lp_planes && ee_half && level_meta && zq_out && seg_sum && seg_cnt,
             VQB200_EINVAL);
  VQ_REQUIRE(mode == VQB200_MODE_FP32_EXACT || mode == VQB200_MODE_BF16_INPUT, VQB200_EINVAL);
  VQ_REQUIRE(shape_ok(D), VQB200_ESHAPE);
  VQ_REQUIRE(aligned16(residual) && aligned16(E) && aligned16(E_lp_planes) && aligned16(zq_out) &&
                 aligned16(residual_out) && aligned16(seg_sum), VQB200_EALIGN);
  VQ_REQUIRE(workspace_bytes >= vqb200_search_workspace_bytes(N, K_per, D, mode) && workspace, VQB200_EWORKSPACE);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool bf = mode == VQB200_MODE_BF16_INPUT;
  const int K_total = K_per * L;
  const int64_t s0 = static_cast<int64_t>(level) * K_per;
  const uint16_t* plane = E_lp_planes + (bf ? 0 : static_cast<size_t>(K_total) * D);
  const float* ee_bf = ee_half + K_total;
  int st;
  if (tc_supported(N, K_per, D)) {
    VQ_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, VQB200_EALIGN);
    st = launch_search_tc(residual, N, D, E + s0 * D, plane + s0 * D, ee_half + s0, ee_bf + s0,
                          level_meta + level * VQB200_LEVEL_META_FLOATS, K_per, mode, s0, idx_out, workspace,
                          workspace_bytes, s);
  } else {
    st = launch_search_simt(residual, nullptr, N, D, E + s0 * D, (bf ? ee_bf : ee_half) + s0, K_per, bf ? 1 : 0, s0,
                            idx_out, nullptr, s);
  }
  if (st != VQB200_OK) return st;
  st = launch_gather(residual, E, idx_out, N, D, K_total, zq_out, level > 0, nullptr, residual_out, nullptr, hist,
                     nullptr, s);
  if (st != VQB200_OK) return st;
  cudaError_t e = cudaMemsetAsync(seg_sum, 0, static_cast<size_t>(K_total) * D * 4, s);
  if (e == cudaSuccess) e = cudaMemsetAsync(seg_cnt, 0, static_cast<size_t>(K_total) * 4, s);
  if (e != cudaSuccess) return status_of(e);
  return launch_scatter_add(residual, idx_out, nullptr, N, D, K_total, seg_sum, seg_cnt, s);
}

int vqb200_residual_prep(const float* z, const float* E_full, const int64_t* idx, int64_t N, int D, int K_total,
                         int mode, const float* next_level_meta, float* residual_out, uint16_t* z16_out,
                         float* margin_out, void* stream) {
  VQ_REQUIRE(N >= 0 && K_total > 0, VQB200_EINVAL);
  VQ_REQUIRE(N == 0 || (z && E_full && idx && next_level_meta && residual_out && z16_out && margin_out), VQB200_EINVAL);
  VQ_REQUIRE(mode == VQB200_MODE_FP32_EXACT || mode == VQB200_MODE_BF16_INPUT, VQB200_EINVAL);
  VQ_REQUIRE(shape_ok(D), VQB200_ESHAPE);
  VQ_REQUIRE(aligned16(z) && aligned16(E_full) && aligned16(residual_out) && aligned16(z16_out), VQB200_EALIGN);
  return launch_residual_prep(z, E_full, idx, N, D, K_total, mode, next_level_meta, residual_out, z16_out, margin_out,
                              static_cast<cudaStream_t>(stream));
}

int vqb200_search_prepped(const float* z, const uint16_t* z16, const float* margin, int64_t N, int D, const float* E,
                          const uint16_t* E_bf16, const float* ee_half, const float* ee_half_bf16,
                          const float* level_meta, int K, int mode, int64_t idx_offset, int64_t* idx_out,
                          void* workspace, size_t workspace_bytes, void* stream) {
  VQ_REQUIRE(N > 0 && K > 0 && z && z16 && margin && idx_out, VQB200_EINVAL);
  VQ_REQUIRE(E && E_bf16 && ee_half && ee_half_bf16 && level_meta, VQB200_EINVAL);
  VQ_REQUIRE(mode == VQB200_MODE_FP32_EXACT || mode == VQB200_MODE_BF16_INPUT, VQB200_EINVAL);
  VQ_REQUIRE(shape_ok(D) && tc_supported(N, K, D), VQB200_ESHAPE);      // the prepared copy only feeds the tensor path
  VQ_REQUIRE(aligned16(z) && aligned16(E) && aligned16(E_bf16) && (reinterpret_cast<uintptr_t>(z16) & 127u) == 0 &&
                 (reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, VQB200_EALIGN);
  VQ_REQUIRE(workspace_bytes >= vqb200_search_workspace_bytes(N, K, D, mode) && workspace, VQB200_EWORKSPACE);
  PrepArgs prep{z16, margin};
  return launch_search_tc(z, N, D, E, E_bf16, ee_half, ee_half_bf16, level_meta, K, mode, idx_offset, idx_out,
                          workspace, workspace_bytes, static_cast<cudaStream_t>(stream), nullptr, &prep);
}

int vqb200_search(const float* z, int64_t N, int D, const float* E, const uint16_t* E_bf16, const float* ee_half,
                  const float* ee_half_bf16, const float* level_meta, int K, int mode, int64_t idx_offset,
                  int64_t* idx_out, void* workspace, size_t workspace_bytes, void* stream) {
  VQ_REQUIRE(N >= 0 && K > 0, VQB200_EINVAL);
  VQ_REQUIRE(N == 0 || (z && idx_out), VQB200_EINVAL);
  VQ_REQUIRE(E && E_bf16 && ee_half && ee_half_bf16 && level_meta, VQB200_EINVAL);
  VQ_REQUIRE(mode == VQB200_MODE_FP32_EXACT || mode == VQB200_MODE_BF16_INPUT, VQB200_EINVAL);
  VQ_REQUIRE(shape_ok(D), VQB200_ESHAPE);
  VQ_REQUIRE(aligned16(z) && aligned16(E), VQB200_EALIGN);
  VQ_REQUIRE(workspace_bytes >= vqb200_search_workspace_bytes(N, K, D, mode) && (workspace || N == 0),
             VQB200_EWORKSPACE);
  const bool bf = mode == VQB200_MODE_BF16_INPUT;
  if (N == 0) return VQB200_OK;
  if (tc_supported(N, K, D)) {
    VQ_REQUIRE(aligned16(E_bf16) && (reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, VQB200_EALIGN);
    return launch_search_tc(z, N, D, E, E_bf16, ee_half, ee_half_bf16, level_meta, K, mode, idx_offset, idx_out,
                            workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
  }
  timing_mark_begin(static_cast<cudaStream_t>(stream));
  const int st = launch_search_simt(z, nullptr, N, D, E, bf ? ee_half_bf16 : ee_half, K, bf ? 1 : 0, idx_offset, idx_out,
                                    nullptr, static_cast<cudaStream_t>(stream));
  timing_mark_end(static_cast<cudaStream_t>(stream));
  return st;
}

int vqb200_gather(const float* z, const float* E, const int64_t* idx, int64_t N, int D, int K_total, float* zq_out,
                  int zq_accumulate, float* zq_st_out, float* residual_out, double* sqerr_sum, int32_t* hist,
                  const uint8_t* row_mask, void* stream) {
  VQ_REQUIRE(N >= 0 && K_total > 0, VQB200_EINVAL);
  VQ_REQUIRE(N == 0 || (z && E && idx), VQB200_EINVAL);
  VQ_REQUIRE(shape_ok(D), VQB200_ESHAPE);
  VQ_REQUIRE(aligned16(z) && aligned16(E) && aligned16(zq_out) && aligned16(zq_st_out) && aligned16(residual_out),
             VQB200_EALIGN);
  return launch_gather(z, E, idx, N, D, K_total, zq_out, zq_accumulate, zq_st_out, residual_out, sqerr_sum, hist,
                       row_mask, static_cast<cudaStream_t>(stream));
}

int vqb200_quantize_fused_supported(int64_t N, int K, int D, int mode) {
  (void)mode;
  return fused_supported(N, K, D) ? 1 : 0;
}

size_t vqb200_quantize_fused_workspace_bytes(int64_t N, int K, int D, int mode) {
  (void)mode;
  return fused_supported(N, K, D) ? fused_workspace_bytes(N) : 0;
}

int vqb200_quantize_fused(const float* z, int64_t N, int D, const float* E, const uint16_t* E_bf16,
                          const float* ee_half, const float* ee_half_bf16, const float* level_meta, int K, int mode,
                          int64_t idx_offset, int64_t* idx_out, float* zq_out, float* zq_st_out, double* sqerr_sum,
                          int32_t* hist, const uint8_t* row_mask, void* workspace, size_t workspace_bytes, void* stream) {
  VQ_REQUIRE(N > 0 && K > 0 && z && idx_out && workspace, VQB200_EINVAL);
  VQ_REQUIRE(E && E_bf16 && ee_half && ee_half_bf16 && level_meta, VQB200_EINVAL);
  VQ_REQUIRE(mode == VQB200_MODE_FP32_EXACT || mode == VQB200_MODE_BF16_INPUT, VQB200_EINVAL);
  VQ_REQUIRE(fused_supported(N, K, D), VQB200_ESHAPE);
  VQ_REQUIRE(aligned16(z) && aligned16(E) && aligned16(E_bf16) && aligned16(zq_out) && aligned16(zq_st_out) &&
                 (reinterpret_cast<uintptr_t>(workspace) & 255u) == 0,
             VQB200_EALIGN);
  return launch_quantize_fused(z, N, D, E, E_bf16, ee_half, ee_half_bf16, level_meta, K, mode, idx_offset, idx_out,
                               zq_out, zq_st_out, sqerr_sum, hist, row_mask, workspace, workspace_bytes,
                               static_cast<cudaStream_t>(stream));
}

int vqb200_quantize(const float* z, int64_t N, int D, const float* E, const uint16_t* E_bf16, const float* ee_half,
                    const float* ee_half_bf16, const float* level_meta, int K, int mode, int64_t idx_offset,
                    int64_t* idx_out, const float* E_full, int K_total, float* zq_out, float* zq_st_out,
                    double* sqerr_sum, int32_t* hist, const uint8_t* row_mask, void* workspace, size_t workspace_bytes,
                    void* stream) {
  VQ_REQUIRE(N >= 0 && K > 0 && K_total > 0, VQB200_EINVAL);
  if (N == 0) return VQB200_OK;
  VQ_REQUIRE(z && idx_out && E && E_bf16 && ee_half && ee_half_bf16 && level_meta && E_full, VQB200_EINVAL);
  VQ_REQUIRE(mode == VQB200_MODE_FP32_EXACT || mode == VQB200_MODE_BF16_INPUT, VQB200_EINVAL);
  VQ_REQUIRE(shape_ok(D), VQB200_ESHAPE);
  VQ_REQUIRE(aligned16(z) && aligned16(E) && aligned16(E_full) && aligned16(zq_out) && aligned16(zq_st_out), VQB200_EALIGN);
  VQ_REQUIRE(workspace_bytes >= vqb200_search_workspace_bytes(N, K, D, mode) && workspace, VQB200_EWORKSPACE);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool bf = mode == VQB200_MODE_BF16_INPUT;
  if (tc_supported(N, K, D)) {
    VQ_REQUIRE(aligned16(E_bf16) && (reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, VQB200_EALIGN);
    GatherArgs ga{E_full, K_total, zq_out, zq_st_out, sqerr_sum, hist, row_mask};
    return launch_search_tc(z, N, D, E, E_bf16, ee_half, ee_half_bf16, level_meta, K, mode, idx_offset, idx_out, workspace,
                            workspace_bytes, s, &ga);
  }
  timing_mark_begin(s);
  int st = launch_search_simt(z, nullptr, N, D, E, bf ? ee_half_bf16 : ee_half, K, bf ? 1 : 0, idx_offset, idx_out, nullptr, s);
  timing_mark_end(s);
  if (st != VQB200_OK) return st;
  return launch_gather(z, E_full, idx_out, N, D, K_total, zq_out, 0, zq_st_out, nullptr, sqerr_sum, hist, row_mask, s);
}

int vqb200_st_loss(const float* z, const float* zq, int64_t n_elems, float* zq_st_out, double* sqerr_sum,
                   void* stream) {
  VQ_REQUIRE(n_elems >= 0, VQB200_EINVAL);
  VQ_REQUIRE(n_elems == 0 || (z && zq), VQB200_EINVAL);
  VQ_REQUIRE(n_elems % 4 == 0, VQB200_ESHAPE);
  VQ_REQUIRE(aligned16(z) && aligned16(zq) && aligned16(zq_st_out), VQB200_EALIGN);
  return launch_st_loss(z, zq, n_elems, zq_st_out, sqerr_sum, static_cast<cudaStream_t>(stream));
}

int vqb200_stats_finalize(const int32_t* hist, int K_total, float count_add, const double* sqerr_sum,
                          double inv_elems, float* ep_usage, float* ep_cnt, float* stats_out, void* stream) {
  VQ_REQUIRE(hist && stats_out && K_total > 0, VQB200_EINVAL);
  return launch_stats_finalize(hist, K_total, count_add, sqerr_sum, inv_elems, ep_usage, ep_cnt, stats_out,
                               static_cast<cudaStream_t>(stream));
}

int vqb200_stats_pack(const int32_t* hist, int K_total, const double* sqerr_sum, double n_elems, double* packed_out,
                      void* stream) {
  VQ_REQUIRE(hist && packed_out && K_total > 0, VQB200_EINVAL);
  return launch_stats_pack(hist, K_total, sqerr_sum, n_elems, packed_out, static_cast<cudaStream_t>(stream));
}

int vqb200_stats_finalize_packed(const double* packed, int K_total, int levels, int D, float* ep_usage, float* ep_cnt,
                                 float* stats_out, void* stream) {
  VQ_REQUIRE(packed && stats_out && K_total > 0 && levels > 0 && D > 0, VQB200_EINVAL);
  return launch_stats_finalize_packed(packed, K_total, levels, D, ep_usage, ep_cnt, stats_out,
                                      static_cast<cudaStream_t>(stream));
}

size_t vqb200_stats_exchange_buffer_bytes(int K_total, int world) {
  return (K_total > 0 && world > 0) ? stats_exchange_buffer_bytes(K_total, world) : 0;
}

int vqb200_stats_exchange(const int32_t* hist, int K_total, const double* sqerr_sum, double n_elems, int levels, int D,
                          const uint64_t* peer_buffers, int rank, int world, uint64_t spin_limit, float* ep_usage,
                          float* ep_cnt, float* stats_out, void* stream) {
  VQ_REQUIRE(hist && peer_buffers && stats_out && K_total > 0 && levels > 0 && D > 0, VQB200_EINVAL);
  return launch_stats_exchange(hist, K_total, sqerr_sum, n_elems, levels, D, peer_buffers, rank, world,
                               static_cast<unsigned long long>(spin_limit), ep_usage, ep_cnt, stats_out,
                               static_cast<cudaStream_t>(stream));
}

int vqb200_scatter_add(const float* z, const int64_t* idx, const uint8_t* row_mask, int64_t N, int D, int K_total,
                       float* seg_sum, float* seg_cnt, void* stream) {
  VQ_REQUIRE(N >= 0 && K_total > 0 && seg_sum && seg_cnt, VQB200_EINVAL);
  VQ_REQUIRE(N == 0 || (z && idx), VQB200_EINVAL);
  VQ_REQUIRE(shape_ok(D), VQB200_ESHAPE);
  VQ_REQUIRE(aligned16(z) && aligned16(seg_sum), VQB200_EALIGN);
  return launch_scatter_add(z, idx, row_mask, N, D, K_total, seg_sum, seg_cnt, static_cast<cudaStream_t>(stream));
}

int vqb200_ema_finalize(const float* seg_sum, const float* seg_cnt, float decay, float one_minus_decay, float eps,
                        int K_total, int D, int K_per, float* ema_cluster_size, float* ema_embedding, float* E,
                        uint16_t* E_bf16, float* ee_half, float* level_meta, void* stream) {
  VQ_REQUIRE(seg_sum && seg_cnt && ema_cluster_size && ema_embedding && E && E_bf16 && ee_half && level_meta,
             VQB200_EINVAL);
  VQ_REQUIRE(K_total > 0 && K_per > 0, VQB200_EINVAL);
  VQ_REQUIRE(shape_ok(D) && K_total % K_per == 0 && K_total / K_per <= VQB200_MAX_LEVELS, VQB200_ESHAPE);
  VQ_REQUIRE(aligned16(seg_sum) && aligned16(ema_embedding) && aligned16(E) && aligned16(E_bf16), VQB200_EALIGN);
  return launch_codebook_refresh(1, seg_sum, seg_cnt, decay, one_minus_decay, eps, K_total, D, K_per,
                                 ema_cluster_size, ema_embedding, E, E_bf16, ee_half, level_meta,
                                 static_cast<cudaStream_t>(stream));
}

int vqb200_kmeans_finalize(const float* seg_sum, const float* seg_cnt, int K_total, int D, int K_per, float* E,
                           uint16_t* E_bf16, float* ee_half, float* level_meta, void* stream) {
  VQ_REQUIRE(seg_sum && seg_cnt && E && E_bf16 && ee_half && level_meta, VQB200_EINVAL);
  VQ_REQUIRE(K_total > 0 && K_per > 0, VQB200_EINVAL);
  VQ_REQUIRE(shape_ok(D) && K_total % K_per == 0 && K_total / K_per <= VQB200_MAX_LEVELS, VQB200_ESHAPE);
  VQ_REQUIRE(aligned16(seg_sum) && aligned16(E) && aligned16(E_bf16), VQB200_EALIGN);
  return launch_codebook_refresh(2, seg_sum, seg_cnt, 0.f, 0.f, 0.f, K_total, D, K_per, nullptr, nullptr, E, E_bf16,
                                 ee_half, level_meta, static_cast<cudaStream_t>(stream));
}

int vqb200_commit_backward(const float* grad_st, const float* grad_commit, const float* z, const float* zq,
                           int64_t n_elems, float scale, float* grad_z_out, void* stream) {
  VQ_REQUIRE(n_elems >= 0, VQB200_EINVAL);
  VQ_REQUIRE(n_elems == 0 || (z && zq && grad_z_out), VQB200_EINVAL);
  VQ_REQUIRE(n_elems % 4 == 0, VQB200_ESHAPE);
  VQ_REQUIRE(aligned16(grad_st) && aligned16(z) && aligned16(zq) && aligned16(grad_z_out), VQB200_EALIGN);
  return launch_commit_backward(grad_st, grad_commit, z, zq, n_elems, scale, grad_z_out,
                                static_cast<cudaStream_t>(stream));
}

int vqb200_relayout_indices(const int64_t* idx_level_major, int Q, int64_t B, int64_t M, void* out,
                            int out_elem_bytes, void* stream) {
  VQ_REQUIRE(Q > 0 && B >= 0 && M >= 0, VQB200_EINVAL);
  VQ_REQUIRE(B * M == 0 || (idx_level_major && out), VQB200_EINVAL);
  return launch_relayout(idx_level_major, Q, B, M, out, out_elem_bytes, static_cast<cudaStream_t>(stream));
}

int vqb200_indices_to_latent(const void* idx, int idx_elem_bytes, int64_t n_tok, int Q, const float* E, int K_total,
                             int D, float* zq_out, void* stream) {
  VQ_REQUIRE(Q > 0 && n_tok >= 0 && K_total > 0, VQB200_EINVAL);
  VQ_REQUIRE(n_tok == 0 || (idx && E && zq_out), VQB200_EINVAL);
  VQ_REQUIRE(shape_ok(D), VQB200_ESHAPE);
  VQ_REQUIRE(aligned16(E) && aligned16(zq_out), VQB200_EALIGN);
  return launch_indices_to_latent(idx, idx_elem_bytes, n_tok, Q, E, K_total, D, zq_out,
                                  static_cast<cudaStream_t>(stream));
}

size_t vqb200_softmax_rows_workspace_bytes(int64_t N, int K) { return softmax_rows_workspace_bytes(N, K); }

int vqb200_softmax_rows(const float* z, int64_t N, int D, const float* E, const float* beta, int K, float alpha,
                        float* row_stats, float* probs_out, float* p_sum, void* workspace, size_t workspace_bytes,
                        void* stream) {
  VQ_REQUIRE(N >= 0 && K > 0, VQB200_EINVAL);
  VQ_REQUIRE(N == 0 || (z && E && workspace && (probs_out || p_sum || row_stats)), VQB200_EINVAL);
  VQ_REQUIRE(shape_ok(D), VQB200_ESHAPE);
  VQ_REQUIRE(aligned16(z) && aligned16(E) && aligned16(probs_out) && (reinterpret_cast<uintptr_t>(workspace) & 7u) == 0,
             VQB200_EALIGN);
  return launch_softmax_rows(z, N, D, E, beta, K, alpha, row_stats, probs_out, p_sum, workspace, workspace_bytes,
                             static_cast<cudaStream_t>(stream));
}

int vqb200_indices_to_memory(const void* idx, int idx_elem_bytes, int64_t n_tok, int Q, const float* P, int K_total,
                             int H, const float* bias, const float* ln_weight, const float* ln_bias, float ln_eps,
                             float* memory_out, void* stream) {
  VQ_REQUIRE(Q > 0 && n_tok >= 0 && K_total > 0 && H > 0, VQB200_EINVAL);
  VQ_REQUIRE(n_tok == 0 || (idx && P && memory_out), VQB200_EINVAL);
  VQ_REQUIRE(H % 4 == 0 && H <= 1024, VQB200_ESHAPE);
  VQ_REQUIRE(aligned16(P) && aligned16(memory_out) && aligned16(bias) && aligned16(ln_weight) && aligned16(ln_bias),
             VQB200_EALIGN);
  return launch_indices_to_memory(idx, idx_elem_bytes, n_tok, Q, P, K_total, H, bias, ln_weight, ln_bias, ln_eps,
                                  memory_out, static_cast<cudaStream_t>(stream));
}

int vqb200_rvq_finalize(const float* z, const int64_t* idx_level_major, int64_t level_stride, int64_t N, int D, int L,
                        const float* E, int K_total, float* zq_out, float* zq_st_out, double* sqerr_sum, int32_t* hist,
                        void* stream) {
  VQ_REQUIRE(N >= 0 && L >= 1 && K_total > 0 && level_stride >= N, VQB200_EINVAL);
  VQ_REQUIRE(N == 0 || (z && idx_level_major && E), VQB200_EINVAL);
  VQ_REQUIRE(shape_ok(D) && L <= 8, VQB200_ESHAPE);
  VQ_REQUIRE(aligned16(z) && aligned16(E) && aligned16(zq_out) && aligned16(zq_st_out), VQB200_EALIGN);
  return launch_rvq_finalize(z, idx_level_major, level_stride, N, D, L, E, K_total, zq_out, zq_st_out, sqerr_sum, hist,
                             static_cast<cudaStream_t>(stream));
}

int vqb200_usage_probs(const float* z, int64_t N, int D, const float* E, int K, float* p_sum, float* row_stats,
                       void* stream) {
  VQ_REQUIRE(N >= 0 && K > 0, VQB200_EINVAL);
  VQ_REQUIRE(N == 0 || (z && E && p_sum && row_stats), VQB200_EINVAL);
  VQ_REQUIRE(shape_ok(D), VQB200_ESHAPE);
  VQ_REQUIRE(aligned16(z) && aligned16(E) && (reinterpret_cast<uintptr_t>(row_stats) & 7u) == 0, VQB200_EALIGN);
  return launch_usage_probs(z, N, D, E, K, p_sum, row_stats, static_cast<cudaStream_t>(stream));
}

int vqb200_usage_probs_backward(const float* z, int64_t N, int D, const float* E, int K, const float* row_stats,
                                const float* grad_p, float scale, float* grad_z_out, void* stream) {
  VQ_REQUIRE(N >= 0 && K > 0, VQB200_EINVAL);
  VQ_REQUIRE(N == 0 || (z && E && row_stats && grad_p && grad_z_out), VQB200_EINVAL);
  VQ_REQUIRE(shape_ok(D), VQB200_ESHAPE);
  VQ_REQUIRE(aligned16(z) && aligned16(E) && aligned16(grad_z_out) && (reinterpret_cast<uintptr_t>(row_stats) & 7u) == 0,
             VQB200_EALIGN);
  return launch_usage_probs_backward(z, N, D, E, K, row_stats, grad_p, scale, grad_z_out,
                                     static_cast<cudaStream_t>(stream));
}

int vqb200_soft_assign(const float* z, int64_t N, int D, const float* E, int K, float tau, float* z_soft_out,
                       void* stream) {
  VQ_REQUIRE(N >= 0 && K > 0, VQB200_EINVAL);
  VQ_REQUIRE(N == 0 || (z && E && z_soft_out), VQB200_EINVAL);
  VQ_REQUIRE(shape_ok(D), VQB200_ESHAPE);
  VQ_REQUIRE(aligned16(z) && aligned16(E) && aligned16(z_soft_out), VQB200_EALIGN);
  return launch_soft_assign(z, N, D, E, K, tau, z_soft_out, static_cast<cudaStream_t>(stream));
}

int vqb200_search_packed(const float* z, int64_t N, int D, const float* E, const float* ee_half, int K,
                         int64_t idx_offset, uint64_t* packed_out, void* stream) {
  VQ_REQUIRE(N >= 0 && K > 0, VQB200_EINVAL);
  VQ_REQUIRE(N == 0 || (z && packed_out), VQB200_EINVAL);
  VQ_REQUIRE(E && ee_half, VQB200_EINVAL);
  VQ_REQUIRE(shape_ok(D), VQB200_ESHAPE);
  VQ_REQUIRE(aligned16(z) && aligned16(E), VQB200_EALIGN);
  return launch_search_simt(z, nullptr, N, D, E, ee_half, K, 0, idx_offset, nullptr, packed_out,
                            static_cast<cudaStream_t>(stream));
}

int vqb200_pack_exact(const float* z, int64_t N, int D, const float* E_full, int K_total, const int64_t* idx,
                      uint64_t* packed_out, void* stream) {
  VQ_REQUIRE(N >= 0 && K_total > 0 && K_total <= (1 << 24), VQB200_EINVAL);
  VQ_REQUIRE(N == 0 || (z && E_full && idx && packed_out), VQB200_EINVAL);
  VQ_REQUIRE(shape_ok(D), VQB200_ESHAPE);
  VQ_REQUIRE(aligned16(z) && aligned16(E_full), VQB200_EALIGN);
  return launch_pack_exact(z, N, D, E_full, K_total, idx, packed_out, static_cast<cudaStream_t>(stream));
}

int vqb200_minloc_unpack24(const uint64_t* packed, int64_t N, int64_t* idx_out, void* stream) {
  VQ_REQUIRE(N >= 0, VQB200_EINVAL);
  VQ_REQUIRE(N == 0 || (packed && idx_out), VQB200_EINVAL);
  return launch_minloc_unpack24(packed, N, idx_out, static_cast<cudaStream_t>(stream));
}

int vqb200_minloc_unpack(const uint64_t* packed, int64_t N, int64_t* idx_out, void* stream) {
  VQ_REQUIRE(N >= 0, VQB200_EINVAL);
  VQ_REQUIRE(N == 0 || (packed && idx_out), VQB200_EINVAL);
  return launch_minloc_unpack(packed, N, idx_out, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
