// Host side of the persistent residual-VQ kernel (vq_rvq_fused.cuh): which shapes it takes, tile shape and workspace, the
// launch, and the training-mode forward around it.
#include "vq_rvq_fused.cuh"

namespace vqb {

// ------------------------------------------------------------------------------------ host side
static int rq_smem_bytes(int BM, int BN, int D, int stages) {
  return 1024 + BM * D * 2 + stages * BN * TC_KB * 2 + rq_misc_bytes(BM, BN) + 8 + 24 * 8 + 64;
}


// Tile shape of a launch.  BM = 64 when the batch has so few 128-row tiles that most of a second wave of SMs would
// idle (VQB200_RVQ_BM=64|128 overrides); BN = 256 wherever the operand tile leaves room for >= 3 ring stages.
static bool rq_config(int64_t N, int D, RqConfig* c) {
  const int64_t tiles128 = (N + 127) / 128;
  int bm = tiles128 * 4 <= kNumSMs * 3 ? 64 : 128;
  if (const char* e = std::getenv("VQB200_RVQ_BM")) {
    if (e[0] == '6') bm = 64;
    else if (e[0] == '1') bm = 128;
  }
  int bn = (bm == 64 || D <= 384) ? 256 : 128;
  int stages = 8;
  while (stages > 3 && rq_smem_bytes(bm, bn, D, stages) > RQ_SMEM_LIMIT) --stages;
  if (rq_smem_bytes(bm, bn, D, stages) > RQ_SMEM_LIMIT) return false;
  const int64_t tiles = (N + bm - 1) / bm;
  c->BM = bm; c->BN = bn; c->stages = stages;
  c->grid = static_cast<int>(tiles < kNumSMs ? tiles : kNumSMs);
  c->smem = rq_smem_bytes(bm, bn, D, stages);
  return true;
}

// Above this many rows the level-by-level pipeline (CTA-pair tensor kernel, HBM-bound passes between levels) is the
// faster one: the persistent kernel serialises search and row passes inside a CTA (measured, profiles/README.md).
// (profiles/r02_rvq_crossover.txt: 4 x 1024 codes; at 65536 rows D = 512: 0.72 vs 0.77 ms, D = 256: 0.57 vs 0.54 ms)
static int64_t rq_max_rows(int D) {
  if (const char* e = std::getenv("VQB200_RVQ_FUSED_MAX_ROWS")) return std::atoll(e);
  return (D == 256 || D == 384) ? 49152 : 65536;
}

bool rvq_fused_supported(int64_t N, int K_per, int D, int L) {
  const char* f = std::getenv("VQB200_FORCE_SIMT");
  if (f && f[0] == '1') return false;
  const char* g = std::getenv("VQB200_NO_RVQ_FUSED");
  if (g && g[0] == '1') return false;
  if (D % 128 != 0 || D < 128 || D > 512 || K_per < TC_BN || L < 2 || L > RQ_MAXL || N < 1 || N > rq_max_rows(D)) return false;
  RqConfig c;
  return rq_config(N, D, &c);
}

bool rvq_fused_train_supported(int64_t N, int K_per, int D, int L) {
  const char* g = std::getenv("VQB200_NO_RVQ_FUSED_TRAIN");
  if (g && g[0] == '1') return false;
  return rvq_fused_supported(N, K_per, D, L);
}

size_t rvq_fused_workspace_bytes(int64_t N, int D) {
  // sized for either tile shape (the environment switch may change between the size query and the launch)
  const int64_t t64 = (N + 63) / 64, t128 = (N + 127) / 128;
  const size_t a = static_cast<size_t>(t64 < kNumSMs ? t64 : kNumSMs) * 64, b = static_cast<size_t>(t128 < kNumSMs ? t128 : kNumSMs) * 128;
  return 256 + (a > b ? a : b) * D * 4;
}

// training workspace: [0, 256) counters | seg_sum | seg_cnt | residual scratch tiles
static size_t rq_train_seg_bytes(int K_total, int D) {
  return (static_cast<size_t>(K_total) * D * 4 + static_cast<size_t>(K_total) * 4 + 255) / 256 * 256;
}
size_t rvq_fused_train_workspace_bytes(int64_t N, int K_per, int D, int L) {
  return rq_train_seg_bytes(K_per * L, D) + rvq_fused_workspace_bytes(N, D);
}

// seg_sum / seg_cnt non-null: training mode (the caller refreshes the codebook before and after, see cabi.cu)
int launch_rvq_fused(const float* z, int64_t N, int D, const float* E, const uint16_t* E_lp, const float* ee_half,
                     const float* level_meta, int K_per, int L, int mode, int64_t* idx_out, float* zq_out,
                     float* zq_st_out, double* sqerr_sum, int32_t* hist, void* workspace, size_t workspace_bytes,
                     cudaStream_t s, float* seg_sum, float* seg_cnt, const RvqStatsTail* tail) {
  if (!rvq_fused_supported(N, K_per, D, L)) return VQB200_ESHAPE;
  if (workspace_bytes < rvq_fused_workspace_bytes(N, D)) return VQB200_EWORKSPACE;
  const bool bf = mode == VQB200_MODE_BF16_INPUT, scatter = seg_sum != nullptr;
  RqConfig c;
  if (!rq_config(N, D, &c)) return VQB200_ESHAPE;
  RvqParams p{};
  p.n_rows = N; p.D = D; p.K_per = K_per; p.L = L; p.mode = mode;
  p.row_tiles = static_cast<int>((N + c.BM - 1) / c.BM);
  p.code_tiles = (K_per + c.BN - 1) / c.BN;
  p.stages = c.stages;
  // kind::f16 instruction descriptor: D = fp32, A = B = bf16 (bf16_input) or fp16, K-major, N = BN, M = BM
  p.idesc = (1u << 4) | (bf ? ((1u << 7) | (1u << 10)) : 0u) | ((static_cast<uint32_t>(c.BN) >> 3) << 17) |
            ((static_cast<uint32_t>(c.BM) >> 4) << 24);
  p.z = z; p.E = E; p.E_lp = E_lp; p.ee_half = ee_half; p.level_meta = level_meta;
  p.seg_sum = seg_sum; p.seg_cnt = seg_cnt;
  if (tail && tail->stats_out && hist) {
    p.stats_out = tail->stats_out; p.ep_usage = tail->ep_usage; p.ep_cnt = tail->ep_cnt;
    p.count_add = tail->count_add; p.inv_elems = tail->inv_elems;
  }
  uint8_t* w = static_cast<uint8_t*>(workspace);
  p.counters = reinterpret_cast<int*>(w);
  p.scratch = reinterpret_cast<float*>(w + 256);
  p.idx_out = idx_out; p.zq_out = zq_out; p.zq_st_out = zq_st_out; p.sqerr_sum = sqerr_sum; p.hist = hist;
  CUtensorMap map_e;
  if (!make_tensor_map_2d(&map_e, E_lp, static_cast<int64_t>(K_per) * L, D, c.BN,
                          bf ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2))
    return VQB200_EDRIVER;
  cudaError_t e = cudaMemsetAsync(p.counters, 0, 4 * sizeof(int), s);
  if (e != cudaSuccess) return status_of(e);
  const char* dbg = std::getenv("VQB200_DEBUG");
  p.trace = nullptr;
  if (dbg && dbg[0] == '5') {
    cudaMallocManaged(&p.trace, 3 * RQ_MAXL * 8 * sizeof(long long));
    cudaMemset(p.trace, 0, 3 * RQ_MAXL * 8 * sizeof(long long));
  }
  int st;
  switch (D / 128) {
    case 1: st = rq_launch_sl1(map_e, p, bf, c, scatter, s); break;
    case 2: st = rq_launch_sl2(map_e, p, bf, c, scatter, s); break;
    case 3: st = rq_launch_sl3(map_e, p, bf, c, scatter, s); break;
    default: st = rq_launch_sl4(map_e, p, bf, c, scatter, s); break;
  }
  if (p.trace) {                                                 // per-role timeline of CTA 0's first tile (cycles)
    cudaStreamSynchronize(s);
    long long t0 = 0;
    for (int i = 0; i < 3 * RQ_MAXL * 8; ++i) if (p.trace[i] && (!t0 || p.trace[i] < t0)) t0 = p.trace[i];
    static const char* names[3] = {"mma ", "scan", "help"};
    fprintf(stderr, "[vqb200] rvq fused%s: BM=%d BN=%d stages=%d grid=%d smem=%d\n", scatter ? " (training)" : "", c.BM, c.BN,
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

// Training-mode forward in two halves around the point where ranks may exchange their segment sums:
//   begin : refresh phase 1 -> the persistent kernel, reducing the residual rows into seg_sum / seg_cnt (zeroed here)
//   finish: refresh phase 2 from the (possibly all-reduced) segment sums
// (see the kernel's header comment for why this is the reference's level-by-level update sequence).
int launch_rvq_train_begin(const float* z, int64_t N, int D, float* E, uint16_t* E_lp_planes, float* ee_half,
                           float* level_meta, int K_per, int L, int mode, float decay, float omd, float eps,
                           float* ema_cs, float* ema_emb, int64_t* idx_out, float* zq_out, float* zq_st_out,
                           double* sqerr_sum, int32_t* hist, float* seg_sum, float* seg_cnt, void* workspace,
                           size_t workspace_bytes, cudaStream_t s, const RvqStatsTail* tail) {
  if (!rvq_fused_train_supported(N, K_per, D, L)) return VQB200_ESHAPE;
  if (workspace_bytes < rvq_fused_workspace_bytes(N, D)) return VQB200_EWORKSPACE;
  const bool bf = mode == VQB200_MODE_BF16_INPUT;
  const int K_total = K_per * L;
  cudaError_t e;
  if (seg_cnt == seg_sum + static_cast<size_t>(K_total) * D) {            // one buffer (the usual case): one memset
    e = cudaMemsetAsync(seg_sum, 0, (static_cast<size_t>(K_total) * D + K_total) * 4, s);
  } else {
    e = cudaMemsetAsync(seg_sum, 0, static_cast<size_t>(K_total) * D * 4, s);
    if (e == cudaSuccess) e = cudaMemsetAsync(seg_cnt, 0, static_cast<size_t>(K_total) * 4, s);
  }
  if (e != cudaSuccess) return status_of(e);
  int st = launch_codebook_refresh(1, seg_sum, seg_cnt, decay, omd, eps, K_total, D, K_per, ema_cs, ema_emb, E, E_lp_planes,
                                   ee_half, level_meta, s, /*chain_phase=*/1);
  if (st != VQB200_OK) return st;
  const uint16_t* plane = E_lp_planes + (bf ? 0 : static_cast<size_t>(K_total) * D);
  return launch_rvq_fused(z, N, D, E, plane, bf ? ee_half + K_total : ee_half, level_meta, K_per, L, mode, idx_out, zq_out,
                          zq_st_out, sqerr_sum, hist, workspace, workspace_bytes, s, seg_sum, seg_cnt, tail);
}

int launch_rvq_train_finish(const float* seg_sum, const float* seg_cnt, float decay, float omd, float eps, int K_per, int L,
                            int D, float* ema_cs, float* ema_emb, float* E, uint16_t* E_lp_planes, float* ee_half,
                            float* level_meta, cudaStream_t s) {
  return launch_codebook_refresh(1, seg_sum, seg_cnt, decay, omd, eps, K_per * L, D, K_per, ema_cs, ema_emb, E, E_lp_planes,
                                 ee_half, level_meta, s, /*chain_phase=*/2);
}

int launch_rvq_fused_train(const float* z, int64_t N, int D, float* E, uint16_t* E_lp_planes, float* ee_half,
                           float* level_meta, int K_per, int L, int mode, float decay, float omd, float eps,
                           float* ema_cs, float* ema_emb, int64_t* idx_out, float* zq_out, float* zq_st_out,
                           double* sqerr_sum, int32_t* hist, void* workspace, size_t workspace_bytes, cudaStream_t s,
                           const RvqStatsTail* tail) {
  if (!rvq_fused_train_supported(N, K_per, D, L)) return VQB200_ESHAPE;
  if (workspace_bytes < rvq_fused_train_workspace_bytes(N, K_per, D, L)) return VQB200_EWORKSPACE;
  const int K_total = K_per * L;
  uint8_t* w = static_cast<uint8_t*>(workspace);
  float* seg_sum = reinterpret_cast<float*>(w);
  float* seg_cnt = seg_sum + static_cast<size_t>(K_total) * D;
  const size_t seg_bytes = rq_train_seg_bytes(K_total, D);
  const int st = launch_rvq_train_begin(z, N, D, E, E_lp_planes, ee_half, level_meta, K_per, L, mode, decay, omd, eps, ema_cs,
                                        ema_emb, idx_out, zq_out, zq_st_out, sqerr_sum, hist, seg_sum, seg_cnt, w + seg_bytes,
                                        workspace_bytes - seg_bytes, s, tail);
  if (st != VQB200_OK) return st;
  return launch_rvq_train_finish(seg_sum, seg_cnt, decay, omd, eps, K_per, L, D, ema_cs, ema_emb, E, E_lp_planes, ee_half,
                                 level_meta, s);
}

}  // namespace vqb
