// Which feature of a tcgen05 kernel pins the runtime's occupancy to one CTA per SM?  Tiny kernels, one feature
// each; prints cudaOccupancyMaxActiveBlocksPerMultiprocessor and, for the TMEM kernel, measured co-residency
// (CTAs that overlapped in time on the same SM).   nvcc -gencode arch=compute_100a,code=sm_100a occ_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__global__ void __launch_bounds__(192, 2) k_plain(int* out) {
  extern __shared__ uint8_t sm[];
  sm[threadIdx.x] = 1;
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = sm[5];
}

__global__ void __launch_bounds__(192, 2) k_mbar(int* out, int off) {
  extern __shared__ uint8_t sm[];
  const uint32_t bar = smem_u32(sm) + off * 8;      // dynamic address: ptxas cannot count the barriers
  if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1) : "memory");
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = 1;
}

template <int COLS>
__global__ void __launch_bounds__(192, 2) k_tmem(long long* t0, long long* t1, int* smid, int spin) {
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  long long a = clock64();
  if (threadIdx.x == 0) {
    unsigned id; asm volatile("mov.u32 %0, %%smid;" : "=r"(id));
    long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
    t0[blockIdx.x] = g; smid[blockIdx.x] = id;
  }
  while (clock64() - a < spin) { }
  __syncthreads();
  if (threadIdx.x == 0) { long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); t1[blockIdx.x] = g; }
  if (threadIdx.x < 32) {
    const uint32_t base = slot;
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(COLS) : "memory");
  }
}

template <typename K> static void occ(const char* name, K k, int threads, int smem) {
  int nb = -1;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000);
  cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, threads, smem);
  printf("%-28s threads %d smem %6d -> %d CTA/SM (%s)\n", name, threads, smem, nb, cudaGetErrorString(e));
}

template <int COLS> static void coreside(int grid) {
  long long *t0, *t1; int* smid;
  cudaMallocManaged(&t0, grid * 8); cudaMallocManaged(&t1, grid * 8); cudaMallocManaged(&smid, grid * 4);
  k_tmem<COLS><<<grid, 192>>>(t0, t1, smid, 2000000);
  cudaError_t e = cudaDeviceSynchronize();
  int overlap = 0;
  for (int i = 0; i < grid; ++i)
    for (int j = i + 1; j < grid; ++j)
      if (smid[i] == smid[j] && t0[i] < t1[j] && t0[j] < t1[i]) ++overlap;
  printf("k_tmem<%d> grid %d: %s, pairs of CTAs overlapping in time on one SM: %d\n", COLS, grid, cudaGetErrorString(e), overlap);
}

int main() {
  occ("plain", k_plain, 192, 50000);
  occ("mbarrier (dynamic addr)", k_mbar, 192, 50000);
  occ("tcgen05.alloc 256 cols", k_tmem<256>, 192, 0);
  occ("tcgen05.alloc 128 cols", k_tmem<128>, 192, 0);
  occ("tcgen05.alloc 512 cols", k_tmem<512>, 192, 0);
  coreside<256>(296);
  coreside<128>(592);
  coreside<512>(296);
  return 0;
}
