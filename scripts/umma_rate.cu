// Microbenchmark: issue rate of tcgen05.mma (bf16, M = 128 per CTA, N = 64/128/256, K = 16, both operands in shared memory)
// for cta_group::1 and for cta_group::2 pairs (M = 256 over two CTAs, each holding N/2 rows of B).  No TMA: the operands are
// whatever is in shared memory (zeros), so this isolates the smem -> tensor-core operand path and the MMA pipe.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/umma_rate scripts/umma_rate.cu && scripts/umma_rate
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../ursonet_b200/csrc/ptx.cuh"
using namespace urso;

// mode 1: one commit per K step (4 MMAs) to a ring of barriers; mode 2: additionally wait, before issuing K step i, for the
// commit of K step i - depth (what the full/empty handshake of the conv kernel amounts to without a producer)
template <int N>
__global__ void __launch_bounds__(128, 1) loop_kernel(long long* out, int iters, int mode, int depth) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t ring[16];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (16384 + N * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    for (int i = 0; i < 16; ++i) mbar_init(&ring[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 0) {
    constexpr uint64_t hi = (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 16) | (1ull << 46) | (2ull << 61);
    const uint64_t ad = hi | (smem_u32(smem) >> 4), bd = hi | ((smem_u32(smem) + 16384) >> 4);
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
    const long long t0 = clock64();
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
      if (mode == 2 && it >= depth) mbar_wait(&ring[stage], phase ^ 1);   // commit of step it - depth has arrived
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem, ad + 2 * k, bd + 2 * k, idesc, 1u);
        umma_commit(&ring[stage]);
      }
      __syncwarp();
      if (++stage == depth) {
        stage = 0;
        phase ^= 1;
      }
    }
    if (elect_one()) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// The same loop for a CTA pair (cta_group::2, M = 256): the leader issues 4 MMAs + one multicast commit per K step and waits for
// the commit of `depth` steps ago; the peer only allocates TMEM and waits for the end.  NOT RUN YET (written at the end of
// round 1 to isolate why the pair kernel of Engine F is slower than the single-CTA one: is it the multicast commit?).
template <int N>
__global__ void __launch_bounds__(128, 1) loop_kernel_2cta(long long* out, int iters, int mode, int depth) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t ring[16];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (16384 + N * 64) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    for (int i = 0; i < 16; ++i) mbar_init(&ring[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2cta(&slot, 512);
  fence_proxy_async();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t rank = cluster_ctarank();
  if (warp == 0 && rank == 0) {
    constexpr uint64_t hi = (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 16) | (1ull << 46) | (2ull << 61);
    const uint64_t ad = hi | (smem_u32(smem) >> 4), bd = hi | ((smem_u32(smem) + 16384) >> 4);
    constexpr uint32_t idesc = umma_idesc_bf16(256, N, 0, 0);
    const long long t0 = clock64();
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
      if (mode == 2 && it >= depth) mbar_wait(&ring[stage], phase ^ 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_2cta(tmem, ad + 2 * k, bd + 2 * k, idesc, 1u);
        umma_commit_2cta(&ring[stage]);
      }
      __syncwarp();
      if (++stage == depth) {
        stage = 0;
        phase ^= 1;
      }
    }
    if (elect_one()) umma_commit_2cta(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  } else if (warp == 0) {
    mbar_wait(&bar, 0);      // the leader's final multicast commit
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem, 512);
  }
}

template <int N>
static void run_loop_2cta(const char* name, int mode, int depth) {
  long long* d;
  const int grid = 148;
  cudaMalloc(&d, grid * sizeof(long long));
  cudaMemset(d, 0, grid * sizeof(long long));
  const int iters = 4096, smem = 16384 + N * 64 + 1024;
  cudaFuncSetAttribute(loop_kernel_2cta<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) cudaLaunchKernelEx(&cfg, loop_kernel_2cta<N>, d, iters, mode, depth);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("%-40s %s  cycles per K step (4 pair MMAs) %7.1f\n", name, cudaGetErrorString(e), (double)mx / iters);
  cudaFree(d);
}

template <int N>
static void run_loop(const char* name, int mode, int depth) {
  long long* d;
  const int grid = 148;
  cudaMalloc(&d, grid * sizeof(long long));
  const int iters = 4096, smem = 16384 + N * 128 + 1024;
  cudaFuncSetAttribute(loop_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) loop_kernel<N><<<grid, 128, smem>>>(d, iters, mode, depth);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("%-40s %s  cycles per K step (4 MMAs) %7.1f\n", name, cudaGetErrorString(e), (double)mx / iters);
  cudaFree(d);
}

template <int N, bool CTA2>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int iters) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (16384 + N * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (CTA2) tmem_alloc_2cta(&slot, 512);
    else tmem_alloc(&slot, 512);
  }
  fence_proxy_async();
  tc_fence_before();
  if constexpr (CTA2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t rank = CTA2 ? cluster_ctarank() : 0u;
  long long t0 = 0, t1 = 0;
  if (warp == 0 && rank == 0) {
    constexpr uint64_t hi = (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 16) | (1ull << 46) | (2ull << 61);
    const uint64_t ad = hi | (smem_u32(smem) >> 4), bd = hi | ((smem_u32(smem) + 16384) >> 4);
    constexpr uint32_t idesc = umma_idesc_bf16(CTA2 ? 256 : 128, N, 0, 0);
    t0 = clock64();
    if (elect_one()) {
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if constexpr (CTA2) umma_bf16_2cta(tmem, ad + 2 * k, bd + 2 * k, idesc, 1u);
          else umma_bf16(tmem, ad + 2 * k, bd + 2 * k, idesc, 1u);
        }
      }
      if constexpr (CTA2) umma_commit_2cta(&bar); else umma_commit(&bar);
    }
    __syncwarp();
  }
  if (warp == 0) {
    mbar_wait(&bar, 0);
    t1 = clock64();
    if (rank == 0 && threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  if constexpr (CTA2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CTA2) tmem_dealloc_2cta(tmem, 512); else tmem_dealloc(tmem, 512);
  }
}

template <int N, bool CTA2>
static void run(const char* name, int grid) {
  long long* d;
  cudaMalloc(&d, grid * sizeof(long long));
  cudaMemset(d, 0, grid * sizeof(long long));
  const int iters = 4096, smem = 16384 + N * 128 + 1024;
  cudaFuncSetAttribute(rate_kernel<N, CTA2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTA2 ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) cudaLaunchKernelEx(&cfg, rate_kernel<N, CTA2>, d, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148] = {0};
  cudaMemcpy(h, d, (grid < 148 ? grid : 148) * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < grid && i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  const double per = (double)mx / (iters * 4.0);
  const double flop = 2.0 * (CTA2 ? 256 : 128) * N * 16;     // per instruction (pair or single CTA)
  printf("%-28s grid %3d  %s  cycles/MMA %7.1f  -> %6.1f FLOP/clk per SM (peak 8192)\n", name, grid, cudaGetErrorString(e), per,
         flop / per / (CTA2 ? 2 : 1));
  cudaFree(d);
}

int main() {
  run<64, false>("1-CTA M128 N64", 148);
  run<128, false>("1-CTA M128 N128", 148);
  run<256, false>("1-CTA M128 N256", 148);
  run<256, false>("1-CTA M128 N256 (1 CTA)", 1);
  run<64, true>("pair  M256 N64", 148);
  run<128, true>("pair  M256 N128", 148);
  run<256, true>("pair  M256 N256", 148);
  run<256, true>("pair  M256 N256 (1 pair)", 2);
  run_loop<64>("N64  commit per K step", 1, 8);
  run_loop<64>("N64  commit + wait depth 8", 2, 8);
  run_loop<64>("N64  commit + wait depth 4", 2, 4);
  run_loop<64>("N64  commit + wait depth 2", 2, 2);
  run_loop<128>("N128 commit + wait depth 6", 2, 6);
  run_loop<256>("N256 commit + wait depth 4", 2, 4);
  run_loop_2cta<256>("pair N256 commit per K step", 1, 6);
  run_loop_2cta<256>("pair N256 commit + wait depth 6", 2, 6);
  return 0;
}
