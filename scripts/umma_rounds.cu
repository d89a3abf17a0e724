// Microbenchmark (round 2): what does one BARRIER ROUND of the MMA-issuing warp cost, and can a second issuing warp hide it?
// A round = wait on an (already completed) mbarrier, tcgen05.fence::after_thread_sync, elect, `n` tcgen05.mma (K = 16 each,
// M = 128, N = 64/128/256, operands in shared memory), one tcgen05.commit.  Variants:
//   issuers = 1 | 2   : one warp, or two warps (different SM sub-partitions) on DIFFERENT accumulators, each running the loop
//   flags & 1         : skip the fence
//   flags & 2         : skip the wait
//   flags & 4         : skip the per-round commit (one commit at the end)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/umma_rounds scripts/umma_rounds.cu
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../ursonet_b200/csrc/ptx.cuh"
using namespace urso;

template <int N>
__global__ void __launch_bounds__(256, 1) rounds_kernel(long long* out, int iters, int n_mma, int issuers, int flags) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t ring[2][16];
  __shared__ uint64_t bar[2];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  // 4 A tiles (16 KB each) + 4 B tiles, so that consecutive MMAs read different smem (as in the real kernel)
  for (int i = threadIdx.x; i < (4 * 16384 + 4 * N * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    for (int w = 0; w < 2; ++w) {
      mbar_init(&bar[w], 1);
      for (int i = 0; i < 16; ++i) mbar_init(&ring[w][i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const int me = warp == 0 ? 0 : (warp == 2 ? 1 : -1);     // warps 0 and 2: different sub-partitions
  long long t0 = 0, t1 = 0;
  if (me >= 0 && me < issuers) {
    constexpr uint64_t hi = (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 16) | (1ull << 46) | (2ull << 61);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 4 * 16384;
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
    const uint32_t d = tmem + me * 256;
    const int depth = 8;
    t0 = clock64();
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
      if (!(flags & 2) && it >= depth) mbar_wait(&ring[me][stage], phase ^ 1);
      if (!(flags & 1)) tc_fence_after();
      if (elect_one()) {
        for (int j = 0; j < n_mma; j += 4) {
          const uint64_t ad = hi | ((a0 + ((j >> 2) & 3) * 16384) >> 4), bd = hi | ((b0 + ((j >> 2) & 3) * N * 128) >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(d, ad + 2 * k, bd + 2 * k, idesc, 1u);
        }
        if (!(flags & 4)) umma_commit(&ring[me][stage]);
      }
      __syncwarp();
      if (++stage == depth) {
        stage = 0;
        phase ^= 1;
      }
    }
    if (elect_one()) umma_commit(&bar[me]);
    __syncwarp();
    mbar_wait(&bar[me], 0);
    t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[blockIdx.x * 2 + me] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <int N>
static void run(int n_mma, int issuers, int flags) {
  long long* d;
  const int grid = 148;
  cudaMalloc(&d, grid * 2 * sizeof(long long));
  cudaMemset(d, 0, grid * 2 * sizeof(long long));
  const int iters = 2048, smem = 4 * 16384 + 4 * N * 128 + 1024;
  cudaFuncSetAttribute(rounds_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) rounds_kernel<N><<<grid, 256, smem>>>(d, iters, n_mma, issuers, flags);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[296];
  cudaMemcpy(h, d, grid * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < grid * 2; ++i) mx = h[i] > mx ? h[i] : mx;
  const double per_round = (double)mx / iters;           // wall cycles per round of ONE warp
  const double per_mma = per_round / (n_mma * issuers);  // aggregate cycles per MMA on the SM
  const double ideal = (N == 64 ? 48.0 : N / 2.0);
  printf("N=%3d mma/round=%2d issuers=%d flags=%d  %s  cycles/round %7.1f  cycles/MMA(SM) %6.1f  (bare %4.0f)  pipe util %.2f\n", N,
         n_mma, issuers, flags, cudaGetErrorString(e), per_round, per_mma, ideal, ideal / per_mma);
  cudaFree(d);
}

int main() {
  for (int issuers = 1; issuers <= 2; ++issuers) {
    for (int n : {4, 8, 12, 36}) run<64>(n, issuers, 0);
    for (int n : {4, 8, 12}) run<128>(n, issuers, 0);
    for (int n : {4, 8}) run<256>(n, issuers, 0);
  }
  // where does a round's overhead come from (one issuer, N = 64, 4 MMAs per round)
  run<64>(4, 1, 1);   // no fence
  run<64>(4, 1, 2);   // no wait
  run<64>(4, 1, 3);   // no fence, no wait
  run<64>(4, 1, 4);   // no commit
  run<64>(4, 1, 7);   // bare
  run<64>(4, 2, 1);
  run<128>(4, 2, 1);
  return 0;
}
