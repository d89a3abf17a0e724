// Dense layers of the two heads (net.py:302,316,336,345,350): fp32, batch <= a few dozen rows, so every kernel is bound
// by streaming the weight matrix W[K, N] (14.6 M parameters at cfg2, 56 MB) through the SMs exactly once.
//
// Round 1 ran these at 270-450 GB/s (4-7 % of HBM): one 4-byte load per thread per loop iteration with the FMAs
// depending on it (latency bound), 16 resident warps per SM.  Here every block stages a [32 or 64 x 256] tile of W
// (32-64 KB) and the matching slice of the small operand in shared memory with 16-byte cp.async copies that are ALL in
// flight at once, then computes from shared memory with 3 LDS.128 per 32 FMAs; grids are (K tiles x N tiles) = several
// hundred blocks, 2-3 resident per SM, so the chip holds > 10 MB of loads in flight (HBM latency x bandwidth ~ 6 MB).
//   fwd   y[b, n] += sum_k x[b,k] w[k,n]            split over K tiles, vector atomics (red.global.add.v4.f32)
//   dgrad dx[b, k] += sum_n dy[b,n] w[k,n]          split over N tiles, 8-lane shuffle reduction + atomics
//   wgrad dw[k, n]  = sum_b x[b,k] dy[b,n]          direct float4 stores
#include "common.cuh"

namespace urso {

constexpr int kDT = 256;          // threads per block
constexpr int kDN = 256;          // columns of W per tile
constexpr int kDKF = 64;          // rows of W per tile: forward
constexpr int kDKB = 32;          // rows of W per tile: dgrad / wgrad
constexpr int kDB = 32;           // batch rows per pass

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
  const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  const int sz = pred ? 16 : 0;   // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void red_add_v4f(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Stage rows [k0, k0 + rows) x columns [n0, n0 + 256) of W[K, N] into ws[rows][256] (zero beyond the matrix).
// vec: N % 4 == 0 and 16-byte aligned rows -> cp.async; otherwise scalar loads (the 3- and 4-column final layers).
__device__ __forceinline__ void stage_w_tile(float* ws, const float* __restrict__ w, int K, int N, int k0, int n0, int rows,
                                             bool vec) {
  if (vec) {
    for (int i = threadIdx.x; i < rows * (kDN / 4); i += kDT) {
      const int r = i / (kDN / 4), q = i % (kDN / 4);
      const bool ok = (k0 + r < K) && (n0 + 4 * q < N);
      cp_async16(ws + r * kDN + 4 * q, ok ? w + (long long)(k0 + r) * N + n0 + 4 * q : w, ok);
    }
  } else {
    for (int i = threadIdx.x; i < rows * kDN; i += kDT) {
      const int r = i / kDN, c = i % kDN;
      ws[i] = (k0 + r < K && n0 + c < N) ? __ldg(w + (long long)(k0 + r) * N + n0 + c) : 0.f;
    }
  }
}

// ---- forward: block = 64 k-rows x 256 columns, batch rows in passes of 32.  Thread: 4 columns x 8 batch rows.
__global__ void __launch_bounds__(kDT) dense_fwd2_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         float* __restrict__ y, int B, int K, int N, int vec) {
  extern __shared__ __align__(16) float dsm[];
  float* ws = dsm;                      // [64][256]
  float* xs = dsm + kDKF * kDN;         // [64][32]  (k-major: the 8 batch values a thread needs are 2 float4)
  const int n0 = blockIdx.x * kDN, k0 = blockIdx.y * kDKF;
  stage_w_tile(ws, w, K, N, k0, n0, kDKF, vec != 0);
  const int nq = threadIdx.x & 63, bo = threadIdx.x >> 6;     // column quad, batch octet
  for (int b0 = 0; b0 < B; b0 += kDB) {
    __syncthreads();                    // previous pass has finished reading xs
    for (int i = threadIdx.x; i < kDKF * kDB; i += kDT) {
      const int kk = i % kDKF, bb = i / kDKF;                 // consecutive threads read consecutive k: coalesced
      xs[kk * kDB + bb] = (k0 + kk < K && b0 + bb < B) ? __ldg(x + (long long)(b0 + bb) * K + k0 + kk) : 0.f;
    }
    cp_async_wait_all();
    __syncthreads();
    float acc[8][4];
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[b][0] = acc[b][1] = acc[b][2] = acc[b][3] = 0.f;
#pragma unroll 8
    for (int kk = 0; kk < kDKF; ++kk) {
      const float4 wv = *reinterpret_cast<const float4*>(ws + kk * kDN + 4 * nq);
      const float4 xa = *reinterpret_cast<const float4*>(xs + kk * kDB + 8 * bo);
      const float4 xb = *reinterpret_cast<const float4*>(xs + kk * kDB + 8 * bo + 4);
      const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        acc[b][0] += xv[b] * wv.x;
        acc[b][1] += xv[b] * wv.y;
        acc[b][2] += xv[b] * wv.z;
        acc[b][3] += xv[b] * wv.w;
      }
    }
    const int n = n0 + 4 * nq;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const int bb = b0 + 8 * bo + b;
      if (bb >= B || n >= N) continue;
      float* dst = y + (long long)bb * N + n;
      if (vec) red_add_v4f(dst, acc[b][0], acc[b][1], acc[b][2], acc[b][3]);
      else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n + j < N) atomicAdd(dst + j, acc[b][j]);
      }
    }
  }
}

// ---- dgrad: block = 32 k-rows x 256 columns.  Thread: one k-row, 8 column quads (seg + 8 i), 32 batch rows; the 8 threads of
// a k-row are adjacent lanes and are reduced with shuffles.
__global__ void __launch_bounds__(kDT) dense_dgrad2_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                           float* __restrict__ dx, int B, int K, int N, int vec) {
  extern __shared__ __align__(16) float dsm[];
  float* ws = dsm;                      // [32][256]
  float* gs = dsm + kDKB * kDN;         // [32 b][256]
  const int n0 = blockIdx.x * kDN, k0 = blockIdx.y * kDKB;
  stage_w_tile(ws, w, K, N, k0, n0, kDKB, vec != 0);
  const int kr = threadIdx.x >> 3, seg = threadIdx.x & 7;
  for (int b0 = 0; b0 < B; b0 += kDB) {
    __syncthreads();
    for (int i = threadIdx.x; i < kDB * kDN; i += kDT) {
      const int bb = i / kDN, c = i % kDN;
      gs[i] = (b0 + bb < B && n0 + c < N) ? __ldg(dy + (long long)(b0 + bb) * N + n0 + c) : 0.f;
    }
    cp_async_wait_all();
    __syncthreads();
    float acc[kDB];
#pragma unroll
    for (int b = 0; b < kDB; ++b) acc[b] = 0.f;
#pragma unroll 2
    for (int i = 0; i < 8; ++i) {
      const int q = seg + 8 * i;
      const float4 wv = *reinterpret_cast<const float4*>(ws + kr * kDN + 4 * q);
#pragma unroll
      for (int b = 0; b < kDB; ++b) {
        const float4 g = *reinterpret_cast<const float4*>(gs + b * kDN + 4 * q);
        acc[b] += g.x * wv.x + g.y * wv.y + g.z * wv.z + g.w * wv.w;
      }
    }
#pragma unroll
    for (int b = 0; b < kDB; ++b) {
      float v = acc[b];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      acc[b] = v;
    }
    if (seg == 0 && k0 + kr < K) {
#pragma unroll
      for (int b = 0; b < kDB; ++b)
        if (b0 + b < B) atomicAdd(dx + (long long)(b0 + b) * K + k0 + kr, acc[b]);
    }
  }
}

// ---- wgrad: block = 32 k-rows x 256 columns of dW.  Thread: 4 columns x 8 k-rows, loops over the batch.
__global__ void __launch_bounds__(kDT) dense_wgrad2_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                           float* __restrict__ dw, int B, int K, int N) {
  __shared__ __align__(16) float gs[kDB * kDN];     // [32 b][256]
  __shared__ __align__(16) float xs[kDB * kDKB];    // [32 b][32 k]
  const int n0 = blockIdx.x * kDN, k0 = blockIdx.y * kDKB;
  const int nq = threadIdx.x & 63, ko = threadIdx.x >> 6;     // column quad, k octet
  float acc[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
  for (int b0 = 0; b0 < B; b0 += kDB) {
    __syncthreads();
    for (int i = threadIdx.x; i < kDB * kDN; i += kDT) {
      const int bb = i / kDN, c = i % kDN;
      gs[i] = (b0 + bb < B && n0 + c < N) ? __ldg(dy + (long long)(b0 + bb) * N + n0 + c) : 0.f;
    }
    for (int i = threadIdx.x; i < kDB * kDKB; i += kDT) {
      const int bb = i / kDKB, kk = i % kDKB;
      xs[i] = (b0 + bb < B && k0 + kk < K) ? __ldg(x + (long long)(b0 + bb) * K + k0 + kk) : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int bb = 0; bb < kDB; ++bb) {
      const float4 g = *reinterpret_cast<const float4*>(gs + bb * kDN + 4 * nq);
      const float4 xa = *reinterpret_cast<const float4*>(xs + bb * kDKB + 8 * ko);
      const float4 xb = *reinterpret_cast<const float4*>(xs + bb * kDKB + 8 * ko + 4);
      const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[j][0] += xv[j] * g.x;
        acc[j][1] += xv[j] * g.y;
        acc[j][2] += xv[j] * g.z;
        acc[j][3] += xv[j] * g.w;
      }
    }
  }
  const int n = n0 + 4 * nq;
  if (n >= N) return;
  const bool vec = (N % 4 == 0);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = k0 + 8 * ko + j;
    if (k >= K) continue;
    float* dst = dw + (long long)k * N + n;
    if (vec) *reinterpret_cast<float4*>(dst) = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
    else {
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (n + t < N) dst[t] = acc[j][t];
    }
  }
}

}  // namespace urso

// host-side launchers used by the C-ABI entry points in layers.cu
namespace urso {
int dense_fwd2(const float* x, const float* w, float* y, int B, int K, int N, cudaStream_t s) {
  static bool attr = false;
  const int smem = (kDKF * kDN + kDKF * kDB) * 4;
  if (!attr) {
    URSO_CUDA_OK(cudaFuncSetAttribute(dense_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  const int vec = (N % 4 == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0) ? 1 : 0;
  dim3 grid((N + kDN - 1) / kDN, (K + kDKF - 1) / kDKF);
  dense_fwd2_kernel<<<grid, kDT, smem, s>>>(x, w, y, B, K, N, vec);
  URSO_CUDA_OK(cudaGetLastError());
  return 0;
}
int dense_dgrad2(const float* dy, const float* w, float* dx, int B, int K, int N, cudaStream_t s) {
  static bool attr = false;
  const int smem = (kDKB * kDN + kDB * kDN) * 4;
  if (!attr) {
    URSO_CUDA_OK(cudaFuncSetAttribute(dense_dgrad2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  const int vec = (N % 4 == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0) ? 1 : 0;
  URSO_CUDA_OK(cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)B * K, s));
  dim3 grid((N + kDN - 1) / kDN, (K + kDKB - 1) / kDKB);
  dense_dgrad2_kernel<<<grid, kDT, smem, s>>>(dy, w, dx, B, K, N, vec);
  URSO_CUDA_OK(cudaGetLastError());
  return 0;
}
int dense_wgrad2(const float* x, const float* dy, float* dw, int B, int K, int N, cudaStream_t s) {
  dim3 grid((N + kDN - 1) / kDN, (K + kDKB - 1) / kDKB);
  dense_wgrad2_kernel<<<grid, kDT, 0, s>>>(x, dy, dw, B, K, N);
  URSO_CUDA_OK(cudaGetLastError());
  return 0;
}
}  // namespace urso
