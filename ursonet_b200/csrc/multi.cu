// Multi-tensor ("job table") versions of the small per-layer kernels: BN fold, weight-operand staging, conv parameter
// gradients.  Round 1 launched ~330 of them per step (one per layer and kind).  Each is tiny, but next to the persistent
// convolution CTAs (384 threads x ~148 registers, > 200 KB shared memory per SM) they barely find room to co-run, so in
// effect they serialised with the convolutions: skipping them shortened the RN-50 step from 13.7 to 12.3 ms
// (gpurun_out/r2k_bench_*.json) although they move < 0.4 GB.  Here ONE launch covers all layers: a device-resident job
// table, blockIdx -> job by binary search over each job's first block.
#include <cuda_bf16.h>

#include "common.cuh"

namespace urso {

__device__ __forceinline__ int find_job(const int32_t* begins, int n_jobs, int block) {
  int lo = 0, hi = n_jobs - 1;      // last job with begins[job] <= block
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (begins[mid] <= block) lo = mid;
    else hi = mid - 1;
  }
  return lo;
}

// ---------------------------------------------------------------------------------------------- BN fold
__global__ void bn_fold_multi_kernel(const urso_bn_job* __restrict__ jobs, float eps) {
  const urso_bn_job j = jobs[blockIdx.y];
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= j.C) return;
  const float b = j.bias != nullptr ? j.bias[c] : 0.f;
  if (j.gamma != nullptr) {
    const float s = j.gamma[c] * rsqrtf(j.var[c] + eps);
    j.scale[c] = s;
    j.shift[c] = (b - j.mean[c]) * s + j.beta[c];
  } else {
    j.scale[c] = 1.f;
    j.shift[c] = b;
  }
}

// ---------------------------------------------------------------------------------------------- weight staging
// kind 0 (fprop operand, a TRANSPOSE): out[co, k] = w[idx[k], co] * scale[co].  Block = one 32 (k) x 32 (co) tile through
//   shared memory: reads coalesced along co, writes coalesced along k (the per-layer kernel read w with stride CO).
// kind 1 (dgrad operand, a gather of taps): out[ci, slot*COp + co] = w[(tap[slot]*CI + ci), co] * scale[co].  Block = 2048
//   consecutive elements of the [ci][k] space (co contiguous on both sides).
__global__ void __launch_bounds__(256) stage_weights_multi_kernel(const urso_stage_job* __restrict__ jobs,
                                                                  const int32_t* __restrict__ begins, int n_jobs) {
  __shared__ float tile[32][33];
  const int jid = find_job(begins, n_jobs, (int)blockIdx.x);
  const urso_stage_job j = jobs[jid];
  const int b = (int)blockIdx.x - j.block_begin;
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(j.out);
  if (j.kind == 0) {
    const int tiles_k = (j.K + 31) >> 5;
    const int k0 = (b % tiles_k) << 5, co0 = (b / tiles_k) << 5;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int kk = ty + 8 * r, k = k0 + kk, co = co0 + tx;
      float v = 0.f;
      if (k < j.K && co < j.CO) {
        const int src = j.index[k];
        if (src >= 0) v = __ldg(j.w + (long long)src * j.CO + co) * (j.scale != nullptr ? j.scale[co] : 1.f);
      }
      tile[kk][tx] = v;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int cc = ty + 8 * r, co = co0 + cc, k = k0 + tx;
      if (co < j.rows_out && k < j.K) {
        const float v = tile[tx][cc];
        __nv_bfloat16 hi = __float2bfloat16(v);
        if (j.part == 1) hi = __float2bfloat16(v - __bfloat162float(hi));   // low half of the split-bf16 pair
        out[(long long)co * j.ld_out + k] = hi;
      }
    }
  } else {
    const int Kc = j.K * j.COp;                       // K = n_slots
    const long long total = (long long)j.rows_out * Kc;
    const long long base = (long long)b * 2048;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const long long i = base + r * 256 + threadIdx.x;
      if (i >= total) break;
      const int k = (int)(i % Kc), ci = (int)(i / Kc);
      const int slot = k / j.COp, co = k - slot * j.COp;
      float v = 0.f;
      const int t = j.index[slot];
      if (t >= 0 && co < j.CO && ci < j.CI)
        v = __ldg(j.w + ((long long)t * j.CI + ci) * j.CO + co) * (j.scale != nullptr ? j.scale[co] : 1.f);
      out[(long long)ci * j.ld_out + k] = __float2bfloat16(v);
    }
  }
}

// ---------------------------------------------------------------------------------------------- conv parameter gradients
// dW[r,co] = scale[co] * G[grow(r),co] and S[co] += sum_r W[r,co] * G[grow(r),co]; block = 32 channels x 8 row lanes over a
// slab of rows (see conv_param_grads_kernel in params.cu); then the per-channel finalisation.
__global__ void __launch_bounds__(256) pgrad_multi_kernel(const urso_pgrad_job* __restrict__ jobs,
                                                          const int32_t* __restrict__ begins, int n_jobs) {
  __shared__ float red[8][33];
  const int jid = find_job(begins, n_jobs, (int)blockIdx.x);
  const urso_pgrad_job j = jobs[jid];
  const int b = (int)blockIdx.x - j.block_begin;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int co = (b % j.cblocks) * 32 + tx;
  const int r0 = (b / j.cblocks) * j.rows_per_slab;
  const int r1 = min(j.R, r0 + j.rows_per_slab);
  const float sc = (co < j.CO && j.scale != nullptr) ? j.scale[co] : 1.f;
  float s = 0.f;
  if (co < j.CO) {
    for (int r = r0 + ty; r < r1; r += 8) {
      const int gr = j.g_row_map != nullptr ? j.g_row_map[r] : r;
      const float g = j.G[(long long)gr * j.CO + co];
      s += j.w[(long long)r * j.CO + co] * g;
      j.dW[(long long)r * j.CO + co] = sc * g;
    }
  }
  if (j.gamma == nullptr) return;
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && co < j.CO) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][tx];
    atomicAdd(j.S + co, t);
  }
}

// dgamma = rstd * (S + (bias - mean) * colsum) ; dbeta = colsum ; dbias = scale * colsum
__global__ void pgrad_finalize_multi_kernel(const urso_pgrad_job* __restrict__ jobs, float eps) {
  const urso_pgrad_job j = jobs[blockIdx.y];
  const int co = blockIdx.x * blockDim.x + threadIdx.x;
  if (co >= j.CO) return;
  const float cs = j.colsum != nullptr ? j.colsum[co] : 0.f;
  const float sc = j.scale != nullptr ? j.scale[co] : 1.f;
  if (j.dbias != nullptr) j.dbias[co] = sc * cs;
  if (j.gamma != nullptr) {
    const float rstd = rsqrtf(j.var[co] + eps);
    const float bb = j.bias != nullptr ? j.bias[co] : 0.f;
    j.dgamma[co] = rstd * (j.S[co] + (bb - j.mean[co]) * cs);
    j.dbeta[co] = cs;
  }
}

}  // namespace urso

extern "C" {

int urso_sizeof_bn_job(void) { return (int)sizeof(urso_bn_job); }
int urso_sizeof_stage_job(void) { return (int)sizeof(urso_stage_job); }
int urso_sizeof_pgrad_job(void) { return (int)sizeof(urso_pgrad_job); }

int urso_bn_fold_multi(const urso_bn_job* jobs_dev, int32_t n_jobs, int32_t max_c, float eps, void* stream) {
  URSO_REQUIRE(jobs_dev != nullptr && n_jobs >= 1 && max_c >= 1, "bad job table");
  dim3 grid((max_c + 127) / 128, n_jobs);
  urso::bn_fold_multi_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(jobs_dev, eps);
  URSO_CUDA_OK(cudaGetLastError());
  return 0;
}

/* fills block_begin of every job (host table) and begins_out[n]; returns the total block count */
int32_t urso_stage_jobs_finalize(urso_stage_job* jobs_host, int32_t n, int32_t* begins_out) {
  int32_t total = 0;
  for (int i = 0; i < n; ++i) {
    urso_stage_job& j = jobs_host[i];
    j.block_begin = total;
    begins_out[i] = total;
    if (j.kind == 0) total += ((j.K + 31) / 32) * ((j.rows_out + 31) / 32);
    else total += (int32_t)(((long long)j.rows_out * j.K * j.COp + 2047) / 2048);
  }
  return total;
}

int urso_stage_weights_multi(const urso_stage_job* jobs_dev, const int32_t* begins_dev, int32_t n_jobs, int32_t total_blocks,
                             void* stream) {
  URSO_REQUIRE(jobs_dev != nullptr && begins_dev != nullptr && n_jobs >= 1 && total_blocks >= 1, "bad job table");
  urso::stage_weights_multi_kernel<<<total_blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(jobs_dev, begins_dev, n_jobs);
  URSO_CUDA_OK(cudaGetLastError());
  return 0;
}

int32_t urso_pgrad_jobs_finalize(urso_pgrad_job* jobs_host, int32_t n, int32_t* begins_out) {
  int32_t total = 0;
  for (int i = 0; i < n; ++i) {
    urso_pgrad_job& j = jobs_host[i];
    j.cblocks = (j.CO + 31) / 32;
    int slabs = (j.R + 63) / 64;                 // >= 64 rows per block: 8 row lanes x 8 iterations
    if (slabs < 1) slabs = 1;
    j.rows_per_slab = (j.R + slabs - 1) / slabs;
    slabs = (j.R + j.rows_per_slab - 1) / j.rows_per_slab;
    j.block_begin = total;
    begins_out[i] = total;
    total += j.cblocks * slabs;
  }
  return total;
}

int urso_conv_param_grads_multi(const urso_pgrad_job* jobs_dev, const int32_t* begins_dev, int32_t n_jobs,
                                int32_t total_blocks, int32_t max_co, float eps, void* stream) {
  URSO_REQUIRE(jobs_dev != nullptr && begins_dev != nullptr && n_jobs >= 1 && total_blocks >= 1, "bad job table");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  urso::pgrad_multi_kernel<<<total_blocks, 256, 0, st>>>(jobs_dev, begins_dev, n_jobs);
  dim3 grid((max_co + 127) / 128, n_jobs);
  urso::pgrad_finalize_multi_kernel<<<grid, 128, 0, st>>>(jobs_dev, eps);
  URSO_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
