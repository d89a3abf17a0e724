// Parameter plumbing: BN folding, bf16 weight staging for the GEMM engines, conv parameter gradients from the
// raw wgrad, and the fused regulariser + global-norm clip + SGD / AMSGrad update over the flat fp32 arenas.
#include <cuda_bf16.h>

#include "common.cuh"

namespace urso {

__device__ __forceinline__ float warp_sum_p(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// scale = gamma * rsqrt(var + eps) ; shift = (bias - mean) * scale + beta.  NULL gamma => no BN (scale 1, shift bias).
__global__ void bn_fold_kernel(const float* gamma, const float* beta, const float* mean, const float* var,
                               const float* bias, float eps, float* scale, float* shift, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float b = bias != nullptr ? bias[c] : 0.f;
  if (gamma != nullptr) {
    const float s = gamma[c] * rsqrtf(var[c] + eps);
    scale[c] = s;
    shift[c] = (b - mean[c]) * s + beta[c];
  } else {
    scale[c] = 1.f;
    shift[c] = b;
  }
}

// Forward operand: out[row, k] = w[idx[k] * CO + row] * scale[row]  (idx[k] = flat (tap, ci) of the HWIO kernel or -1).
// Threads run along k (contiguous in out); the strided read of w is served by L2 (weights are small).
__global__ void stage_weight_rows_kernel(const float* __restrict__ w, const float* __restrict__ scale,
                                         __nv_bfloat16* __restrict__ out, const int* __restrict__ idx, int K, int CO,
                                         int rows_out, long long ld_out, int part) {
  const long long total = (long long)rows_out * K;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    const int row = (int)(i / K);
    float v = 0.f;
    const int src = idx[k];
    if (src >= 0 && row < CO) v = w[(long long)src * CO + row] * (scale != nullptr ? scale[row] : 1.f);
    __nv_bfloat16 hi = __float2bfloat16(v);
    if (part == 1) hi = __float2bfloat16(v - __bfloat162float(hi));   // low half of the split-bf16 pair
    out[(long long)row * ld_out + k] = hi;
  }
}

// Dgrad operand: out[ci, slot*COp + co] = w[(tap[slot]*CI + ci)*CO + co] * scale[co]   (tap[slot] = -1 => zeros).
__global__ void stage_weight_cols_kernel(const float* __restrict__ w, const float* __restrict__ scale,
                                         __nv_bfloat16* __restrict__ out, const int* __restrict__ tap, int n_slots,
                                         int CI, int CO, int COp, int rows_out, long long ld_out) {
  const int K = n_slots * COp;
  const long long total = (long long)rows_out * K;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    const int ci = (int)(i / K);
    const int slot = k / COp, co = k % COp;
    float v = 0.f;
    const int t = tap[slot];
    if (t >= 0 && co < CO && ci < CI) v = w[((long long)t * CI + ci) * CO + co] * (scale != nullptr ? scale[co] : 1.f);
    out[(long long)ci * ld_out + k] = __float2bfloat16(v);
  }
}

// dW[r,co] = scale[co] * G[grow(r),co] for a slab of rows, and S[co] += sum_{r in slab} W[r,co] * G[grow(r),co]
// (block = 32 channels x 8 row lanes, grid = channel blocks x row slabs: enough CTAs to fill the chip even for CO = 64).
__global__ void __launch_bounds__(256) conv_param_grads_kernel(
    const float* __restrict__ G, const int* __restrict__ g_row_map, const float* __restrict__ w,
    const float* __restrict__ scale, float* __restrict__ dW, float* __restrict__ S, int R, int CO, int rows_per_slab,
    int need_s) {
  __shared__ float red[8][33];
  const int co = blockIdx.x * 32 + threadIdx.x;
  const int r0 = blockIdx.y * rows_per_slab;
  const int r1 = min(R, r0 + rows_per_slab);
  const float sc = (co < CO && scale != nullptr) ? scale[co] : 1.f;
  float s = 0.f;
  if (co < CO) {
    for (int r = r0 + threadIdx.y; r < r1; r += 8) {
      const int gr = g_row_map != nullptr ? g_row_map[r] : r;
      const float g = G[(long long)gr * CO + co];
      s += w[(long long)r * CO + co] * g;
      dW[(long long)r * CO + co] = sc * g;
    }
  }
  if (!need_s) return;
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && co < CO) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += red[j][threadIdx.x];
    atomicAdd(S + co, t);
  }
}

// dgamma = rstd * (S + (bias - mean) * colsum) ; dbeta = colsum ; dbias = scale * colsum
__global__ void bn_param_finalize_kernel(const float* __restrict__ S, const float* __restrict__ colsum,
                                         const float* __restrict__ scale, const float* __restrict__ gamma,
                                         const float* __restrict__ mean, const float* __restrict__ var,
                                         const float* __restrict__ bias, float eps, float* __restrict__ dbias,
                                         float* __restrict__ dgamma, float* __restrict__ dbeta, int CO) {
  const int co = blockIdx.x * blockDim.x + threadIdx.x;
  if (co >= CO) return;
  const float cs = colsum != nullptr ? colsum[co] : 0.f;
  const float sc = scale != nullptr ? scale[co] : 1.f;
  if (dbias != nullptr) dbias[co] = sc * cs;
  if (gamma != nullptr) {
    const float rstd = rsqrtf(var[co] + eps);
    const float b = bias != nullptr ? bias[co] : 0.f;
    dgamma[co] = rstd * (S[co] + (b - mean[co]) * cs);
    dbeta[co] = cs;
  }
}

// g <- (trainable ? g * grad_scale + coef * p : 0) ; sumsq += sum g^2.   chunk = 256 elements (arena alignment).
// 16-byte accesses: a 256-thread block covers 4 chunks per iteration, thread t owns elements [4t, 4t+4) of that span
// (n is a multiple of 256 by construction of the arenas; a ragged tail falls back to scalars).
__global__ void __launch_bounds__(256) add_reg_sumsq_kernel(float* __restrict__ grad, const float* __restrict__ param,
                                                            const float* __restrict__ chunk_coef,
                                                            const float* __restrict__ chunk_lr, float grad_scale,
                                                            float* __restrict__ sumsq, long long n) {
  __shared__ float sh[8];
  float acc = 0.f;
  const long long nspans = (n + 1023) / 1024;
  for (long long sp = blockIdx.x; sp < nspans; sp += gridDim.x) {
    const long long i = sp * 1024 + 4 * threadIdx.x;
    const long long ch = i >> 8;
    if (i + 4 <= n) {
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      if (chunk_lr[ch] != 0.f) {
        const float4 g0 = *reinterpret_cast<const float4*>(grad + i);
        const float4 p0 = __ldg(reinterpret_cast<const float4*>(param + i));
        const float c = chunk_coef[ch];
        g = make_float4(g0.x * grad_scale + c * p0.x, g0.y * grad_scale + c * p0.y, g0.z * grad_scale + c * p0.z,
                        g0.w * grad_scale + c * p0.w);
      }
      *reinterpret_cast<float4*>(grad + i) = g;
      acc += g.x * g.x + g.y * g.y + g.z * g.z + g.w * g.w;
    } else {
      for (long long k = i; k < n; ++k) {
        float g = 0.f;
        if (chunk_lr[k >> 8] != 0.f) g = grad[k] * grad_scale + chunk_coef[k >> 8] * param[k];
        grad[k] = g;
        acc += g * g;
      }
    }
  }
  acc = warp_sum_p(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    float t = sh[threadIdx.x];
    t += __shfl_xor_sync(0xffu, t, 4);
    t += __shfl_xor_sync(0xffu, t, 2);
    t += __shfl_xor_sync(0xffu, t, 1);
    // per-block partial, summed in a FIXED order by sumsq_finish_kernel: with atomics the last bits of the norm (hence of
    // the clip factor, hence of every weight) depended on arrival order and data-parallel ranks drifted apart by ulps
    if (threadIdx.x == 0) sumsq[1 + blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(256) sumsq_finish_kernel(float* __restrict__ sumsq, int nparts) {
  __shared__ float sh[256];
  float t = 0.f;
  for (int i = threadIdx.x; i < nparts; i += 256) t += sumsq[1 + i];
  sh[threadIdx.x] = t;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) sumsq[0] = sh[0];
}

// hyper: [0]=lr  [1]=momentum|beta1  [2]=beta2  [3]=eps  [4]=clipnorm   (device memory: CUDA-graph replays see updates)
__device__ __forceinline__ float clip_factor(const float* sumsq, const float* hyper) {
  const float norm = sqrtf(sumsq[0]);
  const float c = hyper[4];
  return (c > 0.f && norm >= c) ? c / norm : 1.f;
}

__global__ void __launch_bounds__(256) sgd_step_kernel(float* __restrict__ param, float* __restrict__ vel,
                                                       const float* __restrict__ grad,
                                                       const float* __restrict__ chunk_lr,
                                                       const float* __restrict__ sumsq, const float* __restrict__ hyper,
                                                       long long n) {
  const float cf = clip_factor(sumsq, hyper);
  const float lr = hyper[0], mom = hyper[1];
  const float lrc = lr * cf;
  const long long nspans = (n + 1023) / 1024;
  for (long long sp = blockIdx.x; sp < nspans; sp += gridDim.x) {
    const long long i = sp * 1024 + 4 * threadIdx.x;
    if (i + 4 <= n) {
      if (chunk_lr[i >> 8] == 0.f) continue;
      const float4 g = __ldg(reinterpret_cast<const float4*>(grad + i));
      float4 v = *reinterpret_cast<const float4*>(vel + i);
      float4 w = *reinterpret_cast<const float4*>(param + i);
      v = make_float4(mom * v.x - lrc * g.x, mom * v.y - lrc * g.y, mom * v.z - lrc * g.z, mom * v.w - lrc * g.w);
      w = make_float4(w.x + v.x, w.y + v.y, w.z + v.z, w.w + v.w);
      *reinterpret_cast<float4*>(vel + i) = v;
      *reinterpret_cast<float4*>(param + i) = w;
    } else {
      for (long long k = i; k < n; ++k) {
        if (chunk_lr[k >> 8] == 0.f) continue;
        const float v = mom * vel[k] - lrc * grad[k];
        vel[k] = v;
        param[k] += v;
      }
    }
  }
}

__global__ void __launch_bounds__(256) amsgrad_step_kernel(float* __restrict__ param, float* __restrict__ m,
                                                           float* __restrict__ v, float* __restrict__ vhat,
                                                           const float* __restrict__ grad,
                                                           const float* __restrict__ chunk_lr,
                                                           const float* __restrict__ sumsq,
                                                           const float* __restrict__ hyper, long long n) {
  const float cf = clip_factor(sumsq, hyper);
  const float lr_t = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3];
  const long long nspans = (n + 1023) / 1024;
  for (long long sp = blockIdx.x; sp < nspans; sp += gridDim.x) {
    const long long i0 = sp * 1024 + 4 * threadIdx.x;
    if (i0 >= n || chunk_lr[i0 >> 8] == 0.f) continue;
    if (i0 + 4 <= n) {
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(grad + i0));
      float4 m4 = *reinterpret_cast<const float4*>(m + i0);
      float4 v4 = *reinterpret_cast<const float4*>(v + i0);
      float4 h4 = *reinterpret_cast<const float4*>(vhat + i0);
      float4 w4 = *reinterpret_cast<const float4*>(param + i0);
      float* gm = reinterpret_cast<float*>(&m4);
      float* gv = reinterpret_cast<float*>(&v4);
      float* gh = reinterpret_cast<float*>(&h4);
      float* gw = reinterpret_cast<float*>(&w4);
      const float* gg = reinterpret_cast<const float*>(&g4);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float g = gg[k] * cf;
        gm[k] = b1 * gm[k] + (1.f - b1) * g;
        gv[k] = b2 * gv[k] + (1.f - b2) * g * g;
        gh[k] = fmaxf(gh[k], gv[k]);
        gw[k] -= lr_t * gm[k] / (sqrtf(gh[k]) + eps);
      }
      *reinterpret_cast<float4*>(m + i0) = m4;
      *reinterpret_cast<float4*>(v + i0) = v4;
      *reinterpret_cast<float4*>(vhat + i0) = h4;
      *reinterpret_cast<float4*>(param + i0) = w4;
    } else {
      for (long long i = i0; i < n; ++i) {
        const float g = grad[i] * cf;
        const float mi = b1 * m[i] + (1.f - b1) * g;
        const float vi = b2 * v[i] + (1.f - b2) * g * g;
        const float vh = fmaxf(vhat[i], vi);
        m[i] = mi;
        v[i] = vi;
        vhat[i] = vh;
        param[i] -= lr_t * mi / (sqrtf(vh) + eps);
      }
    }
  }
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = __float2bfloat16(x[i]);
}
__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = __bfloat162float(x[i]);
}
__global__ void pad_cast_rows_kernel(const float* __restrict__ src, const float* __restrict__ src2,
                                     __nv_bfloat16* __restrict__ dst, long long rows, int C, int Cpad) {
  const long long total = rows * Cpad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cpad);
    const long long r = i / Cpad;
    float v = 0.f;
    if (c < C) v = src[r * C + c] + (src2 != nullptr ? src2[r * C + c] : 0.f);
    dst[i] = __float2bfloat16(v);
  }
}

// out[c] += sum_rows x[r, c] (bf16 [rows, C]).  Block = 32 channel-pairs x 8 row lanes, grid-strided over row slabs.
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out,
                                                          long long rows, int C, int rows_per_block) {
  __shared__ float red[8][65];
  const int c2 = blockIdx.x * 32 + threadIdx.x;  // channel pair index
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  float a0 = 0.f, a1 = 0.f;
  if (2 * c2 < C) {
    for (long long r = r0 + threadIdx.y; r < r1; r += 8) {
      const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(x + r * C + 2 * c2);
      a0 += __low2float(v);
      a1 += __high2float(v);
    }
  }
  red[threadIdx.y][2 * threadIdx.x] = a0;
  red[threadIdx.y][2 * threadIdx.x + 1] = a1;
  __syncthreads();
  if (threadIdx.y == 0 && 2 * c2 < C) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s0 += red[j][2 * threadIdx.x];
      s1 += red[j][2 * threadIdx.x + 1];
    }
    atomicAdd(out + 2 * c2, s0);
    atomicAdd(out + 2 * c2 + 1, s1);
  }
}

// v = y (+ addend) ; relu ; out32 = v ; hi = bf16(v) ; lo = bf16(v - hi)      (split-bf16 parity mode)
__global__ void split_f32_kernel(const float* __restrict__ y, const float* __restrict__ addend, float* __restrict__ out32,
                                 __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, long long n, int relu) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = y[i] + (addend != nullptr ? addend[i] : 0.f);
    if (relu) v = fmaxf(v, 0.f);
    if (out32 != nullptr) out32[i] = v;
    const __nv_bfloat16 h = __float2bfloat16(v);
    hi[i] = h;
    lo[i] = __float2bfloat16(v - __bfloat162float(h));
  }
}

// Gradient accumulation over micro-batches: acc = beta * acc + grad; out = alpha * acc (if out).  16-byte vectors.
__global__ void grad_accumulate_kernel(float* acc, const float* grad, float* out,   // out may alias grad
                                       float beta, float alpha, long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 g = reinterpret_cast<const float4*>(grad)[i];
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (beta != 0.f) a = reinterpret_cast<const float4*>(acc)[i];
    a.x = beta * a.x + g.x; a.y = beta * a.y + g.y; a.z = beta * a.z + g.z; a.w = beta * a.w + g.w;
    if (out != nullptr) reinterpret_cast<float4*>(out)[i] = make_float4(alpha * a.x, alpha * a.y, alpha * a.z, alpha * a.w);
    else reinterpret_cast<float4*>(acc)[i] = a;
  }
}

static inline int grid_for_p(long long n, int block, int max_blocks) {
  long long g = (n + block - 1) / block;
  if (g > max_blocks) g = max_blocks;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace urso

using namespace urso;

extern "C" {

int urso_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, const float* bias,
                 float eps, float* scale, float* shift, int32_t C, void* stream) {
  URSO_REQUIRE(scale && shift, "null pointer");
  URSO_REQUIRE(gamma == nullptr || (beta && mean && var), "BN needs gamma, beta, mean and var");
  bn_fold_kernel<<<(C + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(gamma, beta, mean, var, bias, eps,
                                                                                scale, shift, C);
  URSO_CUDA_OK(cudaGetLastError());
  return 0;
}

int urso_stage_weight_rows(const float* w, const float* scale, void* out, const int32_t* idx_dev, int32_t K, int32_t CO,
                           int32_t rows_out, int64_t ld_out, int32_t part, void* stream) {
  URSO_REQUIRE(w && out && idx_dev, "null pointer");
  URSO_REQUIRE(ld_out >= K, "ld_out < K");
  stage_weight_rows_kernel<<<grid_for_p((long long)rows_out * K, 256, num_sms() * 8), 256, 0,
                             static_cast<cudaStream_t>(stream)>>>(w, scale, static_cast<__nv_bfloat16*>(out), idx_dev,
                                                                  K, CO, rows_out, ld_out, part);
  URSO_CUDA_OK(cudaGetLastError());
  return 0;
}

int urso_stage_weight_cols(const float* w, const float* scale, void* out, const int32_t* tap_dev, int32_t n_slots,
                           int32_t CI, int32_t CO, int32_t COp, int32_t rows_out, int64_t ld_out, void* stream) {
  URSO_REQUIRE(w && out && tap_dev, "null pointer");
  URSO_REQUIRE(ld_out >= (int64_t)n_slots * COp, "ld_out < K");
  stage_weight_cols_kernel<<<grid_for_p((long long)rows_out * n_slots * COp, 256, num_sms() * 8), 256, 0,
                             static_cast<cudaStream_t>(stream)>>>(w, scale, static_cast<__nv_bfloat16*>(out), tap_dev,
                                                                  n_slots, CI, CO, COp, rows_out, ld_out);
  URSO_CUDA_OK(cudaGetLastError());
  return 0;
}

int urso_conv_param_grads(const float* G, const int32_t* g_row_map_dev, const float* w, const float* colsum,
                          const float* scale, const float* gamma, const float* mean, const float* var,
                          const float* bias, float eps, float* dW, float* dbias, float* dgamma, float* dbeta,
                          float* s_scratch, int32_t R, int32_t CO, void* stream) {
  URSO_REQUIRE(G && w && dW, "null pointer");
  URSO_REQUIRE(gamma == nullptr || (mean && var && dgamma && dbeta && colsum && s_scratch),
               "BN gradients need mean/var/colsum and a zeroed [CO] scratch");
  URSO_REQUIRE(dbias == nullptr || colsum != nullptr, "bias gradient needs colsum");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int cblocks = (CO + 31) / 32;
  int sms = num_sms();
  if (sms <= 0) sms = 148;
  int slabs = (2 * sms + cblocks - 1) / cblocks;
  if (slabs > (R + 7) / 8) slabs = (R + 7) / 8;
  if (slabs < 1) slabs = 1;
  const int rps = (R + slabs - 1) / slabs;
  dim3 block(32, 8), grid(cblocks, (R + rps - 1) / rps);
  conv_param_grads_kernel<<<grid, block, 0, st>>>(G, g_row_map_dev, w, scale, dW, s_scratch, R, CO, rps,
                                                  gamma != nullptr ? 1 : 0);
  if (gamma != nullptr || dbias != nullptr)
    bn_param_finalize_kernel<<<(CO + 127) / 128, 128, 0, st>>>(s_scratch, colsum, scale, gamma, mean, var, bias, eps,
                                                              dbias, dgamma, dbeta, CO);
  URSO_CUDA_OK(cudaGetLastError());
  return 0;
}

int urso_grad_accumulate(float* acc, const float* grad, float* out, float beta, float alpha, int64_t n, void* stream) {
  URSO_REQUIRE(acc && grad, "null pointer");
  URSO_REQUIRE(n % 4 == 0, "n must be a multiple of 4 (the arenas are 256-element aligned)");
  grad_accumulate_kernel<<<grid_for_p(n / 4, 256, num_sms() * 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      acc, grad, out, beta, alpha, n / 4);
  URSO_CUDA_OK(cudaGetLastError());
  return 0;
}

int urso_add_reg_sumsq(float* grad, const float* param, const float* chunk_coef, const float* chunk_lr,
                       float grad_scale, float* sumsq_out, int64_t n, void* stream) {
  URSO_REQUIRE(grad && param && chunk_coef && chunk_lr && sumsq_out, "null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int nblocks = grid_for_p((n + 1023) / 1024, 1, num_sms() * 8);
  if (nblocks > URSO_SUMSQ_SCRATCH - 1) nblocks = URSO_SUMSQ_SCRATCH - 1;
  add_reg_sumsq_kernel<<<nblocks, 256, 0, s>>>(grad, param, chunk_coef, chunk_lr, grad_scale, sumsq_out, n);
  sumsq_finish_kernel<<<1, 256, 0, s>>>(sumsq_out, nblocks);
  URSO_CUDA_OK(cudaGetLastError());
  return 0;
}

int urso_sgd_step(float* param, float* vel, const float* grad, const float* chunk_lr, const float* sumsq,
                  const float* hyper_dev, int64_t n, void* stream) {
  URSO_REQUIRE(param && vel && grad && chunk_lr && sumsq && hyper_dev, "null pointer");
  sgd_step_kernel<<<grid_for_p((n + 1023) / 1024, 1, num_sms() * 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      param, vel, grad, chunk_lr, sumsq, hyper_dev, n);
  URSO_CUDA_OK(cudaGetLastError());
  return 0;
}

int urso_amsgrad_step(float* param, float* m, float* v, float* vhat, const float* grad, const float* chunk_lr,
                      const float* sumsq, const float* hyper_dev, int64_t n, void* stream) {
  URSO_REQUIRE(param && m && v && vhat && grad && chunk_lr && sumsq && hyper_dev, "null pointer");
  amsgrad_step_kernel<<<grid_for_p((n + 1023) / 1024, 1, num_sms() * 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      param, m, v, vhat, grad, chunk_lr, sumsq, hyper_dev, n);
  URSO_CUDA_OK(cudaGetLastError());
  return 0;
}

int urso_split_f32(const float* y, const float* addend, float* out32, void* hi, void* lo, int64_t n, int32_t relu,
                   void* stream) {
  URSO_REQUIRE(y && hi && lo, "null pointer");
  split_f32_kernel<<<grid_for_p(n, 256, num_sms() * 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      y, addend, out32, static_cast<__nv_bfloat16*>(hi), static_cast<__nv_bfloat16*>(lo), n, relu);
  URSO_CUDA_OK(cudaGetLastError());
  return 0;
}

int urso_cast_f32_to_bf16(const float* x, void* y, int64_t n, void* stream) {
  URSO_REQUIRE(x && y, "null pointer");
  cast_f32_bf16_kernel<<<grid_for_p(n, 256, num_sms() * 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, static_cast<__nv_bfloat16*>(y), n);
  URSO_CUDA_OK(cudaGetLastError());
  return 0;
}

int urso_cast_bf16_to_f32(const void* x, float* y, int64_t n, void* stream) {
  URSO_REQUIRE(x && y, "null pointer");
  cast_bf16_f32_kernel<<<grid_for_p(n, 256, num_sms() * 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), y, n);
  URSO_CUDA_OK(cudaGetLastError());
  return 0;
}

int urso_pad_cast_rows(const float* src, const float* src2, void* dst, int64_t rows, int32_t C, int32_t Cpad,
                       void* stream) {
  URSO_REQUIRE(src && dst && Cpad >= C, "bad arguments");
  pad_cast_rows_kernel<<<grid_for_p(rows * Cpad, 256, num_sms() * 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, src2, static_cast<__nv_bfloat16*>(dst), rows, C, Cpad);
  URSO_CUDA_OK(cudaGetLastError());
  return 0;
}

int urso_colsum_bf16(const void* x, float* out, int64_t rows, int32_t C, void* stream) {
  URSO_REQUIRE(x && out && C % 2 == 0, "bad arguments");
  int slabs = (int)((rows + 511) / 512);
  int max_slabs = num_sms() * 4;
  if (max_slabs < 1) max_slabs = 592;
  if (slabs > max_slabs) slabs = max_slabs;
  if (slabs < 1) slabs = 1;
  const int rpb = (int)((rows + slabs - 1) / slabs);
  dim3 grid((C / 2 + 31) / 32, slabs), block(32, 8);
  colsum_bf16_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(x), out,
                                                                            rows, C, rpb);
  URSO_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
