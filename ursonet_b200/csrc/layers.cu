// Bandwidth-bound layers of the path: stem input staging, max-pool, the fp32 Dense heads and the losses.
// All are HBM/L2-bound: 16-byte vector accesses, coalesced along the channel (innermost NHWC) axis, grids sized in
// multiples of the SM count where the problem allows.
#include <cuda_bf16.h>

#include "common.cuh"

namespace urso {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum, result valid in every thread; blockDim.x multiple of 32, <= 1024
__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float t = (l < nw) ? sh[l] : 0.0f;
  return warp_sum(t);
}
__device__ __forceinline__ float block_max(float v, float* sh) {
  v = warp_max(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float t = (l < nw) ? sh[l] : -INFINITY;
  return warp_max(t);
}

// ------------------------------------------------------------------------------------------ stem staging
// Compact space-to-depth staging S[b, h2, w2, ph*8 + pw*4 + c] = img[2*h2+ph-3, 2*w2+pw-3, c] - mean (0 outside the image and
// for c == 3), h2 < H/2+3, w2 < W/2+3: 16 bf16 = 32 bytes per staged pixel.  The 7x7/s2 stem reads it through an OVERLAPPING
// tensor view E[b, h2, wo, 64] with a pixel stride of 16 elements: the 64 "channels" of view pixel wo are the staged pixels
// wo .. wo+3, i.e. the four horizontal taps of the space-to-depth filter, without storing them four times (round 1 did:
// 635 MB instead of 160 MB per step at the bench shape, written once and read by the stem's forward and weight gradient).
// One thread writes one 16-byte group (fixed ph: 8 values).
template <bool U8>
__global__ void stem_stage_kernel(const void* __restrict__ img, int subtract_mean, const float* __restrict__ mean3,
                                  __nv_bfloat16* __restrict__ e, int B, int H, int W, int part) {
  const int H2 = H / 2 + 3, W2 = W / 2 + 3;
  const long long total = (long long)B * H2 * W2 * 2;
  const float m0 = subtract_mean ? mean3[0] : 0.f, m1 = subtract_mean ? mean3[1] : 0.f, m2 = subtract_mean ? mean3[2] : 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ph = (int)(i & 1);
    long long r = i >> 1;
    const int w2 = (int)(r % W2); r /= W2;
    const int h2 = (int)(r % H2);
    const int b = (int)(r / H2);
    const int y = 2 * h2 + ph - 3;
    float v[8];
#pragma unroll
    for (int pw = 0; pw < 2; ++pw) {
      const int x = 2 * w2 + pw - 3;
      const bool in = (y >= 0) && (y < H) && (x >= 0) && (x < W);
      float c0 = 0.f, c1 = 0.f, c2 = 0.f;
      if (in) {
        const long long off = (((long long)b * H + y) * W + x) * 3;
        if (U8) {
          const uint8_t* p = static_cast<const uint8_t*>(img) + off;
          c0 = (float)p[0] - m0; c1 = (float)p[1] - m1; c2 = (float)p[2] - m2;
        } else {
          const float* p = static_cast<const float*>(img) + off;
          c0 = p[0] - m0; c1 = p[1] - m1; c2 = p[2] - m2;
        }
      }
      v[pw * 4 + 0] = c0; v[pw * 4 + 1] = c1; v[pw * 4 + 2] = c2; v[pw * 4 + 3] = 0.f;
    }
    if (part == 1) {   // low half of the split-bf16 pair: v - bf16(v)
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] -= __bfloat162float(__float2bfloat16(v[j]));
    }
    __nv_bfloat162 o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    *reinterpret_cast<uint4*>(e + i * 8) = *reinterpret_cast<uint4*>(o);
  }
}

// fp32 variant of the forward pool (split-bf16 parity mode), one thread per output element group of 4 channels
__global__ void maxpool_fwd_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int H, int W, int C) {
  const int HO = H / 2, WO = W / 2, C4 = C / 4;
  const long long total = (long long)B * HO * WO * C4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    long long r = i / C4;
    const int wo = (int)(r % WO); r /= WO;
    const int ho = (int)(r % HO);
    const int b = (int)(r / HO);
    float4 best = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    for (int dr = 0; dr < 3; ++dr) {
      const int h = 2 * ho + dr;
      if (h >= H) continue;
      for (int ds = 0; ds < 3; ++ds) {
        const int w = 2 * wo + ds;
        if (w >= W) continue;
        const float4 u = *reinterpret_cast<const float4*>(x + (((long long)b * H + h) * W + w) * C + c4 * 4);
        best.x = fmaxf(best.x, u.x); best.y = fmaxf(best.y, u.y); best.z = fmaxf(best.z, u.z); best.w = fmaxf(best.w, u.w);
      }
    }
    *reinterpret_cast<float4*>(y + i * 4) = best;
  }
}

// ------------------------------------------------------------------------------------------ max-pool 3x3 / s2 'same'
// Even H, W: TF pads only bottom/right, so window (ho,wo) covers rows 2ho..2ho+2, cols 2wo..2wo+2 clipped to the map.
// One thread = one output pixel x 8 channels, all arithmetic on packed bf16x2 / packed bytes (branch free):
//   m = (new > best) per half-word, best = max, argmax bytes = select(m, code, arg).
// Out-of-range taps are clamped to the last row / column: a duplicate of an earlier element never wins a strict '>'.
__device__ __forceinline__ uint32_t gt_mask2(uint32_t a, uint32_t b) {
  return __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
}
__device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}

__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                          __nv_bfloat16* __restrict__ y, uint8_t* __restrict__ argmax,
                                                          int B, int H, int W, int C) {
  const int HO = H / 2, WO = W / 2, C8 = C / 8;
  const long long total = (long long)B * HO * WO * C8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    long long r = i / C8;
    const int wo = (int)(r % WO); r /= WO;
    const int ho = (int)(r % HO);
    const int b = (int)(r / HO);
    uint4 v[9];
#pragma unroll
    for (int dr = 0; dr < 3; ++dr) {
      const int h = min(2 * ho + dr, H - 1);
#pragma unroll
      for (int ds = 0; ds < 3; ++ds) {
        const int w = min(2 * wo + ds, W - 1);
        v[dr * 3 + ds] = *reinterpret_cast<const uint4*>(x + (((long long)b * H + h) * W + w) * C + c8 * 8);
      }
    }
    uint4 best = v[0];
    uint32_t arg_lo = 0u, arg_hi = 0u;   // packed argmax bytes of elements 0-3 / 4-7
#pragma unroll
    for (int t = 1; t < 9; ++t) {
      const uint32_t code4 = 0x01010101u * t;
      const uint32_t m0 = gt_mask2(v[t].x, best.x), m1 = gt_mask2(v[t].y, best.y);
      const uint32_t m2 = gt_mask2(v[t].z, best.z), m3 = gt_mask2(v[t].w, best.w);
      best.x = max2(v[t].x, best.x); best.y = max2(v[t].y, best.y);
      best.z = max2(v[t].z, best.z); best.w = max2(v[t].w, best.w);
      const uint32_t bm_lo = __byte_perm(m0, m1, 0x6420), bm_hi = __byte_perm(m2, m3, 0x6420);   // half-word -> byte masks
      arg_lo = (arg_lo & ~bm_lo) | (code4 & bm_lo);
      arg_hi = (arg_hi & ~bm_hi) | (code4 & bm_hi);
    }
    *reinterpret_cast<uint4*>(y + i * 8) = best;
    if (argmax != nullptr) *reinterpret_cast<uint2*>(argmax + i * 8) = make_uint2(arg_lo, arg_hi);
  }
}

// dx[h,w,c] = [x>0] * sum over windows whose recorded argmax is (h,w) of dy.
// One thread = a 2x2 block of input pixels x 8 channels.  The block (2i..2i+1, 2j..2j+1) is touched by exactly the four
// windows (i-1..i) x (j-1..j) and by 9 (window, pixel) combinations whose argmax code is a compile-time constant, so the
// routing is a packed byte compare + mask (branch free).  Optionally accumulates the per-channel sums of dx (d beta of
// the stem's BatchNorm): warp-shuffle pre-reduction, shared-memory accumulation, one global atomic per channel per block.
__device__ __forceinline__ void route(float (&acc)[8], uint2 codes, uint4 g, uint32_t code) {
  const uint32_t c4 = 0x01010101u * code;
  const uint32_t e_lo = __vcmpeq4(codes.x, c4), e_hi = __vcmpeq4(codes.y, c4);          // 0xFF per matching byte
  const uint32_t m0 = __byte_perm(e_lo, 0, 0x1100), m1 = __byte_perm(e_lo, 0, 0x3322);   // byte -> half-word masks
  const uint32_t m2 = __byte_perm(e_hi, 0, 0x1100), m3 = __byte_perm(e_hi, 0, 0x3322);
  const uint32_t g0 = g.x & m0, g1 = g.y & m1, g2 = g.z & m2, g3 = g.w & m3;
  acc[0] += __uint_as_float(g0 << 16); acc[1] += __uint_as_float(g0 & 0xFFFF0000u);
  acc[2] += __uint_as_float(g1 << 16); acc[3] += __uint_as_float(g1 & 0xFFFF0000u);
  acc[4] += __uint_as_float(g2 << 16); acc[5] += __uint_as_float(g2 & 0xFFFF0000u);
  acc[6] += __uint_as_float(g3 << 16); acc[7] += __uint_as_float(g3 & 0xFFFF0000u);
}

__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                          const uint8_t* __restrict__ argmax,
                                                          const __nv_bfloat16* __restrict__ dy,
                                                          __nv_bfloat16* __restrict__ dx, float* __restrict__ colsum,
                                                          int B, int H, int W, int C) {
  __shared__ float cs[512];
  const int HO = H / 2, WO = W / 2, C8 = C / 8;
  if (colsum != nullptr) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) cs[c] = 0.f;
    __syncthreads();
  }
  float lsum[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) lsum[k] = 0.f;
  const long long total = (long long)B * HO * WO * C8;
  // grid stride is a multiple of C8 (host guarantees it), so a thread always handles the same 8 channels
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    long long r = i / C8;
    const int j = (int)(r % WO); r /= WO;
    const int ih = (int)(r % HO);
    const int b = (int)(r / HO);
    // the four windows (ih-1..ih) x (j-1..j); out-of-range ones contribute nothing (code 0xFF never matches)
    uint2 cd[2][2];
    uint4 g[2][2];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int ho = ih - 1 + a, wo = j - 1 + c;
        if (ho >= 0 && wo >= 0) {
          const long long o = ((((long long)b * HO + ho) * WO + wo) * C8 + c8) * 8;
          cd[a][c] = *reinterpret_cast<const uint2*>(argmax + o);
          g[a][c] = *reinterpret_cast<const uint4*>(dy + o);
        } else {
          cd[a][c] = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
          g[a][c] = make_uint4(0u, 0u, 0u, 0u);
        }
      }
    }
    float acc[2][2][8];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[a][c][k] = 0.f;
    // pixel (pa,pc) of the block is element (dr,ds) = (pa + 2 - 2a, pc + 2 - 2c) of window (a,c) when that is in 0..2
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int pa = 0; pa < 2; ++pa)
#pragma unroll
          for (int pc = 0; pc < 2; ++pc) {
            const int dr = pa + 2 - 2 * a, ds = pc + 2 - 2 * c;
            if (dr <= 2 && ds <= 2) route(acc[pa][pc], cd[a][c], g[a][c], dr * 3 + ds);
          }
#pragma unroll
    for (int pa = 0; pa < 2; ++pa) {
#pragma unroll
      for (int pc = 0; pc < 2; ++pc) {
        const long long off = ((((long long)b * H + 2 * ih + pa) * W + 2 * j + pc) * C8 + c8) * 8;
        if (x != nullptr) {
          const uint4 xv = *reinterpret_cast<const uint4*>(x + off);
          const __nv_bfloat16* xe = reinterpret_cast<const __nv_bfloat16*>(&xv);
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[pa][pc][k] = __bfloat162float(xe[k]) > 0.f ? acc[pa][pc][k] : 0.f;
        }
        __nv_bfloat162 o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          o[k] = __floats2bfloat162_rn(acc[pa][pc][2 * k], acc[pa][pc][2 * k + 1]);
          lsum[2 * k] += __low2float(o[k]);
          lsum[2 * k + 1] += __high2float(o[k]);
        }
        *reinterpret_cast<uint4*>(dx + off) = *reinterpret_cast<uint4*>(o);
      }
    }
  }
  if (colsum != nullptr) {
    // lanes l, l+8, l+16, l+24 of a warp hold the same channel group (C8 == 8) -> shuffle-reduce, then smem atomics
    const int lane = threadIdx.x & 31;
    if (C8 == 8) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        lsum[k] += __shfl_xor_sync(0xffffffffu, lsum[k], 8);
        lsum[k] += __shfl_xor_sync(0xffffffffu, lsum[k], 16);
      }
    }
    if (C8 != 8 || lane < 8) {
      const int c8 = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) % C8);
#pragma unroll
      for (int k = 0; k < 8; ++k) atomicAdd(&cs[c8 * 8 + k], lsum[k]);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x)
      if (cs[c] != 0.f) atomicAdd(colsum + c, cs[c]);
  }
}

// ------------------------------------------------------------------------------------------ Dense heads (fp32)
// The three GEMM-shaped kernels (fwd, dgrad, wgrad) live in dense.cu; the elementwise parts stay here.
__global__ void dense_bias_act_kernel(float* __restrict__ y, const float* __restrict__ bias, int B, int N, int act) {
  const long long total = (long long)B * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    float v = y[i] + bias[i % N];
    if (act == 1) v = fmaxf(v, 0.f);
    y[i] = v;
  }
}

// dy <- dy * (y>0) (if relu) ; db[n] = sum_b dy[b,n].  One thread per column.
__global__ void dense_mask_bias_grad_kernel(const float* __restrict__ y, float* __restrict__ dy, float* __restrict__ db,
                                            int B, int N, int act) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int b = 0; b < B; ++b) {
    float g = dy[(long long)b * N + n];
    if (act == 1 && !(y[(long long)b * N + n] > 0.f)) {
      g = 0.f;
      dy[(long long)b * N + n] = 0.f;
    }
    s += g;
  }
  if (db != nullptr) db[n] = s;
}

// ------------------------------------------------------------------------------------------ losses
// tf.losses.softmax_cross_entropy(onehot_labels=y, logits=z): mean over batch.  One block per row.
__global__ void __launch_bounds__(256) softmax_xent_kernel(const float* __restrict__ z, const float* __restrict__ y,
                                                           float* __restrict__ dz, float* __restrict__ loss_out, int B,
                                                           int N, float weight) {
  __shared__ float sh[32];
  const int b = blockIdx.x;
  const float* zr = z + (long long)b * N;
  const float* yr = y + (long long)b * N;
  float m = -INFINITY;
  for (int i = threadIdx.x; i < N; i += blockDim.x) m = fmaxf(m, zr[i]);
  m = block_max(m, sh);
  float se = 0.f, sy = 0.f, syz = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    se += expf(zr[i] - m);
    sy += yr[i];
    syz += yr[i] * zr[i];
  }
  se = block_sum(se, sh);
  sy = block_sum(sy, sh);
  syz = block_sum(syz, sh);
  const float lse = m + logf(se);
  const float scale = weight / (float)B;
  if (dz != nullptr) {
    const float inv = 1.0f / se;
    for (int i = threadIdx.x; i < N; i += blockDim.x) dz[(long long)b * N + i] = scale * (expf(zr[i] - m) * inv * sy - yr[i]);
  }
  if (threadIdx.x == 0) atomicAdd(loss_out, scale * (lse * sy - syz));
}

// tf.norm((gt - pred) / tf.norm(gt)) over the whole tensor (net.py:757).  Single block.
__global__ void __launch_bounds__(256) rel_loss_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                       float* __restrict__ dpred, float* __restrict__ loss_out, int n,
                                                       float weight) {
  __shared__ float sh[32];
  float sd = 0.f, sg = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float d = gt[i] - pred[i];
    sd += d * d;
    sg += gt[i] * gt[i];
  }
  sd = block_sum(sd, sh);
  sg = block_sum(sg, sh);
  const float nd = sqrtf(sd), ng = sqrtf(sg);
  if (dpred != nullptr) {
    const float k = nd > 0.f ? weight / (nd * ng) : 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) dpred[i] = -(gt[i] - pred[i]) * k;
  }
  if (threadIdx.x == 0) loss_out[0] = weight * nd / ng;
}

// q = l2_normalize(raw) ; loss = mean_b (1 - |<gt, q>|).  One thread per row (B is small), single block reduce.
__global__ void __launch_bounds__(256) quat_head_kernel(const float* __restrict__ raw, const float* __restrict__ gt,
                                                        float* __restrict__ q_out, float* __restrict__ draw,
                                                        float* __restrict__ loss_out, int B, float weight) {
  __shared__ float sh[32];
  float part = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    float r[4], g[4], q[4];
    float n2 = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      r[j] = raw[b * 4 + j];
      g[j] = gt != nullptr ? gt[b * 4 + j] : 0.f;
      n2 += r[j] * r[j];
    }
    const float inv = rsqrtf(fmaxf(n2, 1e-12f));
    float d = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      q[j] = r[j] * inv;
      d += g[j] * q[j];
      q_out[b * 4 + j] = q[j];
    }
    part += 1.0f - fabsf(d);
    if (draw != nullptr) {
      const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
      const float k = -sgn * weight / (float)B;
      float dq[4], qdq = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) { dq[j] = k * g[j]; qdq += q[j] * dq[j]; }
#pragma unroll
      for (int j = 0; j < 4; ++j) draw[b * 4 + j] = n2 > 1e-12f ? inv * (dq[j] - q[j] * qdq) : inv * dq[j];
    }
  }
  part = block_sum(part, sh);
  if (threadIdx.x == 0 && loss_out != nullptr) loss_out[0] = weight * part / (float)B;
}

static inline int grid_for(long long n, int block, int max_blocks) {
  long long g = (n + block - 1) / block;
  if (g > max_blocks) g = max_blocks;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace urso

using namespace urso;

extern "C" {

int urso_stem_stage(const void* img, int32_t img_is_u8, int32_t subtract_mean, const float* mean3, void* e_out,
                    int32_t B, int32_t H, int32_t W, int32_t part, void* stream) {
  URSO_REQUIRE(img && e_out && (!subtract_mean || mean3), "null pointer");
  URSO_REQUIRE(H % 2 == 0 && W % 2 == 0, "stem input must have even H, W");
  const long long total = (long long)B * (H / 2 + 3) * (W / 2 + 3) * 2;
  const int grid = grid_for(total, 256, num_sms() * 16);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (img_is_u8)
    stem_stage_kernel<true><<<grid, 256, 0, s>>>(img, subtract_mean, mean3, static_cast<__nv_bfloat16*>(e_out), B, H, W, part);
  else
    stem_stage_kernel<false><<<grid, 256, 0, s>>>(img, subtract_mean, mean3, static_cast<__nv_bfloat16*>(e_out), B, H, W, part);
  URSO_CUDA_OK(cudaGetLastError());
  return 0;
}

int urso_maxpool_fwd(const void* x, void* y, void* argmax, int32_t B, int32_t H, int32_t W, int32_t C, void* stream) {
  URSO_REQUIRE(x && y, "null pointer");
  URSO_REQUIRE(H % 2 == 0 && W % 2 == 0 && C % 8 == 0, "maxpool needs even H, W and C %% 8 == 0");
  const long long total = (long long)B * (H / 2) * (W / 2) * (C / 8);
  maxpool_fwd_kernel<<<grid_for(total, 256, num_sms() * 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), static_cast<uint8_t*>(argmax), B, H, W, C);
  URSO_CUDA_OK(cudaGetLastError());
  return 0;
}

int urso_maxpool_fwd_f32(const float* x, float* y, int32_t B, int32_t H, int32_t W, int32_t C, void* stream) {
  URSO_REQUIRE(x && y, "null pointer");
  URSO_REQUIRE(H % 2 == 0 && W % 2 == 0 && C % 4 == 0, "maxpool needs even H, W and C %% 4 == 0");
  const long long total = (long long)B * (H / 2) * (W / 2) * (C / 4);
  maxpool_fwd_f32_kernel<<<grid_for(total, 256, num_sms() * 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, B, H, W, C);
  URSO_CUDA_OK(cudaGetLastError());
  return 0;
}

int urso_maxpool_bwd(const void* x, const void* argmax, const void* dy, void* dx, float* colsum, int32_t B, int32_t H,
                     int32_t W, int32_t C, void* stream) {
  URSO_REQUIRE(argmax && dy && dx, "null pointer");
  URSO_REQUIRE(H % 2 == 0 && W % 2 == 0 && C % 8 == 0 && C <= 512, "maxpool_bwd needs even H, W and C %% 8 == 0, C <= 512");
  const long long total = (long long)B * (H / 2) * (W / 2) * (C / 8);
  URSO_REQUIRE(256 % (C / 8) == 0, "C/8 must divide the block size");   // keeps a thread on one channel group
  maxpool_bwd_kernel<<<grid_for(total, 256, num_sms() * 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<const uint8_t*>(argmax), static_cast<const __nv_bfloat16*>(dy),
      static_cast<__nv_bfloat16*>(dx), colsum, B, H, W, C);
  URSO_CUDA_OK(cudaGetLastError());
  return 0;
}

int urso_dense_fwd(const float* x, const float* w, float* y, int32_t B, int32_t K, int32_t N, void* stream) {
  URSO_REQUIRE(x && w && y, "null pointer");
  return dense_fwd2(x, w, y, B, K, N, static_cast<cudaStream_t>(stream));      // dense.cu
}

int urso_dense_bias_act(float* y, const float* bias, int32_t B, int32_t N, int32_t act, void* stream) {
  URSO_REQUIRE(y && bias, "null pointer");
  dense_bias_act_kernel<<<grid_for((long long)B * N, 256, num_sms() * 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      y, bias, B, N, act);
  URSO_CUDA_OK(cudaGetLastError());
  return 0;
}

int urso_dense_bwd(const float* x, const float* w, const float* y, float* dy, float* dx, float* dw, float* db,
                   int32_t B, int32_t K, int32_t N, int32_t act, void* stream) {
  URSO_REQUIRE(x && w && dy, "null pointer");
  URSO_REQUIRE(act == 0 || y != nullptr, "relu backward needs the forward output");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  dense_mask_bias_grad_kernel<<<(N + 127) / 128, 128, 0, s>>>(y, dy, db, B, N, act);
  if (dw != nullptr) {
    if (int rc = dense_wgrad2(x, dy, dw, B, K, N, s)) return rc;
  }
  if (dx != nullptr) {
    if (int rc = dense_dgrad2(dy, w, dx, B, K, N, s)) return rc;
  }
  URSO_CUDA_OK(cudaGetLastError());
  return 0;
}

int urso_softmax_xent(const float* z, const float* y, float* dz, float* loss_out, int32_t B, int32_t N, float weight,
                      void* stream) {
  URSO_REQUIRE(z && y && loss_out, "null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  URSO_CUDA_OK(cudaMemsetAsync(loss_out, 0, sizeof(float), s));
  softmax_xent_kernel<<<B, 256, 0, s>>>(z, y, dz, loss_out, B, N, weight);
  URSO_CUDA_OK(cudaGetLastError());
  return 0;
}

int urso_rel_loss(const float* pred, const float* gt, float* dpred, float* loss_out, int32_t B, int32_t N, float weight,
                  void* stream) {
  URSO_REQUIRE(pred && gt && loss_out, "null pointer");
  rel_loss_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(pred, gt, dpred, loss_out, B * N, weight);
  URSO_CUDA_OK(cudaGetLastError());
  return 0;
}

int urso_quat_head(const float* raw, const float* gt, float* q_out, float* draw, float* loss_out, int32_t B,
                   float weight, void* stream) {
  URSO_REQUIRE(raw && q_out, "null pointer");
  quat_head_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(raw, gt, q_out, draw, loss_out, B, weight);
  URSO_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------ orientation soft labels
//  device versions of utils.encode_ori_fast (utils.py:319-346) and of the PMF decode of
// pose_estimator.py:406-409 (stable_softmax + the 4x4 moment matrix of se3lib.quat_weighted_avg).
namespace urso {

// enc[b, k] = exp(-2 (acos(min(1, |<q_b, H_k>|)) / pi)^2 / var) for non-redundant bins, normalised over k.
__global__ void __launch_bounds__(256) encode_ori_kernel(const float* __restrict__ quats, const float* __restrict__ hquat,
                                                         const uint8_t* __restrict__ redundant, float* __restrict__ enc,
                                                         int nbins, float var) {
  __shared__ float sh[32];
  const int b = blockIdx.x;
  const float q0 = quats[b * 4 + 0], q1 = quats[b * 4 + 1], q2 = quats[b * 4 + 2], q3 = quats[b * 4 + 3];
  float sum = 0.f;
  for (int k = threadIdx.x; k < nbins; k += blockDim.x) {
    float p = 0.f;
    if (!redundant[k]) {
      const float4 h = *reinterpret_cast<const float4*>(hquat + 4 * k);
      const float d = fminf(1.0f, fabsf(q0 * h.x + q1 * h.y + q2 * h.z + q3 * h.w));
      const float a = acosf(d) * 0.318309886183790672f;
      p = expf(-2.0f * a * a / var);
    }
    enc[(long long)b * nbins + k] = p;
    sum += p;
  }
  sum = block_sum(sum, sh);
  const float inv = 1.0f / sum;
  for (int k = threadIdx.x; k < nbins; k += blockDim.x) enc[(long long)b * nbins + k] *= inv;
}

// A_b = sum_k softmax(z_b)_k * H_k H_k^T   (symmetric 4x4, 10 unique entries written as a full 16-float matrix)
__global__ void __launch_bounds__(256) decode_moments_kernel(const float* __restrict__ logits,
                                                             const float* __restrict__ hquat, float* __restrict__ A,
                                                             int nbins) {
  __shared__ float sh[32];
  const int b = blockIdx.x;
  const float* z = logits + (long long)b * nbins;
  float m = -INFINITY;
  for (int k = threadIdx.x; k < nbins; k += blockDim.x) m = fmaxf(m, z[k]);
  m = block_max(m, sh);
  float acc[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) acc[i] = 0.f;
  float se = 0.f;
  for (int k = threadIdx.x; k < nbins; k += blockDim.x) {
    const float w = expf(z[k] - m);
    se += w;
    const float4 h = *reinterpret_cast<const float4*>(hquat + 4 * k);
    acc[0] += w * h.x * h.x; acc[1] += w * h.x * h.y; acc[2] += w * h.x * h.z; acc[3] += w * h.x * h.w;
    acc[4] += w * h.y * h.y; acc[5] += w * h.y * h.z; acc[6] += w * h.y * h.w;
    acc[7] += w * h.z * h.z; acc[8] += w * h.z * h.w; acc[9] += w * h.w * h.w;
  }
  se = block_sum(se, sh);
  const float inv = 1.0f / se;
  float r[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) r[i] = block_sum(acc[i], sh) * inv;
  if (threadIdx.x == 0) {
    float* o = A + b * 16;
    o[0] = r[0]; o[1] = r[1]; o[2] = r[2]; o[3] = r[3];
    o[4] = r[1]; o[5] = r[4]; o[6] = r[5]; o[7] = r[6];
    o[8] = r[2]; o[9] = r[5]; o[10] = r[7]; o[11] = r[8];
    o[12] = r[3]; o[13] = r[6]; o[14] = r[8]; o[15] = r[9];
  }
}

}  // namespace urso

extern "C" {

int urso_encode_ori(const float* quats, const float* hquat, const uint8_t* redundant, float* enc, int32_t B,
                    int32_t nbins, float var, void* stream) {
  URSO_REQUIRE(quats && hquat && redundant && enc, "null pointer");
  URSO_REQUIRE((reinterpret_cast<uintptr_t>(hquat) & 15) == 0, "hquat must be 16-byte aligned");
  urso::encode_ori_kernel<<<B, 256, 0, static_cast<cudaStream_t>(stream)>>>(quats, hquat, redundant, enc, nbins, var);
  URSO_CUDA_OK(cudaGetLastError());
  return 0;
}

int urso_decode_ori_moments(const float* logits, const float* hquat, float* A, int32_t B, int32_t nbins, void* stream) {
  URSO_REQUIRE(logits && hquat && A, "null pointer");
  urso::decode_moments_kernel<<<B, 256, 0, static_cast<cudaStream_t>(stream)>>>(logits, hquat, A, nbins);
  URSO_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
