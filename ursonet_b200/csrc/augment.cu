// sim2real augmentation on the device (reference: net.py:390-406).
//
// Per image: luma written back into the three uint8 channels (truncation, net.py:391-394), then -- for the images the host
// drew with p = 0.5 -- the five imgaug operations of the reference's iaa.Sequential(random_order=True) in the drawn order:
//   0 AdditiveGaussianNoise(scale = 0.01*255)   1 GaussianBlur(sigma in [0, 1.5]; 5x5 kernel, reflect-101)
//   2 Add([-20, 20])                            3 Multiply([0.5, 2.0])
//   4 CoarseDropout(p in {0, 0.03}, size_percent in [0.02, 0.1]; nearest-upsampled low-resolution mask)
// every operation rounds (half to even) and clips to uint8 like imgaug does between augmenters.  All per-image random
// parameters are drawn on the host (ursonet_b200/augment.py); per-pixel randomness is a counter-based integer hash of
// (seed, pixel / cell index), so the numpy restatement (oracle/sim2real_oracle.py) reproduces the kernel BIT-EXACTLY.
// The image is grey after the luma step, so one channel is processed and written three times.
//
// One CTA = one strip of 16 rows of one image (see the kernel).  HBM-bound: reads 3 B + writes 3 B per pixel.
#include "common.cuh"

namespace urso {

constexpr int kTileW = 64, kTileH = 16, kHalo = 2;
constexpr int kSW = kTileW + 2 * kHalo, kSH = kTileH + 2 * kHalo;

__device__ __forceinline__ uint32_t hash_u32(uint32_t seed, uint32_t idx) {   // lowbias32-style integer hash
  uint32_t x = idx * 0x9E3779B1u + seed;
  x ^= x >> 16; x *= 0x7FEB352Du;
  x ^= x >> 15; x *= 0x846CA68Bu;
  x ^= x >> 16;
  return x;
}
__device__ __forceinline__ float round_clip_u8(float v) { return fminf(fmaxf(rintf(v), 0.f), 255.f); }
__device__ __forceinline__ int reflect101(int i, int lo, int hi) {   // cv2.BORDER_REFLECT_101 on [lo, hi)
  const int n = hi - lo;
  if (n == 1) return lo;
  i -= lo;
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  if (i < 0) i = 0;
  return lo + i;
}

// trunc(0.2126 R + 0.7152 G + 0.0722 B) as numpy evaluates it in float64 (net.py:391).  The exact value is n / 10000 with
// n = 2126 R + 7152 G + 722 B; float64 rounding (a few 1e-14) can only move the truncation when n is a multiple of 10000,
// so everything else is integer arithmetic and only those rare triples take the (slow on this part) FP64 path.
__device__ __forceinline__ float luma_u8(uint32_t r, uint32_t g, uint32_t b) {
  const uint32_t n = 2126u * r + 7152u * g + 722u * b;
  const uint32_t q = n / 10000u;
  if (n - q * 10000u != 0u) return (float)q;
  const double v = __dadd_rn(__dadd_rn(__dmul_rn(0.2126, (double)r), __dmul_rn(0.7152, (double)g)),
                             __dmul_rn(0.0722, (double)b));
  return (float)(int)v;
}

// pointwise operations (everything except the blur) on one grey value at absolute pixel (y, x)
__device__ __forceinline__ float apply_pointwise(int op, float v, const urso_aug_params& a, int y, int x, int W) {
  if (op == 0) {          // additive noise: Irwin-Hall(4) integer approximation of N(0, sigma), exact in integers
    const uint32_t h = hash_u32(a.noise_seed, (uint32_t)(y * W + x));
    const int z = (int)(h & 255u) + (int)((h >> 8) & 255u) + (int)((h >> 16) & 255u) + (int)(h >> 24) - 510;
    const int n = (z * a.noise_q + (z >= 0 ? 32768 : -32768)) / 65536;    // round(z * sigma / 147.8), symmetric
    return fminf(fmaxf(v + (float)n, 0.f), 255.f);
  } else if (op == 2) {   // Add
    return fminf(fmaxf(v + (float)a.add, 0.f), 255.f);
  } else if (op == 3) {   // Multiply
    return round_clip_u8(__fmul_rn(v, a.mul));
  } else if (op == 4) {   // coarse dropout: one Bernoulli(p) draw per low-resolution cell, nearest upsampling
    const int wy0 = a.win[0], wx0 = a.win[1], wh = a.win[2] - a.win[0], ww = a.win[3] - a.win[1];
    // 32-bit arithmetic is exact here: coordinates < 2^16, grid sizes < 2^15 (a 64-bit division costs ~100 instructions)
    const int cy = (int)(((uint32_t)(y - wy0) * (uint32_t)a.drop_h) / (uint32_t)wh);
    const int cx = (int)(((uint32_t)(x - wx0) * (uint32_t)a.drop_w) / (uint32_t)ww);
    const uint32_t h = hash_u32(a.drop_seed, (uint32_t)(cy * a.drop_w + cx));
    return h < a.drop_thresh ? 0.f : v;
  }
  return v;
}

// Grid = (strips of 16 rows, images).  The image's parameter record is staged in shared memory once per CTA.
//   * images without blur (not selected, or sigma below imgaug's cutoff): every thread owns one 4-pixel group (12 bytes =
//     three aligned 32-bit words) per row and walks the 16 rows of the strip, four rows of loads in flight -- no shared
//     tiles, no integer divisions;
//   * blurred images: the strip is processed as 64-pixel-wide tiles with a 2-pixel halo staged in shared memory, so the
//     5x5 separable Gaussian sees neighbours that already went through the augmenters drawn before it.
__global__ void __launch_bounds__(256) sim2real_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                       const urso_aug_params* __restrict__ params, int H, int W) {
  __shared__ float s0[kSH][kSW + 1];
  __shared__ float s1[kSH][kSW + 1];
  __shared__ urso_aug_params s_a;
  const int b = blockIdx.y;
  if (threadIdx.x < sizeof(urso_aug_params) / 4)
    reinterpret_cast<uint32_t*>(&s_a)[threadIdx.x] = reinterpret_cast<const uint32_t*>(params + b)[threadIdx.x];
  __syncthreads();
  const urso_aug_params& a = s_a;
  // blockIdx.x = strip * tiles_x + tile column: a blurred image uses every CTA (one 64-wide tile each, independent CTAs
  // hide the staging latency best); an image without blur is done by the tile-column-0 CTA of each strip, the rest exit
  const int tiles_x = (W + kTileW - 1) / kTileW;
  const int strip = blockIdx.x / tiles_x, tcol = blockIdx.x - strip * tiles_x;
  const int y0 = strip * kTileH;
  const uint8_t* img = src + (size_t)b * H * W * 3;
  uint8_t* out = dst + (size_t)b * H * W * 3;
  const int wy0 = a.win[0], wx0 = a.win[1], wy1 = a.win[2], wx1 = a.win[3];
  const int apply = a.apply;
  // position of the blur in the drawn order (5 = none: sigma below imgaug's 1e-3 cutoff, or image not augmented)
  int blur_at = 5;
  if (apply) {
#pragma unroll
    for (int k = 0; k < 5; ++k)
      if (a.order[k] == 1 && a.blur_sigma >= 1e-3f) blur_at = k;
  }
  auto luma_at = [&](int y, int x) -> float {
    const uint8_t* p = img + ((size_t)y * W + x) * 3;
    return luma_u8(p[0], p[1], p[2]);
  };

  if (blur_at == 5) {
    // ------------------------------------------------------------------ no blur: pointwise only
    if (tcol != 0) return;
    if ((W & 3) == 0 && ((reinterpret_cast<uintptr_t>(img) | reinterpret_cast<uintptr_t>(out)) & 3) == 0) {
      const int groups = W >> 2;
      // per-image values hoisted out of the pixel loops (all block-uniform)
      int ops[5];
#pragma unroll
      for (int k = 0; k < 5; ++k) ops[k] = a.order[k];
      const int noise_q = a.noise_q, add_i = a.add, drop_h = a.drop_h, drop_w = a.drop_w;
      const uint32_t noise_seed = a.noise_seed, drop_seed = a.drop_seed, drop_thresh = a.drop_thresh;
      const float mul = a.mul;
      const uint32_t wh = (uint32_t)(wy1 - wy0), ww = (uint32_t)(wx1 - wx0);
      for (int g = threadIdx.x; g < groups; g += blockDim.x) {
        const int x = g * 4;
        // column-only quantities: which of the 4 pixels lie inside the window, their dropout cell column
        bool in_col[4];
        uint32_t cx[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          in_col[j] = x + j >= wx0 && x + j < wx1;
          cx[j] = (apply && in_col[j] && drop_thresh != 0u) ? ((uint32_t)(x + j - wx0) * (uint32_t)drop_w) / ww : 0u;
        }
#pragma unroll 4
        for (int r = 0; r < kTileH; ++r) {
          const int y = y0 + r;
          if (y >= H) continue;
          const uint32_t* p = reinterpret_cast<const uint32_t*>(img + ((size_t)y * W + x) * 3);
          const uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
          float v[4];
          v[0] = luma_u8(w0 & 255u, (w0 >> 8) & 255u, (w0 >> 16) & 255u);
          v[1] = luma_u8(w0 >> 24, w1 & 255u, (w1 >> 8) & 255u);
          v[2] = luma_u8((w1 >> 16) & 255u, w1 >> 24, w2 & 255u);
          v[3] = luma_u8((w2 >> 8) & 255u, (w2 >> 16) & 255u, w2 >> 24);
          if (apply && y >= wy0 && y < wy1) {
            // one (block-uniform) dispatch per operation for the four pixels, same arithmetic as apply_pointwise
            const uint32_t cyw = drop_thresh != 0u ? (((uint32_t)(y - wy0) * (uint32_t)drop_h) / wh) * (uint32_t)drop_w : 0u;
#pragma unroll
            for (int k = 0; k < 5; ++k) {
              const int op = ops[k];
              if (op == 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const uint32_t h = hash_u32(noise_seed, (uint32_t)(y * W + x + j));
                  const int z = (int)(h & 255u) + (int)((h >> 8) & 255u) + (int)((h >> 16) & 255u) + (int)(h >> 24) - 510;
                  const int n = (z * noise_q + (z >= 0 ? 32768 : -32768)) / 65536;
                  if (in_col[j]) v[j] = fminf(fmaxf(v[j] + (float)n, 0.f), 255.f);
                }
              } else if (op == 2) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  if (in_col[j]) v[j] = fminf(fmaxf(v[j] + (float)add_i, 0.f), 255.f);
              } else if (op == 3) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  if (in_col[j]) v[j] = round_clip_u8(__fmul_rn(v[j], mul));
              } else if (op == 4 && drop_thresh != 0u) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  if (in_col[j] && hash_u32(drop_seed, cyw + cx[j]) < drop_thresh) v[j] = 0.f;
              }
            }
          }
          const uint32_t g0 = (uint32_t)v[0], g1 = (uint32_t)v[1], g2 = (uint32_t)v[2], g3 = (uint32_t)v[3];
          uint32_t* o = reinterpret_cast<uint32_t*>(out + ((size_t)y * W + x) * 3);
          o[0] = g0 * 0x010101u | (g1 << 24);
          o[1] = g1 * 0x0101u | (g2 * 0x0101u << 16);
          o[2] = g2 | (g3 * 0x010101u << 8);
        }
      }
    } else {   // odd widths / unaligned bases: one pixel at a time
      for (int i = threadIdx.x; i < kTileH * W; i += blockDim.x) {
        const int y = y0 + i / W, x = i % W;
        if (y >= H) break;
        float v = luma_at(y, x);
        if (apply && y >= wy0 && y < wy1 && x >= wx0 && x < wx1) {
          for (int k = 0; k < 5; ++k) v = apply_pointwise(a.order[k], v, a, y, x, W);
        }
        const uint8_t g = (uint8_t)v;
        uint8_t* p = out + ((size_t)y * W + x) * 3;
        p[0] = g; p[1] = g; p[2] = g;
      }
    }
    return;
  }

  // -------------------------------------------------------------------- blurred image: this CTA's 64-wide tile
  {
    const int x0 = tcol * kTileW;
    // stage 1: the blur's input -- luma + the operations before the blur -- for the tile and its halo.  Positions outside
    // the image window take the value of the pixel they reflect to (BORDER_REFLECT_101 on the window: the reference pads
    // AFTER augmenting, so the blur never sees padding).
    for (int i = threadIdx.x; i < kSH * kSW; i += blockDim.x) {
      const int sy = i / kSW, sx = i % kSW;
      const int y = reflect101(y0 + sy - kHalo, wy0, wy1), x = reflect101(x0 + sx - kHalo, wx0, wx1);
      float v = luma_at(y, x);
      for (int k = 0; k < blur_at; ++k) v = apply_pointwise(a.order[k], v, a, y, x, W);
      s0[sy][sx] = v;
    }
    __syncthreads();
    // 5-tap separable Gaussian in float32 with un-fused multiply-adds in a fixed order (matches numpy float32)
    for (int i = threadIdx.x; i < kSH * kTileW; i += blockDim.x) {
      const int sy = i / kTileW, sx = i % kTileW + kHalo;
      float acc = __fmul_rn(s0[sy][sx - 2], a.blur_w[0]);
      acc = __fadd_rn(acc, __fmul_rn(s0[sy][sx - 1], a.blur_w[1]));
      acc = __fadd_rn(acc, __fmul_rn(s0[sy][sx], a.blur_w[2]));
      acc = __fadd_rn(acc, __fmul_rn(s0[sy][sx + 1], a.blur_w[3]));
      acc = __fadd_rn(acc, __fmul_rn(s0[sy][sx + 2], a.blur_w[4]));
      s1[sy][sx] = acc;
    }
    __syncthreads();
    // stage 2: vertical pass + the operations after the blur; write the three channels
    for (int i = threadIdx.x; i < kTileH * kTileW; i += blockDim.x) {
      const int ty = i / kTileW, tx = i % kTileW;
      const int y = y0 + ty, x = x0 + tx;
      if (y >= H || x >= W) continue;
      const int sy = ty + kHalo, sx = tx + kHalo;
      float v;
      if (y >= wy0 && y < wy1 && x >= wx0 && x < wx1) {
        float acc = __fmul_rn(s1[sy - 2][sx], a.blur_w[0]);
        acc = __fadd_rn(acc, __fmul_rn(s1[sy - 1][sx], a.blur_w[1]));
        acc = __fadd_rn(acc, __fmul_rn(s1[sy][sx], a.blur_w[2]));
        acc = __fadd_rn(acc, __fmul_rn(s1[sy + 1][sx], a.blur_w[3]));
        acc = __fadd_rn(acc, __fmul_rn(s1[sy + 2][sx], a.blur_w[4]));
        v = round_clip_u8(acc);
        for (int k = blur_at + 1; k < 5; ++k) v = apply_pointwise(a.order[k], v, a, y, x, W);
      } else {
        v = luma_at(y, x);      // padding: luma only
      }
      const uint8_t g = (uint8_t)v;
      uint8_t* p = out + ((size_t)y * W + x) * 3;
      p[0] = g; p[1] = g; p[2] = g;
    }
  }
}

}  // namespace urso

extern "C" {

int urso_sizeof_aug_params(void) { return (int)sizeof(urso_aug_params); }

int urso_sim2real_aug(const uint8_t* src, uint8_t* dst, const urso_aug_params* params_dev, int32_t B, int32_t H,
                      int32_t W, void* stream) {
  using namespace urso;
  URSO_REQUIRE(src && dst && params_dev, "null pointer");
  URSO_REQUIRE(src != dst, "sim2real_aug is out of place (the blur reads neighbours)");
  URSO_REQUIRE(B >= 1 && B <= 65535 && H >= 1 && W >= 1, "bad shape");
  dim3 grid(((H + kTileH - 1) / kTileH) * ((W + kTileW - 1) / kTileW), B);
  sim2real_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, params_dev, H, W);
  URSO_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
