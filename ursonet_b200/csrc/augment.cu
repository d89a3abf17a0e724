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
// One CTA = one 64 x 16 pixel tile of one image; the tile plus a 2-pixel halo is staged in shared memory so that the blur
// (wherever it falls in the drawn order) sees neighbours that already went through the operations before it.
// HBM-bound: reads 3 B + writes 3 B per pixel.
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
    const int cy = (int)(((long long)(y - wy0) * a.drop_h) / wh), cx = (int)(((long long)(x - wx0) * a.drop_w) / ww);
    const uint32_t h = hash_u32(a.drop_seed, (uint32_t)(cy * a.drop_w + cx));
    return h < a.drop_thresh ? 0.f : v;
  }
  return v;
}

__global__ void __launch_bounds__(256) sim2real_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                       const urso_aug_params* __restrict__ params, int H, int W,
                                                       int n_tiles) {
  __shared__ float s0[kSH][kSW + 1];
  __shared__ float s1[kSH][kSW + 1];
  // persistent CTAs (grid = a multiple of the SM count) walk the (image, tile row, tile column) space
  const int tiles_x = (W + kTileW - 1) / kTileW, tiles_y = (H + kTileH - 1) / kTileH;
  const int tiles_per_img = tiles_x * tiles_y;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
  const int b = tile / tiles_per_img;
  const int tr = tile - b * tiles_per_img;
  const urso_aug_params a = params[b];   // (staging the record in shared memory measured slower: 76 registers)
  const int x0 = (tr % tiles_x) * kTileW, y0 = (tr / tiles_x) * kTileH;
  const uint8_t* img = src + (size_t)b * H * W * 3;
  uint8_t* out = dst + (size_t)b * H * W * 3;
  const int wy0 = a.win[0], wx0 = a.win[1], wy1 = a.win[2], wx1 = a.win[3];
  // position of the blur in the drawn order (5 = none: sigma below imgaug's 1e-3 cutoff, or image not augmented)
  int blur_at = 5;
  if (a.apply) {
#pragma unroll
    for (int k = 0; k < 5; ++k)
      if (a.order[k] == 1 && a.blur_sigma >= 1e-3f) blur_at = k;
  }
  const bool need_halo = blur_at < 5;
  // ---- stage 1 (only when blurring): the blur's input -- luma + the operations before the blur -- for the tile and its
  // halo.  Positions outside the image window take the value of the pixel they reflect to (BORDER_REFLECT_101 on the
  // window: the reference pads AFTER augmenting, so the blur never sees padding).
  auto luma_at = [&](int y, int x) -> float {
    const uint8_t* p = img + ((size_t)y * W + x) * 3;
    return luma_u8(p[0], p[1], p[2]);
  };
  // ---- fast path (no blur: luma-only images, or the drawn sigma is below the cutoff): 4 pixels = 12 bytes = three
  // aligned 32-bit words per thread, no shared memory
  if (!need_halo && (W & 3) == 0 && ((reinterpret_cast<uintptr_t>(img) | reinterpret_cast<uintptr_t>(out)) & 3) == 0) {
    const int ty = threadIdx.x >> 4, tx = (threadIdx.x & 15) * 4;
    const int y = y0 + ty, x = x0 + tx;
    if (y < H && x < W) {
      const uint32_t* p = reinterpret_cast<const uint32_t*>(img + ((size_t)y * W + x) * 3);
      const uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
      const uint32_t px[4][3] = {{w0 & 255u, (w0 >> 8) & 255u, (w0 >> 16) & 255u},
                                 {w0 >> 24, w1 & 255u, (w1 >> 8) & 255u},
                                 {(w1 >> 16) & 255u, w1 >> 24, w2 & 255u},
                                 {(w2 >> 8) & 255u, (w2 >> 16) & 255u, w2 >> 24}};
      uint32_t g[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v = luma_u8(px[j][0], px[j][1], px[j][2]);
        if (a.apply && y >= wy0 && y < wy1 && x + j >= wx0 && x + j < wx1) {
          for (int k = 0; k < 5; ++k) v = apply_pointwise(a.order[k], v, a, y, x + j, W);
        }
        g[j] = (uint32_t)v;
      }
      uint32_t* o = reinterpret_cast<uint32_t*>(out + ((size_t)y * W + x) * 3);
      o[0] = g[0] * 0x010101u | (g[1] << 24);
      o[1] = g[1] * 0x0101u | (g[2] * 0x0101u << 16);
      o[2] = g[2] | (g[3] * 0x010101u << 8);
    }
    continue;
  }
  if (need_halo) {
    for (int i = threadIdx.x; i < kSH * kSW; i += blockDim.x) {
      const int sy = i / kSW, sx = i % kSW;
      const int y = reflect101(y0 + sy - kHalo, wy0, wy1), x = reflect101(x0 + sx - kHalo, wx0, wx1);
      float v = luma_at(y, x);
      for (int k = 0; k < blur_at; ++k) v = apply_pointwise(a.order[k], v, a, y, x, W);
      s0[sy][sx] = v;
    }
  }
  __syncthreads();
  if (need_halo) {
    // ---- 5-tap separable Gaussian in float32 with un-fused multiply-adds in a fixed order (matches numpy float32)
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
  }
  // ---- stage 2: vertical pass + the operations after the blur (or all of them when there is no blur); write 3 channels
  for (int i = threadIdx.x; i < kTileH * kTileW; i += blockDim.x) {
    const int ty = i / kTileW, tx = i % kTileW;
    const int y = y0 + ty, x = x0 + tx;
    if (y >= H || x >= W) continue;
    const int sy = ty + kHalo, sx = tx + kHalo;
    const bool inside = y >= wy0 && y < wy1 && x >= wx0 && x < wx1;
    float v;
    if (a.apply && inside) {
      int k0 = 0;
      if (need_halo) {
        float acc = __fmul_rn(s1[sy - 2][sx], a.blur_w[0]);
        acc = __fadd_rn(acc, __fmul_rn(s1[sy - 1][sx], a.blur_w[1]));
        acc = __fadd_rn(acc, __fmul_rn(s1[sy][sx], a.blur_w[2]));
        acc = __fadd_rn(acc, __fmul_rn(s1[sy + 1][sx], a.blur_w[3]));
        acc = __fadd_rn(acc, __fmul_rn(s1[sy + 2][sx], a.blur_w[4]));
        v = round_clip_u8(acc);
        k0 = blur_at + 1;
      } else {
        v = luma_at(y, x);
      }
      for (int k = k0; k < 5; ++k) v = apply_pointwise(a.order[k], v, a, y, x, W);
    } else {
      v = luma_at(y, x);      // padding, or an image the host did not select: luma only
    }
    const uint8_t g = (uint8_t)v;
    uint8_t* p = out + ((size_t)y * W + x) * 3;
    p[0] = g; p[1] = g; p[2] = g;
  }
  __syncthreads();   // the shared tiles are reused by the next iteration
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
  const long long n_tiles = (long long)((W + kTileW - 1) / kTileW) * ((H + kTileH - 1) / kTileH) * B;
  URSO_REQUIRE(n_tiles < 0x7fffffffLL, "too many tiles");
  int sms = num_sms();
  if (sms <= 0) sms = 148;
  const int grid = (int)(n_tiles < (long long)sms * 8 ? n_tiles : (long long)sms * 8);
  sim2real_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, params_dev, H, W, (int)n_tiles);
  URSO_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
