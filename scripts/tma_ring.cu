// Microbenchmark (round 2, last session): what does a TMA -> mbarrier operand ring deliver per SM, as a function of the stage
// size and the number of stages, when every SM runs one?  It isolates the supply side of Engine F's stream mode: per K step a
// stage receives `na` activation boxes (128 rows x 128 B = 16 KB each, rows 2 KB apart: a 1024-channel NHWC tensor streamed
// from DRAM, every byte read once) and optionally one weight box (`brows` rows x 128 B of a small matrix that every CTA
// re-reads: L2 resident); a consumer warp waits for the stage, idles `delay` cycles (the MMAs' time) and releases it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/tma_ring scripts/tma_ring.cu && scripts/tma_ring
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../ursonet_b200/csrc/ptx.cuh"
using namespace urso;

struct RingParams {
  CUtensorMap a_map, b_map;      // a: [rows_a, 1024] bf16, box {64, 128};  b: [256, 1024] bf16, box {64, brows}
  int stages, na, brows, ksteps, tiles_per_cta, delay, a_wrap_tiles;
  int mma;           // 1: the consumer is the real thing -- 4 tcgen05.mma (M = 128, N = 256, K = 16) per K step on the stage's tiles and a
                     // tcgen05.commit that releases the stage when they have completed (na = 1, brows = 256, joint ring)
  int a_stages;      // > 0: SPLIT rings -- the activation boxes get their own ring of a_stages slots (own barriers, own producer warp)
};

__global__ void __launch_bounds__(128, 1) ring_kernel(const __grid_constant__ RingParams p, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full_bar[16], empty_bar[16], afull_bar[16], aempty_bar[16];
  const int warp = threadIdx.x >> 5;
  const int stage_bytes = p.na * 16384 + p.brows * 128;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
      mbar_init(&afull_bar[i], 1);
      mbar_init(&aempty_bar[i], 1);
    }
    fence_barrier_init();
    tma_prefetch_desc(&p.a_map);
    tma_prefetch_desc(&p.b_map);
  }
  __shared__ uint32_t tmem_slot;
  if (p.mma && warp == 3) tmem_alloc(&tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const long long t0 = clock64();
  if (p.mma && warp == 1) {
    // real consumer: descriptor = constant high word | (smem address >> 4); K advance of 16 elements = +2
    constexpr uint64_t kHi = (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 16) | (1ull << 46) | (2ull << 61);
    constexpr uint32_t idesc = umma_idesc_bf16(128, 256, 0, 0);
    const uint32_t d_tmem = tmem_slot;
    const uint32_t base = smem_u32(smem);
    int stage = 0;
    uint32_t phase = 0;
    const int n = p.tiles_per_cta * p.ksteps;
    for (int i = 0; i < n; ++i) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t ad = kHi | ((base + stage * stage_bytes) >> 4);
        const uint64_t bd = kHi | ((base + stage * stage_bytes + 16384) >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (i | k) != 0);
        umma_commit(&empty_bar[stage]);
      }
      __syncwarp();
      if (++stage == p.stages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (p.a_stages > 0) {
    // split rings: [a_stages x na x 16 KB][stages x brows x 128 B]
    uint8_t* a_ring = smem;
    uint8_t* b_ring = smem + p.a_stages * p.na * 16384;
    const int n = p.tiles_per_cta * p.ksteps;
    if (warp == 0) {            // activation producer
      int st = 0;
      uint32_t ph = 0;
      for (int t = 0; t < p.tiles_per_cta; ++t) {
        const int row0 = ((int)blockIdx.x * p.tiles_per_cta + t) * 128 * p.na;
        for (int ks = 0; ks < p.ksteps; ++ks) {
          mbar_wait(&aempty_bar[st], ph ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&afull_bar[st], p.na * 16384);
            for (int a = 0; a < p.na; ++a)
              tma_load_2d(a_ring + (st * p.na + a) * 16384, &p.a_map, &afull_bar[st], ks * 64, row0 + a * 128);
          }
          __syncwarp();
          if (++st == p.a_stages) { st = 0; ph ^= 1; }
        }
      }
    } else if (warp == 2) {     // weight producer
      int st = 0;
      uint32_t ph = 0;
      for (int i = 0; i < n; ++i) {
        mbar_wait(&empty_bar[st], ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&full_bar[st], p.brows * 128);
          tma_load_2d(b_ring + st * p.brows * 128, &p.b_map, &full_bar[st], (i % p.ksteps) * 64, 0);
        }
        __syncwarp();
        if (++st == p.stages) { st = 0; ph ^= 1; }
      }
    } else if (warp == 1) {     // consumer
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      for (int i = 0; i < n; ++i) {
        mbar_wait(&afull_bar[sa], pa);
        mbar_wait(&full_bar[sb], pb);
        if (p.delay > 0) {
          const long long t = clock64();
          while (clock64() - t < p.delay) {}
        }
        if (threadIdx.x == 32) {
          mbar_arrive(&aempty_bar[sa]);
          mbar_arrive(&empty_bar[sb]);
        }
        __syncwarp();
        if (++sa == p.a_stages) { sa = 0; pa ^= 1; }
        if (++sb == p.stages) { sb = 0; pb ^= 1; }
      }
    }
  } else if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int t = 0; t < p.tiles_per_cta; ++t) {
      // tile index: DRAM streaming = every CTA its own rows; a_wrap_tiles > 0: the CTA re-reads the same few tiles (L2 hits)
      const int tt = p.a_wrap_tiles > 0 ? (t % p.a_wrap_tiles) : t;
      const int row0 = ((int)blockIdx.x * (p.a_wrap_tiles > 0 ? p.a_wrap_tiles : p.tiles_per_cta) + tt) * 128 * p.na;
      for (int ks = 0; ks < p.ksteps; ++ks) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* dst = smem + stage * stage_bytes;
          mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
          for (int a = 0; a < p.na; ++a) tma_load_2d(dst + a * 16384, &p.a_map, &full_bar[stage], ks * 64, row0 + a * 128);
          if (p.brows) tma_load_2d(dst + p.na * 16384, &p.b_map, &full_bar[stage], ks * 64, 0);
        }
        __syncwarp();
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1 && !p.mma) {
    int stage = 0;
    uint32_t phase = 0;
    const int n = p.tiles_per_cta * p.ksteps;
    for (int i = 0; i < n; ++i) {
      mbar_wait(&full_bar[stage], phase);
      if (p.delay > 0) {
        const long long t = clock64();
        while (clock64() - t < p.delay) {}
      }
      if (threadIdx.x == 32) mbar_arrive(&empty_bar[stage]);
      __syncwarp();
      if (++stage == p.stages) {
        stage = 0;
        phase ^= 1;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
  if (p.mma) {
    // drain: every MMA has completed once the last stages were released; one more commit + wait keeps it simple
    __shared__ uint64_t done_bar;
    if (threadIdx.x == 32) {
      mbar_init(&done_bar, 1);
      fence_barrier_init();
      umma_commit(&done_bar);
      mbar_wait(&done_bar, 0);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 3) {
      tc_fence_after();
      tmem_dealloc(tmem_slot, 256);
    }
  }
}

typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static void make_map(encode_fn enc, CUtensorMap* m, void* base, long long rows, int box_rows) {
  cuuint64_t dims[2] = {1024, (cuuint64_t)rows};
  cuuint64_t strides[1] = {2048};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("encode failed %d\n", (int)r);
    exit(1);
  }
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  encode_fn enc = reinterpret_cast<encode_fn>(fn);
  const int tiles = 12, ksteps = 16;
  const long long rows_a = (long long)sms * tiles * 128 * 2;      // up to na = 2
  void *a = nullptr, *b = nullptr;
  long long* cyc = nullptr;
  cudaMalloc(&a, rows_a * 2048);
  cudaMalloc(&b, 256 * 2048);
  cudaMalloc(&cyc, sms * sizeof(long long));
  cudaMemset(a, 0, rows_a * 2048);
  cudaMemset(b, 0, 256 * 2048);
  cudaFuncSetAttribute(ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
  printf("%d SMs, %d MHz nominal; per K step: na x 16 KB activation boxes (DRAM stream unless 'L2') + brows x 128 B weight box (L2)\n", sms,
         khz / 1000);
  printf("%-26s %6s %7s %9s %10s %10s %9s\n", "stage", "stages", "delay", "ring KB", "B/clk/SM", "GB/s/SM", "TB/s all");
  struct Cfg { int na, brows, stages, delay, wrap, a_stages, mma; };
  const Cfg cfgs[] = {
      {1, 256, 2, 0, 0}, {1, 256, 3, 0, 0}, {1, 256, 4, 0, 0},                       // Engine F stream stage: 48 KB
      {1, 256, 4, 512, 0},                                                             // + the MMAs' time (512 cycles per K step)
      {1, 128, 4, 0, 0}, {1, 128, 6, 0, 0}, {1, 128, 6, 512, 0},                     // a CTA pair's stage: 32 KB
      {2, 256, 3, 0, 0}, {2, 256, 3, 1024, 0},                                         // dual-M stage: 64 KB
      {1, 0, 4, 0, 0}, {1, 0, 8, 0, 0}, {1, 0, 12, 0, 0},                            // activation stream only: 16 KB boxes
      {2, 0, 6, 0, 0},                                                                 // 32 KB of activations per stage
      {1, 256, 4, 0, 2}, {1, 256, 2, 0, 2}, {1, 128, 6, 0, 2}, {1, 0, 12, 0, 2},     // everything L2 resident
      // consumer time of the real kernel (~316 cycles of issue overhead + 512 of MMAs): joint ring vs SPLIT rings (the DRAM
      // stream of activations gets more slots than the L2-resident weight tiles)
      {1, 256, 4, 830, 0, 0}, {1, 256, 4, 830, 0, 6}, {1, 256, 3, 830, 0, 8}, {1, 256, 4, 512, 0, 6}, {1, 256, 4, 0, 0, 6},
      {1, 256, 4, 1100, 0, 0}, {1, 256, 4, 1100, 0, 6},
      // the real consumer: tcgen05.mma + commit (no epilogue, one accumulator)
      {1, 256, 4, 0, 0, 0, 1}, {1, 256, 3, 0, 0, 0, 1}, {1, 256, 2, 0, 0, 0, 1}, {1, 256, 4, 0, 2, 0, 1},
  };
  for (const Cfg& c : cfgs) {
    RingParams p;
    make_map(enc, &p.a_map, a, rows_a, 128);
    make_map(enc, &p.b_map, b, 256, c.brows ? c.brows : 128);
    p.stages = c.stages; p.na = c.na; p.brows = c.brows; p.ksteps = ksteps; p.tiles_per_cta = tiles; p.delay = c.delay;
    p.a_wrap_tiles = c.wrap;
    p.a_stages = c.a_stages;
    p.mma = c.mma;
    const int stage_bytes = c.na * 16384 + c.brows * 128;
    const int smem = c.a_stages > 0 ? c.a_stages * c.na * 16384 + c.stages * c.brows * 128 : c.stages * stage_bytes;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      ring_kernel<<<sms, 128, smem>>>(p, cyc);
      cudaEventRecord(e1);
      cudaError_t err = cudaEventSynchronize(e1);
      if (err != cudaSuccess) {
        printf("kernel failed: %s\n", cudaGetErrorString(err));
        return 1;
      }
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;
    }
    long long hc[256];
    cudaMemcpy(hc, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < sms; ++i) avg += (double)hc[i] / sms;
    const double bytes_sm = (double)tiles * ksteps * stage_bytes;
    char name[64];
    snprintf(name, sizeof(name), "%dx16KB A + %2d KB B%s%s", c.na, c.brows * 128 / 1024, c.wrap ? " (L2)" : "", c.a_stages ? " SPLIT" : (c.mma ? " MMA" : ""));
    if (c.a_stages) printf("  (next line: %d activation slots + %d weight slots)\n", c.a_stages, c.stages);
    printf("%-26s %6d %7d %9d %10.1f %10.1f %9.2f\n", name, c.stages, c.delay, smem / 1024, bytes_sm / avg,
           bytes_sm / (best * 1e-3) / 1e9, bytes_sm * sms / (best * 1e-3) / 1e12);
  }
  return 0;
}
