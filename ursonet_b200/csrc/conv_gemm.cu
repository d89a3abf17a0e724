// Engine F: persistent, warp-specialised implicit-GEMM convolution on tcgen05 tensor cores.
//
//   D[pixel, n] = sum over K-segments (filter taps / concatenated inputs) of  A_seg[pixel + shift, c] * B[n, k]
//
//  * A tiles: 128 output pixels (a TH x TW patch of one image) x 64 channels, fetched by ONE 4-D TMA box per
//    K-step straight from the NHWC activation tensor.  The tap shift is added to the box coordinates and TMA's
//    out-of-bounds zero fill implements the convolution padding -- no im2col buffer ever exists.
//  * B tiles: BLOCK_N x 64 slices of the staged (BN-folded, K-major) weight matrix, 2-D TMA.
//  * both land in 128B-swizzled shared memory and are consumed by tcgen05.mma (M=128, N=BLOCK_N, K=16) issued by
//    one thread; fp32 accumulators live in TMEM, double buffered so the epilogue of tile i overlaps the MMAs of
//    tile i+1.
//  * epilogue (4 warps, one TMEM lane quadrant = 32 pixel rows each), per 64-channel chunk:
//      tcgen05.ld -> +shift -> +addend -> ReLU -> mask -> bf16
//    ALL global traffic of the epilogue is TMA as well: every warp prefetches its own 32-row slab of the addend /
//    mask tensors chunks ahead into a private smem ring (so HBM latency never sits on the critical path) and writes
//    its output slab with a TMA store from a double-buffered smem slab (full 128-byte lines, automatic clipping of
//    partial tiles, strided views for stride-2 dgrad).  No cross-warp synchronisation in the epilogue.
//    A legacy register epilogue (direct vector stores) remains for fp32 outputs / BLOCK_N = 32 (bottleneck conv).
//  * optional fused per-channel sums of the stored output (d beta): read back from the bf16 output slab
//    (conflict free), accumulated per CTA in shared memory, flushed with one atomic per channel per CTA.
//
// Roles (384 threads): warps 0 and 3 = TMA producers (even / odd K steps), warp 1 = MMA issuer, warp 2 = TMEM
// allocator, warps 4-11 = epilogue
// (4 TMEM lane quadrants x 2 column halves).
#include "common.cuh"
#include "ptx.cuh"

// Bottleneck-isolation knobs (URSO_DBG_NO_TMA / URSO_DBG_NO_MMA, scripts/bench_isolate.py) are compiled in only with
// `make DEBUG_KNOBS=1`: they put a branch into the producer and MMA-issue loops.
#ifndef URSO_DEBUG_KNOBS
#define URSO_DEBUG_KNOBS 0
#endif

namespace urso {

struct SegDev {
  int16_t map_id, dh, dw, c_chunks;
};

struct PixDev {
  void* ptr;
  long long sn, sh, sw;
};

// Division of a 31-bit unsigned value by a launch-time constant (Granlund-Montgomery, round-up variant): 3 instructions
// instead of the ~30 of a runtime integer division.  The tile decode runs once per 64-channel chunk in every epilogue
// warp, so this is on the critical path of the memory-bound layers.
struct FastDiv {
  uint32_t mul, shr, d;
};
__host__ inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  uint32_t l = 0;
  while ((1ull << l) < d) ++l;
  f.mul = (uint32_t)((((1ull << l) - d) << 32) / d + 1);
  f.shr = l;
  f.d = d;
  return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t x, const FastDiv& f) { return (__umulhi(x, f.mul) + x) >> f.shr; }

struct ConvGemmParams {
  CUtensorMap a_maps[URSO_MAX_AMAPS];
  CUtensorMap b_map;
  CUtensorMap out_map, add_map, mask_map;   // epilogue slabs: box {64, bw, bh, 1}
  SegDev seg[URSO_MAX_SEGS];
  int n_seg;
  int OW, OH, NB;
  int TW, TH, tw_shift;
  int tiles_w, tiles_h, n_tiles_n, total_tiles;
  FastDiv fd_ntn, fd_tw, fd_th;
  int ncols;
  int stages;        // mainloop ring depth (runtime: depends on how much smem the epilogue needs)
  int epi_tma;       // 1: TMA epilogue, 0: legacy register epilogue
  int has_add, has_mask;
  int ei_depth;      // per-warp prefetch ring depth of the epilogue inputs (chunks ahead)
  int eo_depth;      // per-warp output slabs (TMA stores in flight)
  int ei_off, eo_off;  // byte offsets of the epilogue input ring / output slabs from the aligned smem base
  int colacc_bytes;    // per-CTA column-sum accumulator (0 when the launch has no colsum: the space goes to stages)
  PixDev out, addend, mask;
  int out_fp32, relu;
  const float* shift;
  float* colsum;
  int dbg_row_shift, dbg_base_offset;   // descriptor experiments (scripts/exp_desc_shift.py)
  int dbg_no_tma, dbg_no_mma;            // bottleneck isolation (URSO_DBG_NO_TMA / URSO_DBG_NO_MMA): results are garbage
  // halo mode (3x3-style taps on one stride-1 view, TW == 8): ONE TMA box per channel chunk holds the whole
  // (TH+dh range) x (TW+dw range) pixel halo; every tap's A operand is a row-shifted window of it (UMMA descriptors
  // swizzle on absolute smem address bits, so a start shifted by whole 128-byte rows reads what TMA wrote).
  CUtensorMap a_halo_map;
  int halo, halo_w, halo_dw_min, halo_dh_min, halo_bytes;
  int a_stages, a_stage_bytes, a_ring_bytes;
  // cluster mode: CTA pairs work on two adjacent M tiles of the same N tile; each CTA fetches HALF of the weight tile
  // and multicasts it to both (halves the L2 -> SM traffic of B, which bounds the large-N tiles)
  CUtensorMap b_half_map;
  int cluster, total_pairs;
  int cta2;   // CTA pairs with tcgen05.mma.cta_group::2 (BLOCK_N = 256 only; see the kernel)
  int kpack;  // K steps (of 64) per pipeline stage / barrier round: 2 halves the per-K-step issue overhead of the MMA warp
              // (wait + fence + elect + commit ~ 190 cycles, scripts/umma_rate.cu), which bounds the N <= 128 launches
};

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kATileBytes = kBlockM * kBlockK * 2;  // 16 KB
constexpr int kMaxStages = 8;
constexpr int kCtrlBytes = 1024;                    // barriers + tmem slot
constexpr int kMaxColsumCols = 2048;
constexpr int kColsumAccBytes = kMaxColsumCols * 4;
constexpr int kSlabBytes = 32 * 128;                // 32 pixel rows x 64 bf16 channels
constexpr int kMaxEiDepth = 3;                      // per-warp prefetch ring depth (chunks ahead), runtime <= this
constexpr int kTrStride = 36;                       // legacy colsum transpose scratch (floats per row)
constexpr int kLegacyScratchBytes = 4 * 32 * kTrStride * 4;
constexpr int kSmemBudget = 227 * 1024;             // the dynamic smem base is 1 KB aligned by declaration: no slack

__device__ __forceinline__ void decode_tile(const ConvGemmParams& p, int tile, int& n_tile, int& img, int& h0,
                                            int& w0) {
  const uint32_t m_tile = fdiv((uint32_t)tile, p.fd_ntn);
  n_tile = tile - (int)m_tile * p.n_tiles_n;
  const uint32_t rest = fdiv(m_tile, p.fd_tw);
  const int twi = (int)(m_tile - rest * (uint32_t)p.tiles_w);
  const uint32_t im = fdiv(rest, p.fd_th);
  const int thi = (int)(rest - im * (uint32_t)p.tiles_h);
  img = (int)im;
  h0 = thi * p.TH;
  w0 = twi * p.TW;
}

// Work items: tiles (one CTA each) or, in cluster mode, tile PAIRS (m_tile = 2*pair + cluster rank, same n_tile).
__device__ __forceinline__ int work_first(const ConvGemmParams& p) { return p.cluster ? (int)(blockIdx.x >> 1) : (int)blockIdx.x; }
__device__ __forceinline__ int work_step(const ConvGemmParams& p) { return p.cluster ? (int)(gridDim.x >> 1) : (int)gridDim.x; }
__device__ __forceinline__ int work_end(const ConvGemmParams& p) { return p.cluster ? p.total_pairs : p.total_tiles; }
__device__ __forceinline__ int work_tile(const ConvGemmParams& p, int wk) {
  if (!p.cluster) return wk;
  const int ntn = p.n_tiles_n;
  const int q = (int)fdiv((uint32_t)wk, p.fd_ntn);
  return (2 * q + (int)(blockIdx.x & 1)) * ntn + (wk - q * ntn);   // may lie beyond the last tile: an all-OOB dummy
}

__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// fp32 pair -> packed bf16 with the ReLU folded into the conversion (cvt.rn.relu.bf16x2.f32: one F2FP instead of F2FP + HMNMX2)
__device__ __forceinline__ uint32_t pack_bf16_relu(float a, float b) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}

// Position of an epilogue warp in its stream of 64-channel chunks (tiles of this CTA x chunks per tile).
template <int BLOCK_N>
struct ChunkIter {
  int wk, tile, j, nch, n_tile, img, h0, w0;
  bool valid;
  __device__ __forceinline__ void load(const ConvGemmParams& p) {
    valid = wk < work_end(p);
    if (valid) {
      tile = work_tile(p, wk);
      decode_tile(p, tile, n_tile, img, h0, w0);
      const int rem = p.ncols - n_tile * BLOCK_N;
      nch = (rem < BLOCK_N ? rem : BLOCK_N) >> 6;
    }
  }
  __device__ __forceinline__ void init(const ConvGemmParams& p) {
    wk = work_first(p);
    j = 0;
    load(p);
  }
  __device__ __forceinline__ void next(const ConvGemmParams& p) {
    if (++j >= nch) {
      j = 0;
      wk += work_step(p);
      load(p);
    }
  }
};

// CTA2 = true: CTA pairs (cluster of 2) issue ONE tcgen05.mma.cta_group::2 per K = 16 slice over two adjacent M tiles of the
// same N tile (M = 256): each CTA stages its own A tile and HALF of the weight tile, so the per-SM operand ingest of a
// 128x256 tile drops from 48 to 32 KB per K step.  The leader CTA (cluster rank 0) owns the full / tempty barriers and
// issues the MMAs; TMA loads of both CTAs signal the leader's full barrier; commits arrive on both CTAs' barriers.
template <int BLOCK_N, bool CTA2>
__global__ void __launch_bounds__(384, 1) conv_gemm_kernel(const __grid_constant__ ConvGemmParams p) {
  constexpr int kBTileBytes = (CTA2 ? BLOCK_N / 2 : BLOCK_N) * kBlockK * 2;
  constexpr int kTmemCols = 2 * BLOCK_N;
  const uint32_t pair_rank = CTA2 ? cluster_ctarank() : 0u;
  // 1024-byte aligned by declaration (SWIZZLE_128B atoms): no integer round trip on the base pointer, so the compiler
  // keeps the shared address space and emits LDS/STS/ATOMS with 32-bit addresses instead of generic accesses
  extern __shared__ __align__(1024) uint8_t smem[];
  const int kStages = p.stages;
  uint8_t* sA = smem;
  uint8_t* sB = smem + p.a_ring_bytes;
  uint8_t* ctrl = sB + kStages * p.kpack * kBTileBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ctrl);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tfull_bar = empty_bar + kMaxStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* ei_bar = tempty_bar + 2;                       // [8 warps][kMaxEiDepth]
  uint64_t* afull_bar = ei_bar + 8 * kMaxEiDepth;          // halo mode: A ring barriers
  uint64_t* aempty_bar = afull_bar + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aempty_bar + 4);
  float* s_colacc = reinterpret_cast<float*>(ctrl + kCtrlBytes);
  float* s_tr = reinterpret_cast<float*>(ctrl + kCtrlBytes + p.colacc_bytes);   // legacy epilogue only

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < URSO_MAX_AMAPS; ++i) tma_prefetch_desc(&p.a_maps[i]);
    tma_prefetch_desc(&p.b_map);
    if (p.halo) tma_prefetch_desc(&p.a_halo_map);
    if (p.epi_tma) {
      tma_prefetch_desc(&p.out_map);
      tma_prefetch_desc(&p.add_map);
      tma_prefetch_desc(&p.mask_map);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], CTA2 ? 2 : 1);         // CTA2: one expect_tx arrival per CTA of the pair (leader's barrier)
      mbar_init(&empty_bar[i], (p.cluster && !CTA2) ? 2 : 1);   // multicast mode: both CTAs' MMAs must have consumed it
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], CTA2 ? 16 : (p.epi_tma ? 8 : 4));   // CTA2: the epilogue warps of BOTH CTAs (leader's barrier)
    }
    for (int i = 0; i < 8 * kMaxEiDepth; ++i) mbar_init(&ei_bar[i], 1);
    for (int i = 0; i < 4; ++i) {
      mbar_init(&afull_bar[i], 1);
      mbar_init(&aempty_bar[i], 1);
    }
    fence_barrier_init();
  }
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();   // the swizzled layouts assume it
  if (warp == 2) {
    if constexpr (CTA2) tmem_alloc_2cta(tmem_slot, kTmemCols);
    else tmem_alloc(tmem_slot, kTmemCols);
  }
  if (p.colsum != nullptr) {
    for (int c = threadIdx.x; c < p.ncols; c += blockDim.x) s_colacc[c] = 0.0f;
  }
  tc_fence_before();
  if (p.cluster) cluster_sync_all();   // the peer's barriers must be initialised before anything is multicast to it
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Producer and MMA issuer: ALL lanes of the warp run the (warp-uniform) control flow and barrier waits; one elected
  // lane issues the TMA / tcgen05 instructions.  Keeping the loop state warp-uniform lets the compiler hold it in
  // uniform registers -- a loop that lives inside `if (lane == 0)` costs ~130 SASS instructions per K step in R2UR /
  // ELECT shuffling and made the single issuing thread the bottleneck of every small-N tile.
  if (warp == 0 || warp == 3) {
    // ------------------------------------------------------------------ TMA producers (two warps: even / odd K steps;
    // a single issuing thread cannot arm a barrier and issue two TMA loads per 128-cycle K step of a small-N tile)
    const int par = warp == 0 ? 0 : 1;
    int g = 0;   // global K-step counter of this CTA
    if (p.halo) {
      int stage = 0, a_stage = 0;
      uint32_t phase = 0, a_phase = 0;
      const int c_chunks = p.seg[0].c_chunks;
      for (int wk = work_first(p); wk < work_end(p); wk += work_step(p)) {
        int n_tile, img, h0, w0;
        decode_tile(p, work_tile(p, wk), n_tile, img, h0, w0);
        for (int c = 0; c < c_chunks; ++c) {
          if (par == 0) mbar_wait(&aempty_bar[a_stage], a_phase ^ 1);
          if (par == 0 && elect_one()) {
            mbar_arrive_expect_tx(&afull_bar[a_stage], p.halo_bytes);
            tma_load_4d(sA + a_stage * p.a_stage_bytes, &p.a_halo_map, &afull_bar[a_stage], c * kBlockK,
                        w0 + p.halo_dw_min, h0 + p.halo_dh_min, img);
          }
          __syncwarp();
          if (++a_stage == p.a_stages) {
            a_stage = 0;
            a_phase ^= 1;
          }
          for (int s = 0; s < p.n_seg; ++s, ++g) {
            if ((g & 1) == par) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              if (elect_one()) {
                mbar_arrive_expect_tx(&full_bar[stage], kBTileBytes);
                tma_load_2d(sB + stage * kBTileBytes, &p.b_map, &full_bar[stage], (s * c_chunks + c) * kBlockK,
                            n_tile * BLOCK_N);
              }
              __syncwarp();
            }
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    } else if (p.kpack >= 2) {
      // kpack (2 or 4) K steps per stage: the owning producer warp (stage groups alternate between the two) arms the barrier once
      // with the bytes of both K steps (one at the odd end of a tile) and issues their four tile loads
      int stage = 0, gg = 0;
      uint32_t phase = 0;
      const int kp = p.kpack;
      int ksteps = 0;
      for (int s = 0; s < p.n_seg; ++s) ksteps += p.seg[s].c_chunks;
      for (int wk = work_first(p); wk < work_end(p); wk += work_step(p)) {
        int n_tile, img, h0, w0;
        decode_tile(p, work_tile(p, wk), n_tile, img, h0, w0);
        int kcol = 0, ks = 0, slot = 0;
        for (int s = 0; s < p.n_seg; ++s) {
          const SegDev sg = p.seg[s];
          for (int c = 0; c < sg.c_chunks; ++c, ++ks) {
            if ((gg & 1) == par) {
              if (slot == 0) mbar_wait(&empty_bar[stage], phase ^ 1);
              if (elect_one()) {
                if (slot == 0) {
                  const int n_grp = ksteps - ks < kp ? ksteps - ks : kp;      // K steps in this barrier round
                  mbar_arrive_expect_tx(&full_bar[stage], n_grp * (kATileBytes + kBTileBytes));
                }
                tma_load_4d(sA + (stage * kp + slot) * kATileBytes, &p.a_maps[sg.map_id], &full_bar[stage], c * kBlockK,
                            w0 + sg.dw, h0 + sg.dh, img);
                tma_load_2d(sB + (stage * kp + slot) * kBTileBytes, &p.b_map, &full_bar[stage], kcol, n_tile * BLOCK_N);
              }
              __syncwarp();
            }
            kcol += kBlockK;
            if (++slot == kp || ks + 1 == ksteps) {
              slot = 0;
              ++gg;
              if (++stage == kStages) {
                stage = 0;
                phase ^= 1;
              }
            }
          }
        }
      }
    } else {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t crank = p.cluster ? (blockIdx.x & 1) : 0;
      for (int wk = work_first(p); wk < work_end(p); wk += work_step(p)) {
        int n_tile, img, h0, w0;
        decode_tile(p, work_tile(p, wk), n_tile, img, h0, w0);
        int kcol = 0;
        for (int s = 0; s < p.n_seg; ++s) {
          const SegDev sg = p.seg[s];
          for (int c = 0; c < sg.c_chunks; ++c, ++g) {
            if ((g & 1) == par) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              if constexpr (CTA2) {
                if (elect_one()) {
                  // both CTAs signal the LEADER's full barrier (shared::cluster address of rank 0)
                  const uint32_t fb = mapa_u32(smem_u32(&full_bar[stage]), 0);
#if URSO_DEBUG_KNOBS
                  if (p.dbg_no_tma && g >= kStages) {
                    mbar_arrive_cluster(fb);      // isolation: keep the handshake, skip the loads
                  } else
#endif
                  {
                  mbar_arrive_expect_tx_cluster(fb, kATileBytes + kBTileBytes);
                  tma_load_4d_2cta(sA + stage * kATileBytes, &p.a_maps[sg.map_id], fb, c * kBlockK, w0 + sg.dw, h0 + sg.dh,
                                   img);
                  tma_load_2d_2cta(sB + stage * kBTileBytes, &p.b_half_map, fb, kcol,
                                   n_tile * BLOCK_N + (int)pair_rank * (BLOCK_N / 2));
                  }
                }
              } else
#if URSO_DEBUG_KNOBS
              if (p.dbg_no_tma && g >= kStages) {
                if (elect_one()) mbar_arrive(&full_bar[stage]);
              } else
#endif
              if (elect_one()) {
                mbar_arrive_expect_tx(&full_bar[stage], kATileBytes + kBTileBytes);
                tma_load_4d(sA + stage * kATileBytes, &p.a_maps[sg.map_id], &full_bar[stage], c * kBlockK, w0 + sg.dw,
                            h0 + sg.dh, img);
                if (p.cluster)   // my half of the weight tile, delivered to both CTAs of the pair
                  tma_load_2d_mcast(sB + stage * kBTileBytes + crank * (kBTileBytes / 2), &p.b_half_map, &full_bar[stage],
                                    kcol, n_tile * BLOCK_N + crank * (BLOCK_N / 2), (uint16_t)3);
                else
                  tma_load_2d(sB + stage * kBTileBytes, &p.b_map, &full_bar[stage], kcol, n_tile * BLOCK_N);
              }
              __syncwarp();
            }
            kcol += kBlockK;
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, BLOCK_N, 0, 0);
    // descriptor = constant high word | (smem address >> 4); advancing K by 16 elements (32 B) adds 2 to the low word
    constexpr uint64_t kDescHiB = (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 16) | (1ull << 46) | (2ull << 61);
    const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
    if (p.halo) {
      int stage = 0, a_stage = 0;
      uint32_t phase = 0, a_phase = 0;
      const int c_chunks = p.seg[0].c_chunks;
      // next 8-pixel group = next patch row = halo_w smem rows further
      const uint64_t desc_hi_a = (uint64_t((p.halo_w * 128) >> 4) << 32) | (uint64_t(1) << 16) | (1ull << 46) | (2ull << 61);
      int it = 0;
      for (int wk = work_first(p); wk < work_end(p); wk += work_step(p), ++it) {
        const int as = it & 1;
        mbar_wait(&tempty_bar[as], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BLOCK_N;
        for (int c = 0; c < c_chunks; ++c) {
          mbar_wait(&afull_bar[a_stage], a_phase);
          const uint32_t a_tile = a_base + a_stage * p.a_stage_bytes;
          for (int s = 0; s < p.n_seg; ++s) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            if (elect_one()) {
              const SegDev sg = p.seg[s];
              const uint32_t a_addr = a_tile + ((sg.dh - p.halo_dh_min) * p.halo_w + (sg.dw - p.halo_dw_min)) * 128;
              const uint64_t ad = desc_hi_a | (a_addr >> 4);
              const uint64_t bd = kDescHiB | ((b_base + stage * kBTileBytes) >> 4);
              umma_bf16(d_tmem, ad, bd, idesc, (c | s) != 0);
#pragma unroll
              for (int k = 1; k < kBlockK / 16; ++k) umma_bf16(d_tmem, ad + 2 * k, bd + 2 * k, idesc, 1u);
              umma_commit(&empty_bar[stage]);
              if (s == p.n_seg - 1) umma_commit(&aempty_bar[a_stage]);   // halo tile free once all taps retire
            }
            __syncwarp();
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
          if (++a_stage == p.a_stages) {
            a_stage = 0;
            a_phase ^= 1;
          }
        }
        if (elect_one()) umma_commit(&tfull_bar[as]);
        __syncwarp();
      }
    } else if (p.kpack >= 2) {
      int stage = 0;
      uint32_t phase = 0;
      const int kp = p.kpack;
      int ksteps = 0;
      for (int s = 0; s < p.n_seg; ++s) ksteps += p.seg[s].c_chunks;
      int it = 0;
      for (int wk = work_first(p); wk < work_end(p); wk += work_step(p), ++it) {
        const int as = it & 1;
        mbar_wait(&tempty_bar[as], ((it >> 1) & 1) ^ 1);  // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BLOCK_N;
        for (int ks = 0; ks < ksteps; ks += kp) {
          const int n = ksteps - ks < kp ? ksteps - ks : kp;
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            for (int j = 0; j < n; ++j) {
              const uint64_t ad = kDescHiB | ((a_base + (stage * kp + j) * kATileBytes) >> 4);
              const uint64_t bd = kDescHiB | ((b_base + (stage * kp + j) * kBTileBytes) >> 4);
              umma_bf16(d_tmem, ad, bd, idesc, (ks + j) != 0);
#pragma unroll
              for (int k = 1; k < kBlockK / 16; ++k) umma_bf16(d_tmem, ad + 2 * k, bd + 2 * k, idesc, 1u);
            }
            umma_commit(&empty_bar[stage]);   // one commit per kpack K steps
          }
          __syncwarp();
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) umma_commit(&tfull_bar[as]);
        __syncwarp();
      }
    } else if (!CTA2 || pair_rank == 0) {     // CTA pairs: the leader issues for both CTAs
      int stage = 0;
      uint32_t phase = 0;
      int ksteps = 0;
      for (int s = 0; s < p.n_seg; ++s) ksteps += p.seg[s].c_chunks;
      constexpr uint32_t idesc2 = umma_idesc_bf16(2 * kBlockM, BLOCK_N, 0, 0);   // M = 256 across the pair
      int it = 0;
      for (int wk = work_first(p); wk < work_end(p); wk += work_step(p), ++it) {
        const int as = it & 1;
        mbar_wait(&tempty_bar[as], ((it >> 1) & 1) ^ 1);  // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BLOCK_N;
        for (int ks = 0; ks < ksteps; ++ks) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t ad = kDescHiB | ((a_base + stage * kATileBytes + p.dbg_row_shift * 128) >> 4);
            const uint64_t bd = kDescHiB | ((b_base + stage * kBTileBytes) >> 4);
            if constexpr (CTA2) {
#if URSO_DEBUG_KNOBS
              if (!p.dbg_no_mma)
#endif
              {
                umma_bf16_2cta(d_tmem, ad, bd, idesc2, ks != 0);
#pragma unroll
                for (int k = 1; k < kBlockK / 16; ++k) umma_bf16_2cta(d_tmem, ad + 2 * k, bd + 2 * k, idesc2, 1u);
              }
              umma_commit_2cta(&empty_bar[stage]);      // the slot is free in BOTH CTAs once these MMAs retire
            } else {
#if URSO_DEBUG_KNOBS
              if (!p.dbg_no_mma)
#endif
              {
                umma_bf16(d_tmem, ad, bd, idesc, ks != 0);
#pragma unroll
                for (int k = 1; k < kBlockK / 16; ++k) umma_bf16(d_tmem, ad + 2 * k, bd + 2 * k, idesc, 1u);
              }
              // smem slot reusable once these MMAs retire (cluster: tell the peer too -- it multicasts into my stage)
              if (p.cluster) umma_commit_mcast(&empty_bar[stage], (uint16_t)3);
              else umma_commit(&empty_bar[stage]);
            }
          }
          __syncwarp();
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) {                              // accumulator complete -> epilogue (of both CTAs of a pair)
          if constexpr (CTA2) umma_commit_2cta(&tfull_bar[as]);
          else umma_commit(&tfull_bar[as]);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4 && (p.epi_tma || warp < 8)) {
    // ------------------------------------------------------------------ epilogue
    // TMA flavour: 8 warps = 4 TMEM lane quadrants x 2 column halves (two warps share each 32-row x 64-channel
    // slab: more warps per scheduler hide the ALU latency of the elementwise work).  Legacy flavour: warps 4-7 only.
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const int half = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    const int rh = row >> p.tw_shift, rw = row & (p.TW - 1);
    if (p.epi_tma) {
      // ---------------- TMA epilogue.  8 warps: warp set `half` (0/1) of quadrant q takes every other 64-channel
      // chunk of the quadrant's 32 pixel rows.  The two warps that share a scheduler (and a TMEM lane quadrant) work
      // on DIFFERENT chunks, each with private smem rings and its own TMA queue: no cross-warp synchronisation.
      const int slab_h = (q * 32) >> p.tw_shift, slab_w = (q * 32) & (p.TW - 1);   // slab origin inside the patch
      const int n_in = p.has_add + p.has_mask;
      const int slot_bytes = n_in * kSlabBytes;
      const int kEiDepth = p.ei_depth, kEoDepth = p.eo_depth;
      const int wslot = half * 4 + q;                                    // this warp's ring index
      uint8_t* ei = smem + p.ei_off + wslot * kEiDepth * slot_bytes;
      uint8_t* eo = smem + p.eo_off + wslot * kEoDepth * kSlabBytes;
      uint64_t* my_bar = ei_bar + wslot * kMaxEiDepth;
      // byte offsets of this lane's eight 16-byte units inside a 128-byte swizzled row (unit u of row r sits at u ^ (r & 7))
      int uoff[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) uoff[u] = lane * 128 + ((u ^ (lane & 7)) << 4);
      ChunkIter<BLOCK_N> cur, pf;
      cur.init(p);
      if (half == 1 && cur.valid) cur.next(p);     // set 1 starts at the second chunk of the stream
      pf = cur;
      int pf_slot = 0;                             // ring slot of the next prefetch
      auto issue_prefetch = [&]() {                // lane 0 only
        const int slot = pf_slot;
        if (++pf_slot == kEiDepth) pf_slot = 0;
        uint8_t* dst = ei + slot * slot_bytes;
        mbar_arrive_expect_tx(&my_bar[slot], slot_bytes);
        const int c = pf.n_tile * BLOCK_N + pf.j * 64;
        if (p.has_add) tma_load_4d(dst, &p.add_map, &my_bar[slot], c, pf.w0 + slab_w, pf.h0 + slab_h, pf.img);
        if (p.has_mask)
          tma_load_4d(dst + p.has_add * kSlabBytes, &p.mask_map, &my_bar[slot], c, pf.w0 + slab_w, pf.h0 + slab_h, pf.img);
        pf.next(p);
        if (pf.valid) pf.next(p);                  // my chunks are every other one
      };
      if (n_in > 0 && lane == 0) {
        for (int i = 0; i < kEiDepth && pf.valid; ++i) issue_prefetch();
      }
      int n_done = 0;        // my chunks consumed so far
      int in_slot = 0, out_slot = 0;   // ring positions (kept incrementally: no runtime modulo on the critical path)
      uint32_t in_phase = 0;
      int it = 0;            // tiles of this CTA visited
      // every tile of the CTA is visited by BOTH warp sets (each must release the accumulator stage exactly once)
      for (int wk = work_first(p); wk < work_end(p); wk += work_step(p), ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait_relaxed(&tfull_bar[as], aphase);
        tc_fence_after();
        while (cur.valid && cur.wk == wk) {
          const int j = cur.j;
          const int col0 = cur.n_tile * BLOCK_N + j * 64;
          const uint8_t* in_slab = ei + in_slot * slot_bytes;
          uint8_t* out_slab = eo + out_slot * kSlabBytes;
          // rows outside the image are clipped by the TMA store; they only have to be zeroed for the column sums
          const bool valid = p.colsum == nullptr ||
                             ((cur.h0 + rh < p.OH) && (cur.w0 + rw < p.OW) && (cur.img < p.NB));
          if (n_done >= kEoDepth) {   // the TMA store that last used this output slab must have drained it
            if (lane == 0) {
              if (kEoDepth >= 3) tma_store_wait_read<2>();
              else if (kEoDepth == 2) tma_store_wait_read<1>();
              else tma_store_wait_read<0>();
            }
            __syncwarp();
          }
          if (n_in > 0) mbar_wait(&my_bar[in_slot], in_phase);
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            uint32_t acc[32];
            tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + as * BLOCK_N + j * 64 + hf * 32, acc);
            tmem_ld_wait();
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[i]);
            if (p.shift != nullptr) {
              const float4* sp = reinterpret_cast<const float4*>(p.shift + col0 + hf * 32);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 s4 = __ldg(sp + i);
                v[4 * i + 0] += s4.x;
                v[4 * i + 1] += s4.y;
                v[4 * i + 2] += s4.z;
                v[4 * i + 3] += s4.w;
              }
            }
            if (p.has_add) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint4 u = *reinterpret_cast<const uint4*>(in_slab + uoff[hf * 4 + i]);
                v[8 * i + 0] += bf16_lo(u.x);
                v[8 * i + 1] += bf16_hi(u.x);
                v[8 * i + 2] += bf16_lo(u.y);
                v[8 * i + 3] += bf16_hi(u.y);
                v[8 * i + 4] += bf16_lo(u.z);
                v[8 * i + 5] += bf16_hi(u.z);
                v[8 * i + 6] += bf16_lo(u.w);
                v[8 * i + 7] += bf16_hi(u.w);
              }
            }
            // pack to bf16 first; ReLU and the ReLU-backward mask are exact on the packed values
            uint32_t pk[16];
            if (p.relu) {
#pragma unroll
              for (int i = 0; i < 16; ++i) pk[i] = pack_bf16_relu(v[2 * i], v[2 * i + 1]);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) pk[i] = pack_bf16(v[2 * i], v[2 * i + 1]);
            }
            if (p.has_mask) {
              const uint8_t* ms = in_slab + p.has_add * kSlabBytes;
              const __nv_bfloat162 z2 = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint4 u = *reinterpret_cast<const uint4*>(ms + uoff[hf * 4 + i]);
                pk[4 * i + 0] &= __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&u.x), z2);
                pk[4 * i + 1] &= __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&u.y), z2);
                pk[4 * i + 2] &= __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&u.z), z2);
                pk[4 * i + 3] &= __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&u.w), z2);
              }
            }
            if (!valid) {
#pragma unroll
              for (int i = 0; i < 16; ++i) pk[i] = 0u;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
              *reinterpret_cast<uint4*>(out_slab + uoff[hf * 4 + i]) =
                  make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
          }
          fence_proxy_async();   // generic-proxy smem writes -> visible to the TMA (async proxy)
          __syncwarp();          // all lanes have finished reading the input slot and writing the output slab
          if (lane == 0) {
            tma_store_4d(&p.out_map, out_slab, col0, cur.w0 + slab_w, cur.h0 + slab_h, cur.img);
            tma_store_commit();
            if (n_in > 0 && pf.valid) issue_prefetch();   // refill the input slot just consumed
          }
          if (p.colsum != nullptr) {
            // lane l sums channel pair l of this chunk over the 32 rows of the bf16 output slab (conflict free:
            // at a fixed row the 32 lanes read the 32 distinct words of one 128-byte line)
            float s0 = 0.f, s1 = 0.f;
            const int c16 = lane >> 2, wsel = (lane & 3) << 2;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
              const uint32_t u = *reinterpret_cast<const uint32_t*>(out_slab + r * 128 + ((c16 ^ (r & 7)) << 4) + wsel);
              s0 += bf16_lo(u);
              s1 += bf16_hi(u);
            }
            atomicAdd(&s_colacc[col0 + 2 * lane], s0);
            atomicAdd(&s_colacc[col0 + 2 * lane + 1], s1);
          }
          ++n_done;
          if (++in_slot == kEiDepth) {
            in_slot = 0;
            in_phase ^= 1;
          }
          if (++out_slot == kEoDepth) out_slot = 0;
          cur.next(p);
          if (cur.valid) cur.next(p);   // skip the other warp set's chunk
        }
        // all of this warp's TMEM reads of the tile are complete: release the accumulator stage (8 arrivals)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CTA2) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[as]), 0));   // the leader's barrier
          else mbar_arrive(&tempty_bar[as]);
        }
      }
      if (lane == 0) tma_store_wait_all();
    } else {
      // ---------------- legacy register epilogue (fp32 output / BLOCK_N == 32)
      int it = 0;
      for (int wk = work_first(p); wk < work_end(p); wk += work_step(p), ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        int n_tile, img, h0, w0;
        decode_tile(p, work_tile(p, wk), n_tile, img, h0, w0);
        const int h = h0 + rh, w = w0 + rw;
        const bool valid = (h < p.OH) && (w < p.OW) && (img < p.NB);
        const long long o_off = (long long)img * p.out.sn + (long long)h * p.out.sh + (long long)w * p.out.sw;
        const long long a_off = (long long)img * p.addend.sn + (long long)h * p.addend.sh + (long long)w * p.addend.sw;
        const long long m_off = (long long)img * p.mask.sn + (long long)h * p.mask.sh + (long long)w * p.mask.sw;
        mbar_wait(&tfull_bar[as], aphase);
        tc_fence_after();
#pragma unroll 1
        for (int j = 0; j < BLOCK_N / 32; ++j) {
          const int col0 = n_tile * BLOCK_N + j * 32;
          if (col0 >= p.ncols) break;  // warp-uniform
          uint32_t acc[32];
          tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + as * BLOCK_N + j * 32, acc);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[i]);
          if (p.shift != nullptr) {
            const float4* sp = reinterpret_cast<const float4*>(p.shift + col0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float4 s4 = __ldg(sp + i);
              v[4 * i + 0] += s4.x;
              v[4 * i + 1] += s4.y;
              v[4 * i + 2] += s4.z;
              v[4 * i + 3] += s4.w;
            }
          }
          if (p.addend.ptr != nullptr && valid) {
            const uint4* ap = reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(p.addend.ptr) + a_off + col0);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4 u = ap[i];
              v[8 * i + 0] += bf16_lo(u.x);
              v[8 * i + 1] += bf16_hi(u.x);
              v[8 * i + 2] += bf16_lo(u.y);
              v[8 * i + 3] += bf16_hi(u.y);
              v[8 * i + 4] += bf16_lo(u.z);
              v[8 * i + 5] += bf16_hi(u.z);
              v[8 * i + 6] += bf16_lo(u.w);
              v[8 * i + 7] += bf16_hi(u.w);
            }
          }
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
          }
          if (p.mask.ptr != nullptr && valid) {
            const uint4* mp = reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(p.mask.ptr) + m_off + col0);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4 u = mp[i];
              v[8 * i + 0] = bf16_lo(u.x) > 0.0f ? v[8 * i + 0] : 0.0f;
              v[8 * i + 1] = bf16_hi(u.x) > 0.0f ? v[8 * i + 1] : 0.0f;
              v[8 * i + 2] = bf16_lo(u.y) > 0.0f ? v[8 * i + 2] : 0.0f;
              v[8 * i + 3] = bf16_hi(u.y) > 0.0f ? v[8 * i + 3] : 0.0f;
              v[8 * i + 4] = bf16_lo(u.z) > 0.0f ? v[8 * i + 4] : 0.0f;
              v[8 * i + 5] = bf16_hi(u.z) > 0.0f ? v[8 * i + 5] : 0.0f;
              v[8 * i + 6] = bf16_lo(u.w) > 0.0f ? v[8 * i + 6] : 0.0f;
              v[8 * i + 7] = bf16_hi(u.w) > 0.0f ? v[8 * i + 7] : 0.0f;
            }
          }
          if (valid) {
            if (p.out_fp32) {
              float4* op = reinterpret_cast<float4*>(static_cast<float*>(p.out.ptr) + o_off + col0);
#pragma unroll
              for (int i = 0; i < 8; ++i) op[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            } else {
              uint4* op = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out.ptr) + o_off + col0);
#pragma unroll
              for (int i = 0; i < 4; ++i)
                op[i] = make_uint4(pack_bf16(v[8 * i], v[8 * i + 1]), pack_bf16(v[8 * i + 2], v[8 * i + 3]),
                                   pack_bf16(v[8 * i + 4], v[8 * i + 5]), pack_bf16(v[8 * i + 6], v[8 * i + 7]));
            }
          }
          if (p.colsum != nullptr) {
            float* tr = s_tr + q * 32 * kTrStride;
#pragma unroll
            for (int i = 0; i < 8; ++i)
              *reinterpret_cast<float4*>(tr + lane * kTrStride + 4 * i) =
                  valid ? make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
            __syncwarp();
            float s = 0.0f;
#pragma unroll
            for (int r = 0; r < 32; ++r) s += tr[r * kTrStride + lane];
            atomicAdd(&s_colacc[col0 + lane], s);
            __syncwarp();
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[as]);
      }
    }
    if (p.colsum != nullptr) {
      const int n_epi = p.epi_tma ? 256 : 128;         // the epilogue warps only
      asm volatile("bar.sync 5, %0;" ::"r"(n_epi) : "memory");
      for (int c = threadIdx.x - 128; c < p.ncols; c += n_epi) {
        const float s = s_colacc[c];
        if (s != 0.0f) atomicAdd(p.colsum + c, s);
      }
    }
  }

  tc_fence_before();
  if (p.cluster) cluster_sync_all();   // no CTA may exit while its peer can still multicast into it
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if constexpr (CTA2) tmem_dealloc_2cta(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace urso

// ====================================================================================== host side
struct urso_convgemm {
  urso::ConvGemmParams params;
  int block_n;
  int grid;
  int smem_bytes;
};

template <int BLOCK_N, bool CTA2 = false>
static int launch_conv_gemm(const urso_convgemm* h, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    URSO_CUDA_OK(cudaFuncSetAttribute(urso::conv_gemm_kernel<BLOCK_N, CTA2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      227 * 1024));
    attr_set = true;
  }
  if (h->params.cluster) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(h->grid);
    cfg.blockDim = dim3(384);
    cfg.dynamicSmemBytes = h->smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    URSO_CUDA_OK(cudaLaunchKernelEx(&cfg, urso::conv_gemm_kernel<BLOCK_N, CTA2>, h->params));
  } else {
    urso::conv_gemm_kernel<BLOCK_N, CTA2><<<h->grid, 384, h->smem_bytes, stream>>>(h->params);
  }
  URSO_CUDA_OK(cudaGetLastError());
  return 0;
}

static int make_pix_map(CUtensorMap* out, const urso_pix& px, int C, int W, int H, int N, int bw, int bh) {
  urso_view4 v;
  v.base = px.ptr;
  v.C = C; v.W = W; v.H = H; v.N = N;
  v.stride_w = px.sw; v.stride_h = px.sh; v.stride_n = px.sn;
  // size-1 dimensions may carry arbitrary strides (torch views): give them a legal multiple of 16 bytes
  if (H == 1) v.stride_h = (int64_t)W * px.sw;
  if (N == 1) v.stride_n = (int64_t)(H == 1 ? 1 : H) * v.stride_h;
  return urso::make_view_map(out, v, bw, bh);
}

extern "C" int urso_convgemm_create(const urso_convgemm_desc* d, urso_convgemm_t** out) {
  using namespace urso;
  URSO_REQUIRE(d != nullptr && out != nullptr, "null argument");
  URSO_REQUIRE(d->n_a >= 1 && d->n_a <= URSO_MAX_AMAPS, "n_a=%d out of range", d->n_a);
  URSO_REQUIRE(d->n_seg >= 1 && d->n_seg <= URSO_MAX_SEGS, "n_seg=%d out of range", d->n_seg);
  URSO_REQUIRE(d->TW * d->TH == 128 && (d->TW & (d->TW - 1)) == 0, "TW*TH must be 128 with TW a power of two (got %dx%d)",
               d->TW, d->TH);
  URSO_REQUIRE(d->b_rows % 32 == 0, "b_rows=%d must be a multiple of 32", d->b_rows);
  URSO_REQUIRE(d->out.ptr != nullptr, "null output");
  URSO_REQUIRE(d->colsum == nullptr || d->b_rows <= urso::kMaxColsumCols, "colsum supports at most %d channels",
               urso::kMaxColsumCols);
  auto* h = new urso_convgemm();
  ConvGemmParams& p = h->params;
  memset(&p, 0, sizeof(p));
  int ktot = 0;
  for (int s = 0; s < d->n_seg; ++s) {
    const urso_seg& sg = d->seg[s];
    if (sg.map_id < 0 || sg.map_id >= d->n_a || sg.c_chunks < 1 || sg.c_chunks * 64 > ((d->a[sg.map_id].C + 63) / 64) * 64) {
      set_error("segment %d invalid (map %d, chunks %d)", s, sg.map_id, sg.c_chunks);
      delete h;
      return 2;
    }
    p.seg[s] = SegDev{(int16_t)sg.map_id, (int16_t)sg.dh, (int16_t)sg.dw, (int16_t)sg.c_chunks};
    ktot += sg.c_chunks * 64;
  }
  if (ktot != d->b_k) {
    set_error("segments cover K=%d but b_k=%d", ktot, d->b_k);
    delete h;
    return 2;
  }
  for (int i = 0; i < URSO_MAX_AMAPS; ++i) {
    const urso_view4& v = d->a[i < d->n_a ? i : 0];
    if (int rc = make_view_map(&p.a_maps[i], v, d->TW, d->TH)) {
      delete h;
      return rc;
    }
  }
  int bn = d->block_n;
  if (bn == 0 && d->b_rows >= 256) {
    if (const char* e = getenv("URSO_BN_MAX")) bn = atoi(e);   // experiments: cap the N tile
  }
  if (bn == 0) bn = d->b_rows >= 256 && d->b_rows % 256 == 0 ? 256 : (d->b_rows >= 128 ? 128 : (d->b_rows >= 64 ? 64 : 32));
  if (bn != 32 && bn != 64 && bn != 128 && bn != 256) {
    set_error("block_n=%d unsupported", bn);
    delete h;
    return 2;
  }
  // epilogue flavour and shared-memory plan
  const int kColAcc = d->colsum != nullptr ? ((d->b_rows * 4 + 1023) / 1024) * 1024 : 0;
  p.colacc_bytes = kColAcc;
  p.has_add = d->addend.ptr != nullptr;
  p.has_mask = d->mask.ptr != nullptr;
  p.epi_tma = (!d->out_fp32 && d->b_rows % 64 == 0 && bn >= 64) ? 1 : 0;
  int epi_bytes;
  if (p.epi_tma) {
    // per-warp private rings (8 epilogue warps): input slots (prefetch depth) and output slabs (stores in flight)
    // Launches with a long K loop visit the epilogue rarely: minimal rings, smem goes to mainloop stages.  Short-K
    // launches are epilogue / store bound: deeper rings keep more TMA traffic in flight.
    const int n_in = p.has_add + p.has_mask;
    int ksteps_total = 0;
    for (int s2 = 0; s2 < d->n_seg; ++s2) ksteps_total += d->seg[s2].c_chunks;
    const bool heavy = ksteps_total >= 4;
    // A depth-1 input ring exposes the full TMA latency of every chunk, a depth-2 ring costs mainloop stages.  Launches
    // with a short K loop (<= 6 K steps: the MMAs of a tile take less time than its epilogue) are epilogue bound and get
    // the deeper ring; long-K launches keep their stages.  Measured per layer: gpurun_out s15 (profiles/r01_progress.md).
    const bool epi_bound = ksteps_total <= 6;
    int min_stages = 3;
    p.ei_depth = (heavy || n_in == 2) ? 1 : 2;
    p.eo_depth = (heavy || n_in == 2) ? 1 : 2;
    if (n_in > 0 && epi_bound) {
      p.ei_depth = 2;
      min_stages = 2;
    }
    int ei_bytes = 8 * p.ei_depth * n_in * kSlabBytes;
    epi_bytes = ei_bytes + 8 * p.eo_depth * kSlabBytes;
    {   // the narrowest tile this launch may fall back to must still get 2 mainloop stages
      const int bn_min = bn == 256 ? 128 : bn;
      if (p.ei_depth == 2 && kSmemBudget - kCtrlBytes - kColAcc - epi_bytes < 2 * (kATileBytes + bn_min * kBlockK * 2)) {
        p.ei_depth = 1;
        ei_bytes = 8 * p.ei_depth * n_in * kSlabBytes;
        epi_bytes = ei_bytes + 8 * p.eo_depth * kSlabBytes;
      }
    }
    if (bn == 256 && kSmemBudget - kCtrlBytes - kColAcc - epi_bytes < min_stages * (kATileBytes + 256 * kBlockK * 2)) bn = 128;
  } else {
    epi_bytes = kLegacyScratchBytes;
  }
  h->block_n = bn;
  // CTA pairs (cta_group::2): each CTA stages half of the weight tile.  Opt-in (URSO_CTA2=1) while it is being evaluated.
  p.cta2 = 0;
  if (const char* e = getenv("URSO_CTA2")) p.cta2 = (atoi(e) && bn == 256 && !d->halo && p.epi_tma) ? 1 : 0;
  const int b_rows_cta = p.cta2 ? bn / 2 : bn;      // rows of B resident per CTA
  p.kpack = 1;
  {
    int kst = 0;
    for (int s2 = 0; s2 < d->n_seg; ++s2) kst += d->seg[s2].c_chunks;
    // measured per layer (gpurun_out kp_*, profiles/r01_progress.md): a win for launches WITHOUT epilogue inputs and a
    // long enough K loop (stem -9 %, stage-2 3x3 -10 %, bottleneck conv -26 %), a loss where the epilogue rings already
    // squeeze the stage count (dgrad with mask / addend).  URSO_KPACK=0 disables, =2 forces it for every N <= 128 launch.
    const bool no_epi_inputs = d->addend.ptr == nullptr && d->mask.ptr == nullptr;
    // (4 K steps per round, URSO_KPACK=4, measured worse for the 9-K-step 3x3 layers: rounds of 4,4,1 on a 2-deep ring)
    int want = (no_epi_inputs && kst >= 4) ? 2 : 1;
    if (const char* e = getenv("URSO_KPACK")) want = atoi(e);
    if ((want == 2 || want == 4) && bn <= 128 && !d->halo && !p.cta2 && kst >= 2 && getenv("URSO_CLUSTER") == nullptr)
      p.kpack = want;
  }
  int stages;
  if (d->halo) {
    // validate + plan the halo ring: one box per channel chunk, B tiles in their own ring
    int dw_min = 1 << 20, dw_max = -(1 << 20), dh_min = 1 << 20, dh_max = -(1 << 20);
    bool ok = d->n_a == 1 && d->TW == 8 && d->TH == 16;
    for (int s2 = 0; s2 < d->n_seg && ok; ++s2) {
      ok = d->seg[s2].map_id == 0 && d->seg[s2].c_chunks == d->seg[0].c_chunks;
      dw_min = d->seg[s2].dw < dw_min ? d->seg[s2].dw : dw_min;
      dw_max = d->seg[s2].dw > dw_max ? d->seg[s2].dw : dw_max;
      dh_min = d->seg[s2].dh < dh_min ? d->seg[s2].dh : dh_min;
      dh_max = d->seg[s2].dh > dh_max ? d->seg[s2].dh : dh_max;
    }
    if (!ok || dw_max - dw_min > 8 || dh_max - dh_min > 16) {
      set_error("halo mode needs one stride-1 view, TW == 8, TH == 16, equal chunk counts and small tap offsets");
      delete h;
      return 2;
    }
    p.halo = 1;
    p.halo_w = 8 + dw_max - dw_min;
    const int halo_h = 16 + dh_max - dh_min;
    p.halo_dw_min = dw_min;
    p.halo_dh_min = dh_min;
    p.halo_bytes = p.halo_w * halo_h * 128;
    p.a_stage_bytes = (p.halo_bytes + 1023) / 1024 * 1024;
    p.a_stages = 3;
    p.a_ring_bytes = p.a_stages * p.a_stage_bytes;
    if (int rc = make_view_map(&p.a_halo_map, d->a[0], p.halo_w, halo_h)) {
      delete h;
      return rc;
    }
    const int btile = bn * kBlockK * 2;
    stages = (kSmemBudget - kCtrlBytes - kColAcc - epi_bytes - p.a_ring_bytes) / btile;
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages < 3) {
      set_error("halo mode: not enough shared memory for the B ring (BLOCK_N=%d)", bn);
      delete h;
      return 2;
    }
  } else {
    int stage_bytes = p.kpack * (kATileBytes + b_rows_cta * kBlockK * 2);
    stages = (kSmemBudget - kCtrlBytes - kColAcc - epi_bytes) / stage_bytes;
    while (stages < 2 && p.kpack > 1) {   // not enough room for two packed stages: fewer K steps per stage
      p.kpack /= 2;
      stage_bytes = p.kpack * (kATileBytes + b_rows_cta * kBlockK * 2);
      stages = (kSmemBudget - kCtrlBytes - kColAcc - epi_bytes) / stage_bytes;
    }
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages < 2) {
      set_error("not enough shared memory for 2 pipeline stages (BLOCK_N=%d)", bn);
      delete h;
      return 2;
    }
    p.a_ring_bytes = stages * p.kpack * kATileBytes;
  }
  p.stages = stages;
  const int fixed = p.a_ring_bytes + stages * p.kpack * (d->halo ? bn : b_rows_cta) * kBlockK * 2 + kCtrlBytes + kColAcc;
  if (p.epi_tma) {
    p.ei_off = (fixed + 1023) / 1024 * 1024;
    p.eo_off = p.ei_off + 8 * p.ei_depth * (p.has_add + p.has_mask) * kSlabBytes;
    h->smem_bytes = p.eo_off + 8 * p.eo_depth * kSlabBytes;
  } else {
    h->smem_bytes = fixed + kLegacyScratchBytes;
  }
  if (h->smem_bytes > 227 * 1024) {
    set_error("internal: shared memory plan %d bytes exceeds 227 KB", h->smem_bytes);
    delete h;
    return 2;
  }
  if (int rc = make_mat_map(&p.b_map, d->b, d->b_rows, d->b_k, bn)) {
    delete h;
    return rc;
  }
  if (p.epi_tma) {
    const int bw = d->TW < 32 ? d->TW : 32, bh = 32 / bw;
    int rc = make_pix_map(&p.out_map, d->out, d->b_rows, d->OW, d->OH, d->NB, bw, bh);
    if (!rc) rc = make_pix_map(&p.add_map, p.has_add ? d->addend : d->out, d->b_rows, d->OW, d->OH, d->NB, bw, bh);
    if (!rc) rc = make_pix_map(&p.mask_map, p.has_mask ? d->mask : d->out, d->b_rows, d->OW, d->OH, d->NB, bw, bh);
    if (rc) {
      delete h;
      return rc;
    }
  }
  p.n_seg = d->n_seg;
  p.OW = d->OW; p.OH = d->OH; p.NB = d->NB;
  p.TW = d->TW; p.TH = d->TH;
  p.tw_shift = 0;
  while ((1 << p.tw_shift) < d->TW) ++p.tw_shift;
  p.tiles_w = (d->OW + d->TW - 1) / d->TW;
  p.tiles_h = (d->OH + d->TH - 1) / d->TH;
  p.n_tiles_n = (d->b_rows + bn - 1) / bn;
  p.fd_ntn = make_fastdiv((uint32_t)p.n_tiles_n);
  p.fd_tw = make_fastdiv((uint32_t)p.tiles_w);
  p.fd_th = make_fastdiv((uint32_t)p.tiles_h);
  long long total = (long long)p.tiles_w * p.tiles_h * d->NB * p.n_tiles_n;
  if (total <= 0 || total > 0x7fffffffLL) {
    set_error("bad tile count %lld", total);
    delete h;
    return 2;
  }
  p.total_tiles = (int)total;
  {
    // cluster pairs (experimental, URSO_CLUSTER=1): measured slower on every layer -- a 2-CTA multicast does not reduce
    // L2 -> SM traffic (the L2 already de-duplicates near-simultaneous unicast requests), see profiles/r01_progress.md
    int want = 0;
    if (const char* e = getenv("URSO_CLUSTER")) want = atoi(e) && !p.halo && total >= 64;
    if (p.cta2) want = 1;
    if (want) {
      const long long mt = (long long)p.tiles_w * p.tiles_h * d->NB;
      p.cluster = 1;
      p.total_pairs = (int)(((mt + 1) / 2) * p.n_tiles_n);
      if (int rc = make_mat_map(&p.b_half_map, d->b, d->b_rows, d->b_k, bn / 2)) {
        delete h;
        return rc;
      }
    }
  }
  p.ncols = d->b_rows;
  p.out = PixDev{d->out.ptr, d->out.sn, d->out.sh, d->out.sw};
  p.addend = PixDev{d->addend.ptr, d->addend.sn, d->addend.sh, d->addend.sw};
  p.mask = PixDev{d->mask.ptr, d->mask.sn, d->mask.sh, d->mask.sw};
  p.out_fp32 = d->out_fp32;
  p.relu = d->relu;
  p.shift = d->shift;
  p.colsum = d->colsum;
  if (const char* e = getenv("URSO_DBG_NO_TMA")) p.dbg_no_tma = atoi(e);
  if (const char* e = getenv("URSO_DBG_NO_MMA")) p.dbg_no_mma = atoi(e);
  if (const char* e = getenv("URSO_DBG_ROW_SHIFT")) p.dbg_row_shift = atoi(e);
  if (const char* e = getenv("URSO_DBG_BASE_OFFSET")) p.dbg_base_offset = atoi(e);
  int sms = num_sms();
  if (sms <= 0) sms = 148;
  h->grid = p.total_tiles < sms ? p.total_tiles : sms;
  if (p.cluster) {
    const int nclusters = p.total_pairs < sms / 2 ? p.total_pairs : sms / 2;
    h->grid = 2 * nclusters;
  }
  *out = h;
  return 0;
}

extern "C" int urso_convgemm_launch(urso_convgemm_t* h, void* stream) {
  URSO_REQUIRE(h != nullptr, "null handle");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (h->block_n) {
    case 32: return launch_conv_gemm<32>(h, s);
    case 64: return launch_conv_gemm<64>(h, s);
    case 128: return launch_conv_gemm<128>(h, s);
    case 256: return h->params.cta2 ? launch_conv_gemm<256, true>(h, s) : launch_conv_gemm<256>(h, s);
  }
  urso::set_error("unsupported BLOCK_N %d", h->block_n);
  return 2;
}

extern "C" void urso_convgemm_destroy(urso_convgemm_t* h) { delete h; }
