// Engine F: persistent, warp-specialised implicit-GEMM convolution on tcgen05 tensor cores.
//
//   D[pixel, n] = sum over K-segments (filter taps / concatenated inputs) of  A_seg[pixel + shift, c] * B[n, k]
//
//  * A tiles: 128 output pixels (a TH x TW patch of one image) x 64 channels, fetched by 4-D TMA boxes straight from the
//    NHWC activation tensor.  The tap shift is added to the box coordinates and TMA's out-of-bounds zero fill implements
//    the convolution padding -- no im2col buffer ever exists.
//  * B tiles: BLOCK_N x 64 slices of the staged (BN-folded, K-major) weight matrix, 2-D TMA.
//  * both land in 128B-swizzled shared memory and are consumed by tcgen05.mma (M=128, N=BLOCK_N, K=16) issued by
//    one thread per pipeline; fp32 accumulators live in TMEM, double buffered per pipeline so the epilogue of a tile
//    overlaps the MMAs of the next.
//
//  Operand supply.  Two things bound the small-N (64 / 128 channel) layers in round 1 (profiles/r01_progress.md,
//  profiles/r02_umma_rounds.txt): (a) the MMA-issuing warp spends ~316 cycles per barrier round (wait + fence + elect +
//  commit) on top of its MMAs, and (b) the L2 -> SM path delivers ~67 B/clk per SM while a 3x3 tile of N = 64 asks for
//  128 B/clk.  This kernel answers both:
//    - TWO PIPELINES (npipe = 2, BLOCK_N <= 128): two producer warps and two MMA-issuing warps work on ALTERNATE tiles
//      of the CTA, each pair with its own shared-memory ring and its own two accumulator stages (4 x BLOCK_N TMEM
//      columns).  While one issuer sits in its barrier round the other one's MMAs keep the tensor pipe busy
//      (microbenchmark: pipe utilisation 0.38 -> 1.00 at N = 64 with 8 MMAs per round).
//    - HALO MODE (3x3-style taps on one stride-1 view, 8 x 16 pixel patch): ONE TMA box per 64-channel chunk holds the
//      patch plus its halo; every tap's A operand is a row-shifted window of it (UMMA descriptors swizzle on absolute
//      smem address bits, so a start shifted by whole 128-byte rows reads what TMA wrote) -- 6.4x less A traffic.
//    - RESIDENT B (halo mode, one N tile, weight operand <= 100 KB): the whole weight matrix is loaded once per CTA; a
//      tile then costs one barrier wait and two commits for all of its (up to 36) MMAs.
//    - kpack (stream mode): two K steps per barrier round.
//
//  * epilogue (8 warps = 4 TMEM lane quadrants x 2 alternating chunk sets), per 64-channel chunk:
//      tcgen05.ld -> +shift -> +addend -> ReLU -> mask -> bf16
//    ALL global traffic of the epilogue is TMA as well: every warp prefetches its own 32-row slab of the addend /
//    mask tensors chunks ahead into a private smem ring (so HBM latency never sits on the critical path) and writes
//    its output slab with a TMA store from a private smem slab (full 128-byte lines, automatic clipping of
//    partial tiles, strided views for stride-2 dgrad).  No cross-warp synchronisation in the epilogue.
//    A legacy register epilogue (direct vector stores) remains for fp32 outputs / BLOCK_N = 32 (bottleneck conv).
//  * optional fused per-channel sums of the stored output (d beta): read back from the bf16 output slab
//    (conflict free), accumulated per CTA in shared memory, flushed with one atomic per channel per CTA.
//
// Roles (384 threads): warps 0 and 3 = TMA producers, warps 1 and 2 = MMA issuers (warp 2 also allocates TMEM),
// warps 4-11 = epilogue.  With one pipeline the two producers take alternate barrier rounds and warp 2 only allocates.
#include "common.cuh"
#include "ptx.cuh"

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
  CUtensorMap a_maps[URSO_MAX_AMAPS];       // stream mode: box {64, TW, TH, 1} per view
  CUtensorMap a_halo_map;                   // halo mode: box {64, halo_w, halo_h, 1} of view 0
  CUtensorMap b_map;                        // box {64, BLOCK_N}
  CUtensorMap out_map, add_map, mask_map;   // epilogue slabs: box {64, bw, bh, 1}
  CUtensorMap radd_map;                     // addend as an A operand: box {64, TW, TH, 1} on the output grid
  SegDev seg[URSO_MAX_SEGS];
  int n_seg, ksteps;
  int OW, OH, NB;
  int TW, TH, tw_shift;
  int tiles_w, tiles_h, n_tiles_n, total_tiles;
  FastDiv fd_ntn, fd_tw, fd_th;
  int ncols;
  // N-SPLIT TAIL.  A launch whose tile count is not a multiple of the CTA count ends with a partial wave in which most SMs
  // idle (600 tiles on 148 CTAs: 4.05 waves of work in 5 waves of time).  For K-heavy BLOCK_N = 256 launches in stream mode
  // the R tiles of that last wave are cut along N into tail_nsub (2 or 4) sub-tiles of 128 / 64 output channels, one per CTA:
  // a sub-tile streams the same pixel tile but only its slice of the weight tile (16 + 8 KB per K step instead of 48) and
  // issues N = 64 / 128 MMAs, so the last wave costs about half a wave -- and, unlike a split along K, needs no reduction.
  // Units 0 .. tail_first-1 are whole tiles (tail_first = the full waves, a multiple of the grid); unit tail_first + u is
  // sub-tile u % tail_nsub of tile tail_first + u / tail_nsub.  Off: tail_first = n_units = total_tiles, tail_nsub = 1.
  int tail_first, tail_nsub, n_units;
  CUtensorMap b_sub_map;   // box {64, BLOCK_N / tail_nsub}
  int npipe;         // 1 or 2 producer -> MMA pipelines working on alternate tiles
  int stages;        // per pipeline: stream mode = ring of (A, B) stages; halo mode = ring of B tiles (unless resident)
  int kpack;         // stream mode: K steps (of 64) per stage / barrier round
  int halo, halo_w, halo_dw_min, halo_dh_min, halo_bytes;
  int a_stages, a_stage_bytes;   // halo mode, per pipeline
  int bres;          // halo mode: the whole weight operand is resident in shared memory
  int pipe_bytes;    // shared memory of one pipeline (its A ring followed by its B ring)
  int b_ring_off;    // offset of a pipeline's B ring from its base
  int bres_off, ctrl_off;
  int radd;          // the addend is accumulated by the tensor core: per 64-channel chunk of the tile one extra K step
                     // A = addend tile, B = 64x64 identity (resident, ident_off), N = 64 at the chunk's TMEM columns.
                     // The epilogue then has no input stream at all (no ring, no unpack, no add).  One pipeline only
                     // (BLOCK_N = 256): producer warp 3 feeds the addend tiles through their own ring of kRaddStages
                     // (barriers: the halo mode's A-ring pair), producer warp 0 owns every operand round.
  int ident_off, radd_off;
  int epi_tma;       // 1: TMA epilogue, 0: legacy register epilogue
  int has_add, has_mask;
  int ei_depth;      // per-warp prefetch ring depth of the epilogue inputs (chunks ahead)
  int eo_depth;      // per-warp output slabs (TMA stores in flight)
  int ei_off, eo_off;  // byte offsets of the epilogue input ring / output slabs from the aligned smem base
  int colacc_bytes;    // per-CTA column-sum accumulator (0 when the launch has no colsum: the space goes to stages)
  int shift_bytes;     // TMA epilogue: shared-memory copy of shift[] (the epilogue read it with 16 global loads per chunk whose
                       // L2 latency sat on every warp's critical path: ncu long-scoreboard stalls on the shift FADDs)
  PixDev out, addend, mask;
  PixDev bits_out, mask_bits;   // bit-packed ReLU masks (1 bit per element, 32-channel words): strides in BYTES
  int out_fp32, relu;
  const float* shift;
  float* colsum;
};

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kATileBytes = kBlockM * kBlockK * 2;  // 16 KB
constexpr int kMaxStages = 8;                       // per pipeline
constexpr int kCtrlBytes = 1024;                    // barriers + tmem slot
constexpr int kMaxColsumCols = 2048;
constexpr int kSlabBytes = 32 * 128;                // 32 pixel rows x 64 bf16 channels
constexpr int kMaxEiDepth = 3;                      // per-warp prefetch ring depth (chunks ahead), runtime <= this
constexpr int kTrStride = 36;                       // legacy colsum transpose scratch (floats per row)
constexpr int kLegacyScratchBytes = 4 * 32 * kTrStride * 4;
constexpr int kSmemBudget = 227 * 1024;             // the dynamic smem base is 1 KB aligned by declaration: no slack
constexpr int kMaxBresBytes = 100 * 1024;
constexpr int kRaddStages = 4;                      // addend ring: one BLOCK_N = 256 tile (4 chunks of 64 channels)

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

__device__ __forceinline__ int fdiv_ntile(const ConvGemmParams& p, int tile) {
  return tile - (int)fdiv((uint32_t)tile, p.fd_ntn) * p.n_tiles_n;
}

// work unit -> tile and N sub-tile (q = 0 for whole tiles)
__device__ __forceinline__ int unit_tile(const ConvGemmParams& p, int wk, int& q) {
  if (wk < p.tail_first) {
    q = 0;
    return wk;
  }
  const int u = wk - p.tail_first;
  const int t = p.tail_nsub == 4 ? (u >> 2) : (u >> 1);
  q = u - t * p.tail_nsub;
  return p.tail_first + t;
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
  int wk, j, nch, n_tile, img, h0, w0;
  int jb;        // first 64-channel chunk of this unit within its tile (N-split tail: sub-tile q starts at chunk q * nch)
  bool valid;
  __device__ __forceinline__ void load(const ConvGemmParams& p) {
    valid = wk < p.n_units;
    if (valid) {
      int q;
      const int tile = unit_tile(p, wk, q);
      decode_tile(p, tile, n_tile, img, h0, w0);
      const int rem = p.ncols - n_tile * BLOCK_N;
      nch = (rem < BLOCK_N ? rem : BLOCK_N) >> 6;
      jb = 0;
      if (wk >= p.tail_first) {       // (only planned when every tile is BLOCK_N channels wide)
        nch = (BLOCK_N / 64) / p.tail_nsub;
        jb = q * nch;
      }
    }
  }
  __device__ __forceinline__ void init(const ConvGemmParams& p) {
    wk = (int)blockIdx.x;
    j = 0;
    load(p);
  }
  __device__ __forceinline__ void next(const ConvGemmParams& p) {
    if (++j >= nch) {
      j = 0;
      wk += (int)gridDim.x;
      load(p);
    }
  }
};

// FLAVOR specialises the TMA epilogue at compile time (the epilogue's instruction issue bounds every short-K launch, and
// ~10 % of its instructions were uniform branches on launch flags): 0 = every option at run time; 1 = forward convs (no
// mask / mask bits / column sums compiled in); 2 = gradient launches (no shift / ReLU / mask-bit output compiled in).
template <int BLOCK_N, int FLAVOR>
__global__ void __launch_bounds__(384, 1) conv_gemm_kernel(const __grid_constant__ ConvGemmParams p) {
  constexpr bool kFwdOps = FLAVOR != 2, kBwdOps = FLAVOR != 1;
  constexpr int kBTileBytes = BLOCK_N * kBlockK * 2;
  // 1024-byte aligned by declaration (SWIZZLE_128B atoms): no integer round trip on the base pointer, so the compiler
  // keeps the shared address space and emits LDS/STS/ATOMS with 32-bit addresses instead of generic accesses
  extern __shared__ __align__(1024) uint8_t smem[];
  const int npipe = p.npipe;
  const int pshift = npipe - 1;          // npipe is 1 or 2: (it & pshift) = pipeline of tile `it`, it >> pshift = its index there
  uint8_t* ctrl = smem + p.ctrl_off;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ctrl);            // [2][kMaxStages]
  uint64_t* empty_bar = full_bar + 2 * kMaxStages;                   // [2][kMaxStages]
  uint64_t* afull_bar = empty_bar + 2 * kMaxStages;                  // [2][4]   halo mode: A ring
  uint64_t* aempty_bar = afull_bar + 8;                              // [2][4]
  uint64_t* tfull_bar = aempty_bar + 8;                              // [4] accumulator stages (2 per pipeline)
  uint64_t* tempty_bar = tfull_bar + 4;                              // [4]
  uint64_t* bres_bar = tempty_bar + 4;                               // [1] resident weight operand loaded
  uint64_t* ei_bar = bres_bar + 1;                                   // [8 warps][kMaxEiDepth]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ei_bar + 8 * kMaxEiDepth);
  float* s_colacc = reinterpret_cast<float*>(ctrl + kCtrlBytes);
  float* s_shift = reinterpret_cast<float*>(ctrl + kCtrlBytes + p.colacc_bytes);
  float* s_tr = reinterpret_cast<float*>(ctrl + kCtrlBytes + p.colacc_bytes + p.shift_bytes);   // legacy epilogue only

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < URSO_MAX_AMAPS; ++i) tma_prefetch_desc(&p.a_maps[i]);
    tma_prefetch_desc(&p.b_map);
    if (p.halo) tma_prefetch_desc(&p.a_halo_map);
    if (p.radd) tma_prefetch_desc(&p.radd_map);
    if (p.epi_tma) {
      tma_prefetch_desc(&p.out_map);
      tma_prefetch_desc(&p.add_map);
      tma_prefetch_desc(&p.mask_map);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2 * kMaxStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 8; ++i) {
      mbar_init(&afull_bar[i], 1);
      mbar_init(&aempty_bar[i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], p.epi_tma ? 8 : 4);
    }
    mbar_init(bres_bar, 1);
    for (int i = 0; i < 8 * kMaxEiDepth; ++i) mbar_init(&ei_bar[i], 1);
    fence_barrier_init();
  }
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();   // the swizzled layouts assume it
  const uint32_t tmem_cols = (uint32_t)(2 * npipe * BLOCK_N);        // 64 .. 512: a power of two
  if (warp == 2) tmem_alloc(tmem_slot, tmem_cols);
  if (p.colsum != nullptr) {
    for (int c = threadIdx.x; c < p.ncols; c += blockDim.x) s_colacc[c] = 0.0f;
  }
  if (p.shift_bytes) {
    for (int c = threadIdx.x; c < p.ncols; c += blockDim.x) s_shift[c] = __ldg(p.shift + c);
  }
  if (p.radd) {
    // 64 x 64 bf16 identity as a K-major SWIZZLE_128B operand tile: row n = 128 bytes, its 16-byte chunk j (k = 8j..8j+7)
    // stored at chunk position j ^ (n & 7)
    const uint32_t ident = smem_u32(smem + p.ident_off);
    for (int i = threadIdx.x; i < 512; i += blockDim.x) {
      const int n = i >> 3, j = (i & 7) ^ (n & 7);
      uint32_t w[4] = {0u, 0u, 0u, 0u};
      if ((n >> 3) == j) w[(n & 7) >> 1] = (n & 1) ? 0x3F800000u : 0x00003F80u;
      sts128(ident + i * 16, w[0], w[1], w[2], w[3]);
    }
    fence_proxy_async();     // generic-proxy writes -> visible to the tensor core's (async proxy) operand reads
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above ran while the previous kernel of the stream was still draining;
  // from here on its results are read.  (No-ops when the launch carries no programmatic dependency.)
  pdl_wait();
  pdl_launch_dependents();

  // Producers and MMA issuers: ALL lanes of the warp run the (warp-uniform) control flow and barrier waits; one elected
  // lane issues the TMA / tcgen05 instructions.  Keeping the loop state warp-uniform lets the compiler hold it in
  // uniform registers -- a loop that lives inside `if (lane == 0)` costs ~130 SASS instructions per K step in R2UR /
  // ELECT shuffling and made the single issuing thread the bottleneck of every small-N tile.
  if (warp == 0 || warp == 3) {
    // ------------------------------------------------------------------ TMA producers
    const int pw = warp == 0 ? 0 : 1;                 // producer index
    const int pipe = npipe == 2 ? pw : 0;             // two pipelines: one producer each; one pipeline: alternate rounds
    uint8_t* pbase = smem + pipe * p.pipe_bytes;
    uint8_t* sA = pbase;
    uint8_t* sB = pbase + p.b_ring_off;
    uint64_t* fullb = full_bar + pipe * kMaxStages;
    uint64_t* emptyb = empty_bar + pipe * kMaxStages;
    const int kStages = p.stages;
    if (p.halo) {
      uint64_t* afullb = afull_bar + pipe * 4;
      uint64_t* aemptyb = aempty_bar + pipe * 4;
      const int c_chunks = p.seg[0].c_chunks;
      if (p.bres && pw == 0) {          // the whole weight operand, once (n_tiles_n == 1)
        if (elect_one()) {
          mbar_arrive_expect_tx(bres_bar, p.ksteps * kBTileBytes);
          for (int ks = 0; ks < p.ksteps; ++ks)
            tma_load_2d(smem + p.bres_off + ks * kBTileBytes, &p.b_map, bres_bar, ks * kBlockK, 0);
        }
        __syncwarp();
      }
      int stage = 0, a_stage = 0, g = 0;
      uint32_t phase = 0, a_phase = 0;
      const bool a_mine = npipe == 2 || pw == 0;
      for (int wk = (int)blockIdx.x + pipe * (int)gridDim.x; wk < p.total_tiles; wk += npipe * (int)gridDim.x) {
        int n_tile, img, h0, w0;
        decode_tile(p, wk, n_tile, img, h0, w0);
        for (int c = 0; c < c_chunks; ++c) {
          if (a_mine) {
            mbar_wait(&aemptyb[a_stage], a_phase ^ 1);
            if (elect_one()) {
              mbar_arrive_expect_tx(&afullb[a_stage], p.halo_bytes);
              tma_load_4d(sA + a_stage * p.a_stage_bytes, &p.a_halo_map, &afullb[a_stage], c * kBlockK,
                          w0 + p.halo_dw_min, h0 + p.halo_dh_min, img);
            }
            __syncwarp();
          }
          if (++a_stage == p.a_stages) {
            a_stage = 0;
            a_phase ^= 1;
          }
          if (!p.bres) {
            for (int s = 0; s < p.n_seg; ++s, ++g) {
              if (npipe == 2 || (g & 1) == pw) {
                mbar_wait(&emptyb[stage], phase ^ 1);
                if (elect_one()) {
                  mbar_arrive_expect_tx(&fullb[stage], kBTileBytes);
                  tma_load_2d(sB + stage * kBTileBytes, &p.b_map, &fullb[stage], (s * c_chunks + c) * kBlockK,
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
      }
    } else if (p.radd && pw == 1) {
      // addend ring (one pipeline): a whole tile's 64-channel chunks ahead of the MMA warp, which consumes them after the
      // tile's operand rounds
      int slot = 0;
      uint32_t rphase = 0;
      uint8_t* ring = smem + p.radd_off;
      for (int wk = (int)blockIdx.x; wk < p.total_tiles; wk += (int)gridDim.x) {
        int n_tile, img, h0, w0;
        decode_tile(p, wk, n_tile, img, h0, w0);
        const int n0 = n_tile * BLOCK_N;
        const int rn = (p.ncols - n0 < BLOCK_N ? p.ncols - n0 : BLOCK_N) >> 6;
        for (int r = 0; r < rn; ++r) {
          mbar_wait(&aempty_bar[slot], rphase ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&afull_bar[slot], kATileBytes);
            tma_load_4d(ring + slot * kATileBytes, &p.radd_map, &afull_bar[slot], n0 + r * kBlockK, w0, h0, img);
          }
          __syncwarp();
          if (++slot == kRaddStages) {
            slot = 0;
            rphase ^= 1;
          }
        }
      }
    } else {
      // stream mode: a stage holds kpack K steps (A tile + B tile each); the owning producer arms the barrier once
      // with the bytes of the whole round and issues its tile loads
      const bool own_all = npipe == 2 || p.radd != 0;      // (with an addend ring producer 0 issues every operand round)
      int stage = 0, gg = 0;
      uint32_t phase = 0;
      const int kp = p.kpack;
      const int ksteps = p.ksteps;
      for (int wk = (int)blockIdx.x + pipe * (int)gridDim.x; wk < p.n_units; wk += npipe * (int)gridDim.x) {
        int n_tile, img, h0, w0, q;
        decode_tile(p, unit_tile(p, wk, q), n_tile, img, h0, w0);
        // N-split tail unit: only rows [q * BLOCK_N / nsub, ...) of the weight tile
        const bool sub = wk >= p.tail_first;
        const int b_bytes = sub ? kBTileBytes / p.tail_nsub : kBTileBytes;
        const int b_row0 = n_tile * BLOCK_N + (sub ? q * (BLOCK_N / p.tail_nsub) : 0);
        const CUtensorMap* bmap = sub ? &p.b_sub_map : &p.b_map;
        int kcol = 0, ks = 0, slot = 0;
        for (int s = 0; s < p.n_seg; ++s) {
          const SegDev sg = p.seg[s];
          for (int c = 0; c < sg.c_chunks; ++c, ++ks) {
            if (own_all || (gg & 1) == pw) {
              if (slot == 0) mbar_wait(&emptyb[stage], phase ^ 1);
              if (elect_one()) {
                if (slot == 0) {
                  const int n_grp = ksteps - ks < kp ? ksteps - ks : kp;      // K steps in this barrier round
                  mbar_arrive_expect_tx(&fullb[stage], n_grp * (kATileBytes + b_bytes));
                }
                tma_load_4d(sA + (stage * kp + slot) * kATileBytes, &p.a_maps[sg.map_id], &fullb[stage], c * kBlockK,
                            w0 + sg.dw, h0 + sg.dh, img);
                tma_load_2d(sB + (stage * kp + slot) * kBTileBytes, bmap, &fullb[stage], kcol, b_row0);
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
    }
  } else if (warp == 1 || (warp == 2 && npipe == 2)) {
    // ------------------------------------------------------------------ MMA issuers (one per pipeline)
    const int pipe = warp == 1 ? 0 : 1;
    constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, BLOCK_N, 0, 0);
    // descriptor = constant high word | (smem address >> 4); advancing K by 16 elements (32 B) adds 2 to the low word
    constexpr uint64_t kDescHiB = (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 16) | (1ull << 46) | (2ull << 61);
    const uint32_t a_base = smem_u32(smem + pipe * p.pipe_bytes);
    const uint32_t b_base = a_base + p.b_ring_off;
    uint64_t* fullb = full_bar + pipe * kMaxStages;
    uint64_t* emptyb = empty_bar + pipe * kMaxStages;
    uint64_t* tfullb = tfull_bar + pipe * 2;
    uint64_t* temptyb = tempty_bar + pipe * 2;
    const int kStages = p.stages;
    if (p.halo) {
      uint64_t* afullb = afull_bar + pipe * 4;
      uint64_t* aemptyb = aempty_bar + pipe * 4;
      int stage = 0, a_stage = 0;
      uint32_t phase = 0, a_phase = 0;
      const int c_chunks = p.seg[0].c_chunks;
      // next 8-pixel group = next patch row = halo_w smem rows further
      const uint64_t desc_hi_a = (uint64_t((p.halo_w * 128) >> 4) << 32) | (uint64_t(1) << 16) | (1ull << 46) | (2ull << 61);
      const uint32_t bres_base = smem_u32(smem + p.bres_off);
      if (p.bres) {
        mbar_wait(bres_bar, 0);
        tc_fence_after();
      }
      int li = 0;
      for (int wk = (int)blockIdx.x + pipe * (int)gridDim.x; wk < p.total_tiles; wk += npipe * (int)gridDim.x, ++li) {
        const int as = li & 1;
        mbar_wait(&temptyb[as], ((li >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (pipe * 2 + as) * BLOCK_N;
        for (int c = 0; c < c_chunks; ++c) {
          mbar_wait(&afullb[a_stage], a_phase);
          tc_fence_after();
          const uint32_t a_tile = a_base + a_stage * p.a_stage_bytes;
          if (p.bres) {
            // one barrier round for all taps of this channel chunk
            if (elect_one()) {
              for (int s = 0; s < p.n_seg; ++s) {
                const SegDev sg = p.seg[s];
                const uint32_t a_addr = a_tile + ((sg.dh - p.halo_dh_min) * p.halo_w + (sg.dw - p.halo_dw_min)) * 128;
                const uint64_t ad = desc_hi_a | (a_addr >> 4);
                const uint64_t bd = kDescHiB | ((bres_base + (s * c_chunks + c) * kBTileBytes) >> 4);
                umma_bf16(d_tmem, ad, bd, idesc, (c | s) != 0);
#pragma unroll
                for (int k = 1; k < kBlockK / 16; ++k) umma_bf16(d_tmem, ad + 2 * k, bd + 2 * k, idesc, 1u);
              }
              umma_commit(&aemptyb[a_stage]);
            }
            __syncwarp();
          } else {
            for (int s = 0; s < p.n_seg; ++s) {
              mbar_wait(&fullb[stage], phase);
              tc_fence_after();
              if (elect_one()) {
                const SegDev sg = p.seg[s];
                const uint32_t a_addr = a_tile + ((sg.dh - p.halo_dh_min) * p.halo_w + (sg.dw - p.halo_dw_min)) * 128;
                const uint64_t ad = desc_hi_a | (a_addr >> 4);
                const uint64_t bd = kDescHiB | ((b_base + stage * kBTileBytes) >> 4);
                umma_bf16(d_tmem, ad, bd, idesc, (c | s) != 0);
#pragma unroll
                for (int k = 1; k < kBlockK / 16; ++k) umma_bf16(d_tmem, ad + 2 * k, bd + 2 * k, idesc, 1u);
                umma_commit(&emptyb[stage]);
                if (s == p.n_seg - 1) umma_commit(&aemptyb[a_stage]);   // halo tile free once all taps retire
              }
              __syncwarp();
              if (++stage == kStages) {
                stage = 0;
                phase ^= 1;
              }
            }
          }
          if (++a_stage == p.a_stages) {
            a_stage = 0;
            a_phase ^= 1;
          }
        }
        if (elect_one()) umma_commit(&tfullb[as]);
        __syncwarp();
      }
    } else {
      int stage = 0, rslot = 0;
      uint32_t phase = 0, rphase = 0;
      const uint32_t radd_base = smem_u32(smem + p.radd_off);
      const int kp = p.kpack;
      const int ksteps = p.ksteps;
      int li = 0;
      for (int wk = (int)blockIdx.x + pipe * (int)gridDim.x; wk < p.n_units; wk += npipe * (int)gridDim.x, ++li) {
        const int as = li & 1;
        mbar_wait(&temptyb[as], ((li >> 1) & 1) ^ 1);  // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (pipe * 2 + as) * BLOCK_N;
        // N-split tail unit: the same A tile against BLOCK_N / nsub rows of the weight tile (narrower MMAs)
        constexpr uint32_t idesc_h = umma_idesc_bf16(kBlockM, BLOCK_N >= 64 ? BLOCK_N / 2 : 32, 0, 0);
        constexpr uint32_t idesc_q = umma_idesc_bf16(kBlockM, BLOCK_N >= 128 ? BLOCK_N / 4 : 32, 0, 0);
        const uint32_t idesc_u = wk < p.tail_first ? idesc : (p.tail_nsub == 4 ? idesc_q : idesc_h);
        for (int ks = 0; ks < ksteps; ks += kp) {
          const int n = ksteps - ks < kp ? ksteps - ks : kp;
          mbar_wait(&fullb[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            for (int j = 0; j < n; ++j) {
              const uint64_t ad = kDescHiB | ((a_base + (stage * kp + j) * kATileBytes) >> 4);
              const uint64_t bd = kDescHiB | ((b_base + (stage * kp + j) * kBTileBytes) >> 4);
              umma_bf16(d_tmem, ad, bd, idesc_u, (ks + j) != 0);
#pragma unroll
              for (int k = 1; k < kBlockK / 16; ++k) umma_bf16(d_tmem, ad + 2 * k, bd + 2 * k, idesc_u, 1u);
            }
            umma_commit(&emptyb[stage]);   // one commit per barrier round
          }
          __syncwarp();
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (p.radd) {       // D[:, 64r .. 64r+63] += addend chunk r x I
          constexpr uint32_t idesc64 = umma_idesc_bf16(kBlockM, 64, 0, 0);
          const uint64_t bd = kDescHiB | (smem_u32(smem + p.ident_off) >> 4);
          const int n0 = fdiv_ntile(p, wk) * BLOCK_N;
          const int rn = (p.ncols - n0 < BLOCK_N ? p.ncols - n0 : BLOCK_N) >> 6;
          for (int r = 0; r < rn; ++r) {
            mbar_wait(&afull_bar[rslot], rphase);
            tc_fence_after();
            if (elect_one()) {
              const uint64_t ad = kDescHiB | ((radd_base + rslot * kATileBytes) >> 4);
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k) umma_bf16(d_tmem + r * 64, ad + 2 * k, bd + 2 * k, idesc64, 1u);
              umma_commit(&aempty_bar[rslot]);
            }
            __syncwarp();
            if (++rslot == kRaddStages) {
              rslot = 0;
              rphase ^= 1;
            }
          }
        }
        if (elect_one()) umma_commit(&tfullb[as]);      // accumulator complete -> epilogue
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
      // column sums: byte offset of this lane's channel pair in row k (0..7) of an 8-row group of the swizzled slab
      uint32_t csoff[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) csoff[k] = (uint32_t)(k * 128 + ((((lane >> 2) ^ k) << 4) + ((lane & 3) << 2)));
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
        const int c = pf.n_tile * BLOCK_N + (pf.jb + pf.j) * 64;
        if (p.has_add) tma_load_4d(dst, &p.add_map, &my_bar[slot], c, pf.w0 + slab_w, pf.h0 + slab_h, pf.img);
        if (p.has_mask)
          tma_load_4d(dst + p.has_add * kSlabBytes, &p.mask_map, &my_bar[slot], c, pf.w0 + slab_w, pf.h0 + slab_h, pf.img);
        pf.next(p);
        if (pf.valid) pf.next(p);                  // my chunks are every other one
      };
      if (n_in > 0 && lane == 0) {
        for (int i = 0; i < kEiDepth && pf.valid; ++i) issue_prefetch();
      }
      // bit-packed ReLU mask of the dgrad epilogue: 64 bits per pixel row and chunk, read straight from global memory into
      // two registers one chunk AHEAD (the load has a whole chunk of epilogue work to arrive) -- no shared-memory ring
      const bool has_mbits = kBwdOps && p.mask_bits.ptr != nullptr;
      auto load_mask_bits = [&](const ChunkIter<BLOCK_N>& ci) -> uint2 {
        uint2 r = make_uint2(0u, 0u);
        if (ci.valid && (ci.h0 + rh < p.OH) && (ci.w0 + rw < p.OW) && (ci.img < p.NB)) {
          const char* a = static_cast<const char*>(p.mask_bits.ptr) + (long long)ci.img * p.mask_bits.sn +
                          (long long)(ci.h0 + rh) * p.mask_bits.sh + (long long)(ci.w0 + rw) * p.mask_bits.sw +
                          ((ci.n_tile * BLOCK_N + (ci.jb + ci.j) * 64) >> 3);
          r = __ldg(reinterpret_cast<const uint2*>(a));
        }
        return r;
      };
      ChunkIter<BLOCK_N> nxt = cur;
      uint2 mb_next = make_uint2(0u, 0u);
      if (has_mbits) mb_next = load_mask_bits(cur);
      int n_done = 0;        // my chunks consumed so far
      int in_slot = 0, out_slot = 0;   // ring positions (kept incrementally: no runtime modulo on the critical path)
      uint32_t in_phase = 0;
      int it = 0;            // tiles of this CTA visited
      // every tile of the CTA is visited by BOTH warp sets (each must release the accumulator stage exactly once)
      for (int wk = blockIdx.x; wk < p.n_units; wk += gridDim.x, ++it) {
        const int li = it >> pshift;                         // index of the tile within its pipeline
        const int as = ((it & pshift) << 1) | (li & 1);      // accumulator stage: 2 per pipeline
        const uint32_t aphase = (li >> 1) & 1;
        mbar_wait_relaxed(&tfull_bar[as], aphase);
        tc_fence_after();
        while (cur.valid && cur.wk == wk) {
          const int j = cur.j;
          const int col0 = cur.n_tile * BLOCK_N + (cur.jb + j) * 64;     // (TMEM columns below: j, the unit's own accumulator)
          const uint8_t* in_slab = ei + in_slot * slot_bytes;
          uint8_t* out_slab = eo + out_slot * kSlabBytes;
          // rows outside the image are clipped by the TMA store; they only have to be zeroed for the column sums
          const bool in_img = (cur.h0 + rh < p.OH) && (cur.w0 + rw < p.OW) && (cur.img < p.NB);
          const bool valid = !(kBwdOps && p.colsum != nullptr) || in_img;
          uint2 mb = make_uint2(0u, 0u);
          if (has_mbits) {   // this chunk's mask words were loaded one chunk ago; start the load for my next chunk
            mb = mb_next;
            nxt.next(p);
            if (nxt.valid) nxt.next(p);
            mb_next = load_mask_bits(nxt);
          }
          if (n_done >= kEoDepth) {   // the TMA store that last used this output slab must have drained it
            if (lane == 0) {
              if (kEoDepth >= 3) tma_store_wait_read<2>();
              else if (kEoDepth == 2) tma_store_wait_read<1>();
              else tma_store_wait_read<0>();
            }
            __syncwarp();
          }
          if (n_in > 0) mbar_wait(&my_bar[in_slot], in_phase);
          uint2 obits = make_uint2(0u, 0u);
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            uint32_t acc[32];
            tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + as * BLOCK_N + j * 64 + hf * 32, acc);
            tmem_ld_wait();
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[i]);
            if (kFwdOps && p.shift != nullptr) {
              const float4* sp = reinterpret_cast<const float4*>(s_shift + col0 + hf * 32);    // broadcast LDS.128
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 s4 = sp[i];
                fadd2(v[4 * i + 0], v[4 * i + 1], s4.x, s4.y);
                fadd2(v[4 * i + 2], v[4 * i + 3], s4.z, s4.w);
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
            if (kFwdOps && p.relu) {
#pragma unroll
              for (int i = 0; i < 16; ++i) pk[i] = pack_bf16_relu(v[2 * i], v[2 * i + 1]);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) pk[i] = pack_bf16(v[2 * i], v[2 * i + 1]);
            }
            if (kFwdOps && p.bits_out.ptr != nullptr) {
              // ReLU mask of the stored output, 1 bit per element: channel c (0..31) of this half  <->  bit
              // (7 - (c >> 2)) + 8 (c & 3).  The stored values are >= 0, so "!= 0" is the carry of (half + 0x7FFF) into
              // the half's top bit; PRMT gathers the four flag bytes of a word pair, 2.5 instructions per word.
              uint32_t bw0 = 0u, bw1 = 0u;      // two independent dependency chains (even / odd word pairs)
#pragma unroll
              for (int sp = 6; sp >= 0; sp -= 2) {
                const uint32_t ya = pk[2 * sp] + 0x7FFF7FFFu, yb = pk[2 * sp + 1] + 0x7FFF7FFFu;
                const uint32_t yc = pk[2 * sp + 2] + 0x7FFF7FFFu, yd = pk[2 * sp + 3] + 0x7FFF7FFFu;
                bw0 = (bw0 >> 2) | (prmt(ya, yb, 0x7531u) & 0x80808080u);
                bw1 = (bw1 >> 2) | (prmt(yc, yd, 0x7531u) & 0x80808080u);
              }
              const uint32_t bw = bw0 | (bw1 >> 1);
              if (hf == 0) obits.x = bw; else obits.y = bw;
            }
            if (has_mbits) {
              // shift the four bits of word pair sp to the top of the four bytes; PRMT's sign-replication mode expands
              // them to 0xFFFF / 0 half masks (2.5 instructions per word)
              const uint32_t b = hf == 0 ? mb.x : mb.y;
#pragma unroll
              for (int sp = 0; sp < 8; ++sp) {
                const uint32_t t = b << sp;
                pk[2 * sp] &= prmt(t, 0u, 0x9988u);
                pk[2 * sp + 1] &= prmt(t, 0u, 0xBBAAu);
              }
            }
            if (kBwdOps && p.has_mask) {
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
          if (kFwdOps && p.bits_out.ptr != nullptr && in_img) {
            char* a = static_cast<char*>(p.bits_out.ptr) + (long long)cur.img * p.bits_out.sn +
                      (long long)(cur.h0 + rh) * p.bits_out.sh + (long long)(cur.w0 + rw) * p.bits_out.sw + (col0 >> 3);
            *reinterpret_cast<uint2*>(a) = obits;
          }
          fence_proxy_async();   // generic-proxy smem writes -> visible to the TMA (async proxy)
          __syncwarp();          // all lanes have finished reading the input slot and writing the output slab
          if (lane == 0) {
            tma_store_4d(&p.out_map, out_slab, col0, cur.w0 + slab_w, cur.h0 + slab_h, cur.img);
            tma_store_commit();
            if (n_in > 0 && pf.valid) issue_prefetch();   // refill the input slot just consumed
          }
          if (kBwdOps && p.colsum != nullptr) {
            // lane l sums channel pair l of this chunk over the 32 rows of the bf16 output slab (conflict free:
            // at a fixed row the 32 lanes read the 32 distinct words of one 128-byte line)
            // (row r = 8 g + k sits at g * 1024 + k * 128, its 16-byte unit c16 at (c16 ^ k) << 4: the eight per-k offsets
            // are loop invariant, so the unrolled loop is LDS [reg + immediate] + 2 unpack + 2 FADD per row -- it used to
            // spend 13 instructions per row, more than the rest of the epilogue)
            float s0 = 0.f, s1 = 0.f, t0 = 0.f, t1 = 0.f;
            const uint32_t slab_a = smem_u32(out_slab);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
#pragma unroll
              for (int k = 0; k < 8; k += 2) {
                const uint32_t u = lds32(slab_a + csoff[k] + g * 1024);
                const uint32_t v2 = lds32(slab_a + csoff[k + 1] + g * 1024);
                s0 += bf16_lo(u);
                s1 += bf16_hi(u);
                t0 += bf16_lo(v2);
                t1 += bf16_hi(v2);
              }
            }
            atomicAdd(&s_colacc[col0 + 2 * lane], s0 + t0);
            atomicAdd(&s_colacc[col0 + 2 * lane + 1], s1 + t1);
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
        if (lane == 0) mbar_arrive(&tempty_bar[as]);
      }
      if (lane == 0) tma_store_wait_all();
    } else {
      // ---------------- legacy register epilogue (fp32 output / BLOCK_N == 32)
      int it = 0;
      for (int wk = blockIdx.x; wk < p.total_tiles; wk += gridDim.x, ++it) {
        const int li = it >> pshift;
        const int as = ((it & pshift) << 1) | (li & 1);
        const uint32_t aphase = (li >> 1) & 1;
        int n_tile, img, h0, w0;
        decode_tile(p, wk, n_tile, img, h0, w0);
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
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

}  // namespace urso

// ====================================================================================== host side
struct urso_convgemm {
  urso::ConvGemmParams params;
  int flavor;      // epilogue specialisation of the kernel template (0 general, 1 forward, 2 gradient)
  int block_n;
  int grid;
  int smem_bytes;
};

template <int BLOCK_N, int FLAVOR>
static int launch_conv_gemm_f(const urso_convgemm* h, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    URSO_CUDA_OK(cudaFuncSetAttribute(urso::conv_gemm_kernel<BLOCK_N, FLAVOR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      227 * 1024));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(h->grid);
  cfg.blockDim = dim3(384);
  cfg.dynamicSmemBytes = h->smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = urso::pdl_enabled() ? 1 : 0;
  URSO_CUDA_OK(cudaLaunchKernelEx(&cfg, urso::conv_gemm_kernel<BLOCK_N, FLAVOR>, h->params));
  return 0;
}
template <int BLOCK_N>
static int launch_conv_gemm(const urso_convgemm* h, cudaStream_t stream) {
  if constexpr (BLOCK_N >= 64) {
    if (h->flavor == 1) return launch_conv_gemm_f<BLOCK_N, 1>(h, stream);
    if (h->flavor == 2) return launch_conv_gemm_f<BLOCK_N, 2>(h, stream);
  }
  return launch_conv_gemm_f<BLOCK_N, 0>(h, stream);
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

// Shared-memory plan of the operand pipelines.  Returns false when the requested mode does not fit.
namespace {
struct PipePlan {
  int npipe, stages, kpack, a_stages, bres;
  int pipe_bytes, b_ring_off, bres_bytes;
};

bool plan_stream(int avail, int bn, int ksteps, bool epi_inputs, bool radd, long long tiles_per_cta, PipePlan* pl) {
  const int step_bytes = urso::kATileBytes + bn * urso::kBlockK * 2;
  // preference order: two pipelines (hides the issue-side cost of a barrier round, the bound of the N <= 128 launches),
  // two K steps per round where the ring still gets >= 2 stages per pipeline
  const bool dual_ok = bn <= 128 && tiles_per_cta >= 2;
  const int cand[4][2] = {{2, 2}, {2, 1}, {1, 2}, {1, 1}};
  for (int i = 0; i < 4; ++i) {
    const int np = cand[i][0], kp = cand[i][1];
    if (np == 2 && !dual_ok) continue;
    if (kp == 2 && (ksteps < 4 || bn > 128)) continue;
    if (kp == 2 && np == 2 && bn == 128) continue;     // N = 128: 4 MMAs per round already run at the full rate with 2 issuers
    if (kp == 2 && np == 1 && epi_inputs) continue;    // measured in round 1: loses where the epilogue rings squeeze the ring
    if (kp == 2 && radd) continue;                     // the addend rounds are single K steps
    int stages = avail / (np * kp * step_bytes);
    if (stages > urso::kMaxStages) stages = urso::kMaxStages;
    const int need = (np == 2 || kp == 2) ? 2 : 2;
    if (stages < need) continue;
    if (np == 2 && kp == 1 && stages < 3 && avail / step_bytes >= 3) continue;   // prefer one deeper ring over two shallow ones
    pl->npipe = np; pl->kpack = kp; pl->stages = stages; pl->a_stages = 0; pl->bres = 0; pl->bres_bytes = 0;
    pl->b_ring_off = stages * kp * urso::kATileBytes;
    pl->pipe_bytes = stages * kp * step_bytes;
    return true;
  }
  return false;
}

bool plan_halo(int avail, int bn, int ksteps, int n_tiles_n, int a_stage_bytes, long long tiles_per_cta, PipePlan* pl) {
  const int bt = bn * urso::kBlockK * 2;
  const bool dual_ok = bn <= 128 && tiles_per_cta >= 2;
  for (int np = dual_ok ? 2 : 1; np >= 1; --np) {
    // resident weight operand: one N tile and it fits next to >= 2 halo stages per pipeline
    if (n_tiles_n == 1 && ksteps * bt <= urso::kMaxBresBytes) {
      int a_st = (avail - ksteps * bt) / (np * a_stage_bytes);
      if (a_st > 3) a_st = 3;
      if (a_st >= 2) {
        pl->npipe = np; pl->kpack = 1; pl->stages = 1; pl->a_stages = a_st; pl->bres = 1; pl->bres_bytes = ksteps * bt;
        pl->b_ring_off = a_st * a_stage_bytes;
        pl->pipe_bytes = a_st * a_stage_bytes;
        return true;
      }
    }
    // streamed weight tiles: 2 halo stages + >= 3 B tiles per pipeline
    const int per_pipe = avail / np;
    int b_st = (per_pipe - 2 * a_stage_bytes) / bt;
    if (b_st > urso::kMaxStages) b_st = urso::kMaxStages;
    if (b_st >= 3) {
      int a_st = (per_pipe - b_st * bt) / a_stage_bytes;
      if (a_st > 3) a_st = 3;
      pl->npipe = np; pl->kpack = 1; pl->stages = b_st; pl->a_stages = a_st; pl->bres = 0; pl->bres_bytes = 0;
      pl->b_ring_off = a_st * a_stage_bytes;
      pl->pipe_bytes = a_st * a_stage_bytes + b_st * bt;
      return true;
    }
  }
  return false;
}
}  // namespace

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
  int ktot = 0, ksteps = 0;
  for (int s = 0; s < d->n_seg; ++s) {
    const urso_seg& sg = d->seg[s];
    if (sg.map_id < 0 || sg.map_id >= d->n_a || sg.c_chunks < 1 || sg.c_chunks * 64 > ((d->a[sg.map_id].C + 63) / 64) * 64) {
      set_error("segment %d invalid (map %d, chunks %d)", s, sg.map_id, sg.c_chunks);
      delete h;
      return 2;
    }
    p.seg[s] = SegDev{(int16_t)sg.map_id, (int16_t)sg.dh, (int16_t)sg.dw, (int16_t)sg.c_chunks};
    ktot += sg.c_chunks * 64;
    ksteps += sg.c_chunks;
  }
  if (ktot != d->b_k) {
    set_error("segments cover K=%d but b_k=%d", ktot, d->b_k);
    delete h;
    return 2;
  }
  p.ksteps = ksteps;
  for (int i = 0; i < URSO_MAX_AMAPS; ++i) {
    const urso_view4& v = d->a[i < d->n_a ? i : 0];
    if (int rc = make_view_map(&p.a_maps[i], v, d->TW, d->TH)) {
      delete h;
      return rc;
    }
  }
  int bn = d->block_n;
  if (bn == 0) bn = d->b_rows >= 256 && d->b_rows % 256 == 0 ? 256 : (d->b_rows >= 128 ? 128 : (d->b_rows >= 64 ? 64 : 32));
  if (bn != 32 && bn != 64 && bn != 128 && bn != 256) {
    set_error("block_n=%d unsupported", bn);
    delete h;
    return 2;
  }
  // epilogue flavour and its shared memory
  const int kColAccOnly = d->colsum != nullptr ? ((d->b_rows * 4 + 1023) / 1024) * 1024 : 0;
  p.colacc_bytes = kColAccOnly;
  p.has_mask = d->mask.ptr != nullptr;
  p.epi_tma = (!d->out_fp32 && d->b_rows % 64 == 0 && bn >= 64) ? 1 : 0;
  // residual / fan-in addend through the tensor core (see ConvGemmParams::radd): stream mode with the TMA epilogue, one
  // pipeline (BLOCK_N = 256), short K loops -- the store-bound launches, whose epilogue warps are the bottleneck.  (With
  // >= 8 K steps the third operand stage that the addend ring displaces is worth more: res5x_2c 54 -> 67 us measured.)
  p.radd = (residual_mma_enabled() && d->addend.ptr != nullptr && p.epi_tma && !d->halo && bn == 256 && ksteps <= 4) ? 1 : 0;
  p.shift_bytes = (p.epi_tma && d->shift != nullptr) ? ((d->b_rows * 4 + 1023) / 1024) * 1024 : 0;
  const int kColAcc = kColAccOnly + p.shift_bytes;     // everything between the control block and the epilogue rings
  if ((d->relu_bits.ptr != nullptr || d->mask_bits.ptr != nullptr) && !p.epi_tma) {
    set_error("bit-packed masks need the bf16 TMA epilogue (bf16 output, channels a multiple of 64)");
    delete h;
    return 2;
  }
  if (d->mask_bits.ptr != nullptr && d->mask.ptr != nullptr) {
    set_error("mask and mask_bits are mutually exclusive");
    delete h;
    return 2;
  }
  int n_in = 0, radd_bytes = 0, epi_bytes = 0;
  for (const int bn_asked = bn;; p.radd = 0, bn = bn_asked) {       // second pass: the addend ring did not fit
    p.has_add = d->addend.ptr != nullptr && !p.radd;
    n_in = p.has_add + p.has_mask;
    radd_bytes = p.radd ? 8192 + kRaddStages * kATileBytes : 0;     // identity tile + addend ring
    if (p.epi_tma) {
      // per-warp private rings (8 epilogue warps): input slots (prefetch depth) and output slabs (stores in flight).
      // Launches with a long K loop visit the epilogue rarely: minimal rings, smem goes to the operand pipelines.  Launches
      // with a short K loop (<= 6 K steps: the MMAs of a tile take less time than its epilogue) are epilogue / store bound:
      // a depth-1 input ring would expose the full TMA latency of every chunk (measured per layer in round 1).
      const bool heavy = ksteps >= 4;
      const bool epi_bound = ksteps <= 6;
      p.ei_depth = (heavy || n_in == 2) ? 1 : 2;
      p.eo_depth = (heavy || n_in == 2 || p.radd) ? 1 : 2;
      if (n_in > 0 && epi_bound) p.ei_depth = 2;
      epi_bytes = 8 * p.ei_depth * n_in * kSlabBytes + 8 * p.eo_depth * kSlabBytes;
      const int room = kSmemBudget - kCtrlBytes - kColAcc - radd_bytes;
      const int bn_min = bn == 256 ? 128 : bn;   // the narrowest tile this launch may fall back to must still get 2 stages
      if (p.ei_depth == 2 && room - epi_bytes < 2 * (kATileBytes + bn_min * kBlockK * 2)) {
        p.ei_depth = 1;
        epi_bytes = 8 * p.ei_depth * n_in * kSlabBytes + 8 * p.eo_depth * kSlabBytes;
      }
      const int min_stages = ((n_in > 0 && epi_bound) || p.radd) ? 2 : 3;
      if (bn == 256 && room - epi_bytes < min_stages * (kATileBytes + 256 * kBlockK * 2)) bn = 128;
    } else {
      epi_bytes = kLegacyScratchBytes;
    }
    if (!p.radd || bn == 256) break;
  }
  h->block_n = bn;
  p.n_tiles_n = (d->b_rows + bn - 1) / bn;
  p.tiles_w = (d->OW + d->TW - 1) / d->TW;
  p.tiles_h = (d->OH + d->TH - 1) / d->TH;
  const long long total = (long long)p.tiles_w * p.tiles_h * d->NB * p.n_tiles_n;
  if (total <= 0 || total > 0x7fffffffLL) {
    set_error("bad tile count %lld", total);
    delete h;
    return 2;
  }
  p.total_tiles = (int)total;
  int sms = max_ctas();
  if (sms <= 0) sms = 148;
  h->grid = p.total_tiles < sms ? p.total_tiles : sms;
  const long long tiles_per_cta = (total + h->grid - 1) / h->grid;
  PipePlan pl;
  bool planned = false;
  int halo_w = 0, halo_h = 0, dw_min = 0, dh_min = 0;
  if (d->halo) {
    // validate the halo geometry: one stride-1 view, 8 x 16 patch, equal chunk counts, small tap offsets
    int dw_max = -(1 << 20), dh_max = -(1 << 20);
    dw_min = dh_min = 1 << 20;
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
    halo_w = 8 + dw_max - dw_min;
    halo_h = 16 + dh_max - dh_min;
  }
  // Plan the operand pipelines in what the epilogue leaves; if nothing fits, shrink the epilogue rings to depth 1 and retry.
  // (All regions are multiples of 1 KB, so the epilogue rings need no alignment slack.)
  for (int attempt = 0; attempt < 2 && !planned; ++attempt) {
    if (attempt == 1) {
      if (!p.epi_tma || (p.ei_depth == 1 && p.eo_depth == 1)) break;
      p.ei_depth = p.eo_depth = 1;
      epi_bytes = 8 * n_in * kSlabBytes + 8 * kSlabBytes;
    }
    const int avail = kSmemBudget - kCtrlBytes - kColAcc - epi_bytes - radd_bytes;
    if (d->halo) {
      const int a_stage_bytes = (halo_w * halo_h * 128 + 1023) / 1024 * 1024;
      if (plan_halo(avail, bn, ksteps, p.n_tiles_n, a_stage_bytes, tiles_per_cta, &pl)) {
        planned = true;
        p.halo = 1;
        p.halo_w = halo_w;
        p.halo_dw_min = dw_min;
        p.halo_dh_min = dh_min;
        p.halo_bytes = halo_w * halo_h * 128;
        p.a_stage_bytes = a_stage_bytes;
        continue;
      }   // else: not enough shared memory for the halo rings -> plain stream mode on the same patch
    }
    planned = plan_stream(avail, bn, ksteps, n_in > 0, p.radd != 0, tiles_per_cta, &pl);
  }
  if (!planned) {
    set_error("not enough shared memory for 2 pipeline stages (BLOCK_N=%d)", bn);
    delete h;
    return 2;
  }
  if (p.halo) {
    if (int rc = make_view_map(&p.a_halo_map, d->a[0], halo_w, halo_h)) {
      delete h;
      return rc;
    }
  }
  p.npipe = pl.npipe;
  p.stages = pl.stages;
  p.kpack = pl.kpack;
  p.a_stages = pl.a_stages;
  p.bres = pl.bres;
  p.pipe_bytes = pl.pipe_bytes;
  p.b_ring_off = pl.b_ring_off;
  p.bres_off = pl.npipe * pl.pipe_bytes;
  p.ident_off = p.bres_off + pl.bres_bytes;
  p.radd_off = p.ident_off + (p.radd ? 8192 : 0);
  p.ctrl_off = p.radd_off + (p.radd ? kRaddStages * kATileBytes : 0);
  const int fixed = p.ctrl_off + kCtrlBytes + kColAcc;
  if (p.epi_tma) {
    p.ei_off = (fixed + 1023) / 1024 * 1024;
    p.eo_off = p.ei_off + 8 * p.ei_depth * n_in * kSlabBytes;
    h->smem_bytes = p.eo_off + 8 * p.eo_depth * kSlabBytes;
  } else {
    h->smem_bytes = fixed + kLegacyScratchBytes;
  }
  if (h->smem_bytes > kSmemBudget) {
    set_error("internal: shared memory plan %d bytes exceeds 227 KB", h->smem_bytes);
    delete h;
    return 2;
  }
  if (int rc = make_mat_map(&p.b_map, d->b, d->b_rows, d->b_k, bn)) {
    delete h;
    return rc;
  }
  // ---- N-split tail (see ConvGemmParams::tail_*): K-heavy BLOCK_N = 256 launches in stream mode with a partial last wave
  p.tail_first = p.n_units = p.total_tiles;
  p.tail_nsub = 1;
  if (tail_split_enabled() && bn == 256 && !p.halo && !p.radd && p.epi_tma && pl.npipe == 1 && pl.kpack == 1 && ksteps >= 8 &&
      d->b_rows % 256 == 0 && p.total_tiles > h->grid) {
    const int full = p.total_tiles / h->grid * h->grid;
    const int R = p.total_tiles - full;
    const int nsub = R == 0 ? 1 : (4 * R <= h->grid ? 4 : (2 * R <= h->grid ? 2 : 1));
    if (nsub > 1) {
      p.tail_first = full;
      p.tail_nsub = nsub;
      p.n_units = full + R * nsub;
      if (int rc = make_mat_map(&p.b_sub_map, d->b, d->b_rows, d->b_k, bn / nsub)) {
        delete h;
        return rc;
      }
    }
  }
  if (p.epi_tma) {
    const int bw = d->TW < 32 ? d->TW : 32, bh = 32 / bw;
    int rc = make_pix_map(&p.out_map, d->out, d->b_rows, d->OW, d->OH, d->NB, bw, bh);
    if (!rc) rc = make_pix_map(&p.add_map, p.has_add ? d->addend : d->out, d->b_rows, d->OW, d->OH, d->NB, bw, bh);
    if (!rc) rc = make_pix_map(&p.mask_map, p.has_mask ? d->mask : d->out, d->b_rows, d->OW, d->OH, d->NB, bw, bh);
    if (!rc && p.radd) rc = make_pix_map(&p.radd_map, d->addend, d->b_rows, d->OW, d->OH, d->NB, d->TW, d->TH);
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
  p.fd_ntn = make_fastdiv((uint32_t)p.n_tiles_n);
  p.fd_tw = make_fastdiv((uint32_t)p.tiles_w);
  p.fd_th = make_fastdiv((uint32_t)p.tiles_h);
  p.ncols = d->b_rows;
  p.out = PixDev{d->out.ptr, d->out.sn, d->out.sh, d->out.sw};
  p.addend = PixDev{d->addend.ptr, d->addend.sn, d->addend.sh, d->addend.sw};
  p.mask = PixDev{d->mask.ptr, d->mask.sn, d->mask.sh, d->mask.sw};
  p.bits_out = PixDev{d->relu_bits.ptr, d->relu_bits.sn, d->relu_bits.sh, d->relu_bits.sw};
  p.mask_bits = PixDev{d->mask_bits.ptr, d->mask_bits.sn, d->mask_bits.sh, d->mask_bits.sw};
  p.out_fp32 = d->out_fp32;
  p.relu = d->relu;
  p.shift = d->shift;
  p.colsum = d->colsum;
  const bool bwd_ops = d->mask.ptr != nullptr || d->mask_bits.ptr != nullptr || d->colsum != nullptr;
  const bool fwd_ops = d->shift != nullptr || d->relu != 0 || d->relu_bits.ptr != nullptr;
  h->flavor = !p.epi_tma ? 0 : (!bwd_ops ? 1 : (!fwd_ops ? 2 : 0));
  *out = h;
  return 0;
}

extern "C" int urso_convgemm_launch(urso_convgemm_t* h, void* stream) {
  URSO_REQUIRE(h != nullptr, "null handle");
  URSO_REQUIRE(!urso::dry_run(), "dry run: nothing can be launched");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (h->block_n) {
    case 32: return launch_conv_gemm<32>(h, s);
    case 64: return launch_conv_gemm<64>(h, s);
    case 128: return launch_conv_gemm<128>(h, s);
    case 256: return launch_conv_gemm<256>(h, s);
  }
  urso::set_error("unsupported BLOCK_N %d", h->block_n);
  return 2;
}

extern "C" void urso_convgemm_destroy(urso_convgemm_t* h) { delete h; }

/* plan introspection (tests / profiling): writes {block_n, npipe, stages, kpack, halo, bres, a_stages, smem_bytes, grid} */
extern "C" int urso_convgemm_plan_info(const urso_convgemm_t* h, int32_t* out9) {
  URSO_REQUIRE(h != nullptr && out9 != nullptr, "null argument");
  const urso::ConvGemmParams& p = h->params;
  const int32_t v[9] = {h->block_n, p.npipe, p.stages, p.kpack, p.halo, p.bres, p.a_stages, h->smem_bytes, h->grid};
  for (int i = 0; i < 9; ++i) out9[i] = v[i];
  return 0;
}

/* N-split tail of the launch: number of sub-tiles each tile of the last partial wave is cut into (1 = none) */
extern "C" int urso_convgemm_tail_split(const urso_convgemm_t* h) { return h != nullptr ? h->params.tail_nsub : 0; }
