// Engine W: weight-gradient GEMM on tcgen05 tensor cores.
//
//   G[t][p, q] += sum over pixels  P_t[pixel + shift_t, p] * Q[pixel, q]
//
// The reduction (MMA K) dimension is the PIXEL axis, so both operands are "MN-major": a TMA box of
// {64 channels, TW, TH, 1} lands in shared memory as 64 pixel rows of 128 bytes -- exactly the canonical
// SWIZZLE_128B MN-major UMMA atom (8 pixel rows x 64 channels per 1 KB group, SBO = 1 KB between groups,
// LBO = distance between 64-channel atoms).  No transposition of activations or gradients is ever materialised.
//
// One CTA owns a 128-channel P tile x BLOCK_Q-channel Q tile for up to T filter taps (T*BLOCK_Q <= 512 TMEM columns,
// all taps reuse the single Q tile per K step) over a contiguous range of pixel blocks (split-K); results are
// reduced into the fp32 gradient with atomics.
// HALO MODE (several filter taps on one stride-1 view, 8 x 8 pixel blocks): the tap-shifted P operands of a CTA's taps are
// overlapping windows of ONE box per 64-channel atom (block + halo, e.g. 10 x 10 pixels for a 3x3 filter) instead of one
// 8 KB atom per tap: the MN-major descriptor of tap (dh, dw) starts (dh * halo_w + dw) pixel rows into the box, its
// 8-pixel groups are halo_w rows apart (SBO) -- UMMA descriptors swizzle on absolute shared-memory address bits, so a
// start shifted by whole 128-byte rows reads what TMA wrote (the trick of Engine F's halo mode).  The 64-channel 3x3
// layers go from 88 KB to 21 KB per K step (94 -> 22 B/clk of operand supply, against ~67 B/clk the SM can ingest).
// Roles (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warps 4-7 = epilogue.
#include "common.cuh"
#include "ptx.cuh"

namespace urso {

struct WSegDev {
  int16_t map_id, dh, dw, pad;
};

struct WgradParams {
  CUtensorMap p_maps[URSO_MAX_AMAPS];
  CUtensorMap q_map;
  CUtensorMap p_halo_map;   // halo mode: box {64, halo_w, halo_h, 1} of view 0
  WSegDev seg[URSO_MAX_SEGS];
  int n_seg, taps_per_cta, n_seg_groups;
  int order;       // work-item decode order (see kernel)
  int interleave;  // split s takes pixel blocks s, s + split_k, ... instead of one contiguous range: all CTAs then walk the
                   // tensor front to back together, like Engine F's round-robin tiles (an L2-sharing pair needs that)
  int pair_mode;   // PC <= 64: the two 64-row halves of the MMA M dimension carry two different filter taps
  int halo, halo_w, halo_box_bytes, halo_tx_bytes;   // box pitch in the stage (1 KB multiple) / bytes one box load delivers
  int PC, QC;
  int p_tiles, q_tiles;
  int tiles_w, tiles_h, TW, TH;
  int n_pix_blocks, split_k;
  int stages, stage_bytes;
  int tmem_cols;
  int bulk_red;    // epilogue: accumulator rows go through shared memory and one cp.reduce.async.bulk (fp32 add) per row
                   // instead of 16-byte REDG instructions (measured 1.29 cycles per LANE on the SM side: 11 us per launch)
  int red_bufs;    // row buffers in the (idle) operand ring: 1 or 2
  float* g;
  long long g_seg_stride, g_sp, g_sq;
};

constexpr int kAtomBytes = 64 * 64 * 2;  // 64 pixels x 64 channels bf16 = 8 KB

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int BLOCK_Q>
__global__ void __launch_bounds__(256, 1) wgrad_kernel(const __grid_constant__ WgradParams p) {
  constexpr int QA = BLOCK_Q / 64;  // 64-channel atoms in the Q tile
  extern __shared__ __align__(1024) uint8_t smem[];   // SWIZZLE_128B atoms need 1 KB alignment
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * p.stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tfull_bar = empty_bar + p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // work item decode: blockIdx.x -> (split, q_tile, p_tile, seg_group); split fastest so that CTAs that run
  // concurrently stream different pixels of the same operand columns.
  int wi = blockIdx.x;
  int split, q_tile, p_tile, seg_group;
  if (p.order == 0) {   // split fastest
    split = wi % p.split_k;  wi /= p.split_k;
    q_tile = wi % p.q_tiles; wi /= p.q_tiles;
    p_tile = wi % p.p_tiles; wi /= p.p_tiles;
    seg_group = wi;
  } else {              // output tiles fastest: co-resident CTAs stream the SAME pixels (operand tiles shared through L2)
    q_tile = wi % p.q_tiles; wi /= p.q_tiles;
    p_tile = wi % p.p_tiles; wi /= p.p_tiles;
    seg_group = wi % p.n_seg_groups; wi /= p.n_seg_groups;
    split = wi;
  }
  // "units" = taps (normal) or tap pairs (pair mode)
  const int n_units = p.pair_mode ? (p.n_seg + 1) / 2 : p.n_seg;
  const int seg0 = seg_group * p.taps_per_cta;
  const int T = min(p.taps_per_cta, n_units - seg0);
  // halo mode: origin of this CTA's box = the smallest tap offsets of its group
  int hdh = 0, hdw = 0;
  if (p.halo) {
    const int t_first = p.pair_mode ? 2 * seg0 : seg0;
    const int t_last = min(p.pair_mode ? 2 * (seg0 + T) : seg0 + T, p.n_seg);
    hdh = hdw = 1 << 20;
    for (int t = t_first; t < t_last; ++t) {
      hdh = min(hdh, (int)p.seg[t].dh);
      hdw = min(hdw, (int)p.seg[t].dw);
    }
  }
  // my pixel blocks: kb = kb_begin + i * kb_step, i < kb_count
  const int kb_step = p.interleave ? p.split_k : 1;
  const int kb_begin = p.interleave ? split : (int)((long long)p.n_pix_blocks * split / p.split_k);
  const int kb_count = p.interleave ? (p.n_pix_blocks - split + p.split_k - 1) / p.split_k
                                    : (int)((long long)p.n_pix_blocks * (split + 1) / p.split_k) - kb_begin;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < URSO_MAX_AMAPS; ++i) tma_prefetch_desc(&p.p_maps[i]);
    tma_prefetch_desc(&p.q_map);
    if (p.halo) tma_prefetch_desc(&p.p_halo_map);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();   // the swizzled layouts assume it
  if (warp == 2) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                 // programmatic dependent launch: the prologue above overlapped the previous kernel's tail
  pdl_launch_dependents();

  // Producer / MMA issuer: warp-uniform control flow in all lanes, one elected lane issues (keeps the loop state in
  // uniform registers; a loop inside `if (lane == 0)` costs >100 SASS instructions per K step in R2UR shuffling).
  // Two producer warps (0 and 3) take alternate K steps: each K step is (QA + 2T) separate 8 KB TMA boxes.
  if (warp == 0 || warp == 3) {
    const int par = warp == 0 ? 0 : 1;
    int stage = 0;
    uint32_t phase = 0;
    const int n_pa = p.pair_mode ? 1 : 2;      // 64-channel atoms of the P tile
    const uint32_t bytes = p.halo ? QA * kAtomBytes + n_pa * p.halo_tx_bytes : (QA + 2 * T) * kAtomBytes;
    for (int i = 0; i < kb_count; ++i) {
      const int kb = kb_begin + i * kb_step;
      if ((i & 1) == par) {
        const int twi = kb % p.tiles_w;
        const int rest = kb / p.tiles_w;
        const int thi = rest % p.tiles_h;
        const int img = rest / p.tiles_h;
        const int h0 = thi * p.TH, w0 = twi * p.TW;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&full_bar[stage], bytes);
          uint8_t* base = smem + stage * p.stage_bytes;
#pragma unroll 1
          for (int a = 0; a < QA; ++a)
            tma_load_4d(base + a * kAtomBytes, &p.q_map, &full_bar[stage], q_tile * BLOCK_Q + a * 64, w0, h0, img);
          uint8_t* pbase = base + QA * kAtomBytes;
          if (p.halo) {
            for (int a = 0; a < n_pa; ++a)
              tma_load_4d(pbase + a * p.halo_box_bytes, &p.p_halo_map, &full_bar[stage],
                          (p.pair_mode ? 0 : p_tile * 128) + a * 64, w0 + hdw, h0 + hdh, img);
          }
#pragma unroll 1
          for (int t = 0; t < (p.halo ? 0 : T); ++t) {
            if (p.pair_mode) {
              const int ta = 2 * (seg0 + t), tb = min(ta + 1, p.n_seg - 1);   // odd tap count: the last atom repeats a tap
              const WSegDev sa = p.seg[ta], sb = p.seg[tb];
              tma_load_4d(pbase + (2 * t) * kAtomBytes, &p.p_maps[sa.map_id], &full_bar[stage], 0, w0 + sa.dw, h0 + sa.dh,
                          img);
              tma_load_4d(pbase + (2 * t + 1) * kAtomBytes, &p.p_maps[sb.map_id], &full_bar[stage], 0, w0 + sb.dw,
                          h0 + sb.dh, img);
            } else {
              const WSegDev sg = p.seg[seg0 + t];
              tma_load_4d(pbase + (2 * t) * kAtomBytes, &p.p_maps[sg.map_id], &full_bar[stage], p_tile * 128, w0 + sg.dw,
                          h0 + sg.dh, img);
              tma_load_4d(pbase + (2 * t + 1) * kAtomBytes, &p.p_maps[sg.map_id], &full_bar[stage], p_tile * 128 + 64,
                          w0 + sg.dw, h0 + sg.dh, img);
            }
          }
        }
        __syncwarp();
      }
      if (++stage == p.stages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, BLOCK_Q, 1, 1);  // both operands MN-major
    // MN-major SW128 descriptor: LBO = distance between 64-channel atoms (8 KB), SBO = 1 KB between 8-pixel groups
    constexpr uint64_t kDescHi = (uint64_t(1024 >> 4) << 32) | (uint64_t(kAtomBytes >> 4) << 16) | (1ull << 46) | (2ull << 61);
    const uint32_t s_base = smem_u32(smem);
    // halo mode: per unit the window's byte offset into the box and the LBO of its descriptor (the second 64-row atom of M
    // is the next 64 channels = the next box, or -- pair mode -- the unit's second tap = another window of the same box)
    uint32_t h_off[8], h_lbo[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      h_off[t] = h_lbo[t] = 0;
      if (p.halo && t < T) {
        if (p.pair_mode) {
          const int ta = 2 * (seg0 + t), tb = min(ta + 1, p.n_seg - 1);
          const uint32_t oa = (uint32_t)(((int)p.seg[ta].dh - hdh) * p.halo_w + ((int)p.seg[ta].dw - hdw)) * 128u;
          const uint32_t ob = (uint32_t)(((int)p.seg[tb].dh - hdh) * p.halo_w + ((int)p.seg[tb].dw - hdw)) * 128u;
          h_off[t] = oa;
          h_lbo[t] = ob - oa;      // taps are listed row-major: ob >= oa
        } else {
          const WSegDev sg = p.seg[seg0 + t];
          h_off[t] = (uint32_t)(((int)sg.dh - hdh) * p.halo_w + ((int)sg.dw - hdw)) * 128u;
          h_lbo[t] = (uint32_t)p.halo_box_bytes;
        }
      }
    }
    const uint32_t h_kadv = (uint32_t)p.halo_w * 16u;     // 16 pixels = 2 block rows = 2 * halo_w box rows of 128 B, in 16-byte units
    const uint64_t h_sbo = uint64_t((uint32_t)p.halo_w * 128u >> 4) << 32;
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < kb_count; ++i) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t q_addr = s_base + stage * p.stage_bytes;
        const uint64_t bd = kDescHi | (q_addr >> 4);
        const uint32_t first = i > 0;
        if (p.halo) {
          const uint32_t pbox = q_addr + QA * kAtomBytes;
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            if (t < T) {
              const uint64_t ad = h_sbo | (uint64_t(h_lbo[t] >> 4) << 16) | (1ull << 46) | (2ull << 61) | ((pbox + h_off[t]) >> 4);
              const uint32_t d = tmem_base + t * BLOCK_Q;
              umma_bf16(d, ad, bd, idesc, first);
              umma_bf16(d, ad + h_kadv, bd + 128, idesc, 1u);
              umma_bf16(d, ad + 2 * h_kadv, bd + 256, idesc, 1u);
              umma_bf16(d, ad + 3 * h_kadv, bd + 384, idesc, 1u);
            }
          }
        }
#pragma unroll 1
        for (int t = 0; t < (p.halo ? 0 : T); ++t) {
          const uint64_t ad = kDescHi | ((q_addr + (QA + 2 * t) * kAtomBytes) >> 4);
          const uint32_t d = tmem_base + t * BLOCK_Q;
          // 64 pixels = 4 x UMMA_K(16); 16 pixels = two 8-pixel groups = 2 KB -> +128 in descriptor units
          umma_bf16(d, ad, bd, idesc, first);
          umma_bf16(d, ad + 128, bd + 128, idesc, 1u);
          umma_bf16(d, ad + 256, bd + 256, idesc, 1u);
          umma_bf16(d, ad + 384, bd + 384, idesc, 1u);
        }
        umma_commit(&empty_bar[stage]);
      }
      __syncwarp();
      if (++stage == p.stages) {
        stage = 0;
        phase ^= 1;
      }
    }
    if (elect_one()) umma_commit(tfull_bar);
    __syncwarp();
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    if (kb_count > 0) {
      mbar_wait(tfull_bar, 0);
      tc_fence_after();
      // bulk-reduce path: the operand ring is idle now (every MMA that read it has completed); lane = accumulator row
      // owns one padded row buffer (pitch = row bytes + 16: conflict-free 16-byte stores of 8 consecutive lanes)
      const int ncols = min(BLOCK_Q, p.QC - q_tile * BLOCK_Q);
      constexpr uint32_t kPitch = BLOCK_Q * 4 + 16;
      for (int t = 0; t < T; ++t) {
        // accumulator row -> (filter tap, P channel)
        const int tap = p.pair_mode ? 2 * (seg0 + t) + (row >> 6) : seg0 + t;
        const int pch = p.pair_mode ? (row & 63) : p_tile * 128 + row;
        const bool row_ok = pch < p.PC && tap < p.n_seg;
        float* grow = p.g + (long long)tap * p.g_seg_stride + (long long)pch * p.g_sp;
        if (p.bulk_red) {
          const uint32_t rbuf = smem_u32(smem) + (uint32_t)((t % p.red_bufs) * 128 + row) * kPitch;
          if (t >= p.red_bufs) tma_store_wait_read<0>();   // my own earlier reduction has finished reading this row buffer
#pragma unroll 1
          for (int j = 0; j < BLOCK_Q / 32; ++j) {
            if (j * 32 >= ncols) break;
            uint32_t acc[32];
            tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + t * BLOCK_Q + j * 32, acc);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i)
              sts128(rbuf + j * 128 + i * 16, acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
          }
          fence_proxy_async();
          if (row_ok) bulk_reduce_add_f32(grow + q_tile * BLOCK_Q, rbuf, (uint32_t)ncols * 4u);
          tma_store_commit();
          continue;
        }
        const bool vec_ok = (reinterpret_cast<uintptr_t>(grow) & 15) == 0;
#pragma unroll 1
        for (int j = 0; j < BLOCK_Q / 32; ++j) {
          const int col0 = q_tile * BLOCK_Q + j * 32;
          if (col0 >= p.QC) break;
          uint32_t acc[32];
          tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + t * BLOCK_Q + j * 32, acc);
          tmem_ld_wait();
          if (row_ok) {
            if (p.g_sq == 1 && col0 + 32 <= p.QC && vec_ok) {
              // contiguous along q: 16-byte vector reductions (REDG.E.ADD.F32x4), 4x fewer L2 atomic sectors
#pragma unroll
              for (int i = 0; i < 8; ++i)
                red_add_v4(grow + col0 + 4 * i, __uint_as_float(acc[4 * i]), __uint_as_float(acc[4 * i + 1]),
                           __uint_as_float(acc[4 * i + 2]), __uint_as_float(acc[4 * i + 3]));
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                if (col0 + i < p.QC) atomicAdd(grow + (long long)(col0 + i) * p.g_sq, __uint_as_float(acc[i]));
              }
            }
          }
        }
      }
      if (p.bulk_red) tma_store_wait_all();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

}  // namespace urso

struct urso_wgrad {
  urso::WgradParams params;
  int block_q, grid, smem_bytes;
};

template <int BLOCK_Q>
static int launch_wgrad(const urso_wgrad* h, cudaStream_t stream) {
  static int attr_smem = 0;
  if (attr_smem < h->smem_bytes) {
    URSO_CUDA_OK(cudaFuncSetAttribute(urso::wgrad_kernel<BLOCK_Q>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      220 * 1024));
    attr_smem = 220 * 1024;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(h->grid);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = h->smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = urso::pdl_enabled() ? 1 : 0;
  URSO_CUDA_OK(cudaLaunchKernelEx(&cfg, urso::wgrad_kernel<BLOCK_Q>, h->params));
  return 0;
}

extern "C" int urso_wgrad_create(const urso_wgrad_desc* d, urso_wgrad_t** out) {
  using namespace urso;
  URSO_REQUIRE(d != nullptr && out != nullptr, "null argument");
  URSO_REQUIRE(d->n_p >= 1 && d->n_p <= URSO_MAX_AMAPS, "n_p=%d out of range", d->n_p);
  URSO_REQUIRE(d->n_seg >= 1 && d->n_seg <= URSO_MAX_SEGS, "n_seg=%d out of range", d->n_seg);
  URSO_REQUIRE(d->TW * d->TH == 64, "TW*TH must be 64 (got %dx%d)", d->TW, d->TH);
  URSO_REQUIRE(d->g != nullptr, "null gradient output");
  auto* h = new urso_wgrad();
  WgradParams& p = h->params;
  memset(&p, 0, sizeof(p));
  int bq = d->block_q;
  if (bq == 0) bq = d->QC >= 256 ? 256 : (d->QC > 64 ? 128 : 64);
  if (bq != 64 && bq != 128 && bq != 256) {
    set_error("block_q=%d unsupported", bq);
    delete h;
    return 2;
  }
  h->block_q = bq;
  for (int i = 0; i < URSO_MAX_AMAPS; ++i) {
    if (int rc = make_view_map(&p.p_maps[i], d->p[i < d->n_p ? i : 0], d->TW, d->TH)) {
      delete h;
      return rc;
    }
  }
  if (int rc = make_view_map(&p.q_map, d->q, d->TW, d->TH)) {
    delete h;
    return rc;
  }
  for (int s = 0; s < d->n_seg; ++s) {
    if (d->seg[s].map_id < 0 || d->seg[s].map_id >= d->n_p) {
      set_error("segment %d: bad map id %d", s, d->seg[s].map_id);
      delete h;
      return 2;
    }
    p.seg[s] = WSegDev{(int16_t)d->seg[s].map_id, (int16_t)d->seg[s].dh, (int16_t)d->seg[s].dw, 0};
  }
  p.n_seg = d->n_seg;
  p.pair_mode = (d->PC <= 64 && d->n_seg >= 2) ? 1 : 0;
  const int n_units = p.pair_mode ? (d->n_seg + 1) / 2 : d->n_seg;
  int tmax = 512 / bq;
  const int qa = bq / 64;
  // ---- halo mode: every tap on one stride-1 view, 8 x 8 pixel blocks.  The units of a CTA are then limited by tensor
  // memory only (the stage holds one box per P atom, not two atoms per unit); taken when the boxes are < 70 % of the bytes.
  p.halo = 0;
  bool one_view = d->n_seg >= 2 && d->TW == 8 && d->TH == 8 && wgrad_halo_enabled();
  for (int s2 = 1; s2 < d->n_seg && one_view; ++s2) one_view = d->seg[s2].map_id == d->seg[0].map_id;
  if (one_view) {
    const int t_units = tmax > 8 ? 8 : tmax;          // the kernel unrolls at most 8 units
    const int groups = (n_units + t_units - 1) / t_units;
    const int per = (n_units + groups - 1) / groups;
    int span_w = 0, span_h = 0;
    for (int g0 = 0; g0 < groups; ++g0) {
      const int first = (p.pair_mode ? 2 : 1) * g0 * per;
      int last = (p.pair_mode ? 2 : 1) * (g0 + 1) * per;
      if (last > d->n_seg) last = d->n_seg;
      int lo_w = 1 << 20, hi_w = -(1 << 20), lo_h = 1 << 20, hi_h = -(1 << 20);
      for (int t = first; t < last; ++t) {
        lo_w = d->seg[t].dw < lo_w ? d->seg[t].dw : lo_w; hi_w = d->seg[t].dw > hi_w ? d->seg[t].dw : hi_w;
        lo_h = d->seg[t].dh < lo_h ? d->seg[t].dh : lo_h; hi_h = d->seg[t].dh > hi_h ? d->seg[t].dh : hi_h;
      }
      if (last > first) {
        span_w = hi_w - lo_w > span_w ? hi_w - lo_w : span_w;
        span_h = hi_h - lo_h > span_h ? hi_h - lo_h : span_h;
      }
    }
    // taps must be listed in row-major (dh, dw) order: pair mode takes LBO = offset(second tap) - offset(first tap) >= 0
    bool ordered = true;
    for (int t = 1; t < d->n_seg && ordered; ++t)
      ordered = d->seg[t].dh > d->seg[t - 1].dh || (d->seg[t].dh == d->seg[t - 1].dh && d->seg[t].dw >= d->seg[t - 1].dw);
    const int hw = 8 + span_w, hh = 8 + span_h;
    const int box = hw * hh * 128, box_pad = (box + 1023) / 1024 * 1024;
    const int n_pa = p.pair_mode ? 1 : 2;
    if (ordered && span_w <= 8 && span_h <= 8 && n_pa * box_pad * 10 < 2 * per * kAtomBytes * 7) {
      p.halo = 1;
      p.halo_w = hw;
      p.halo_box_bytes = box_pad;
      p.halo_tx_bytes = box;
      p.n_seg_groups = groups;
      p.taps_per_cta = per;
      p.stage_bytes = qa * kAtomBytes + n_pa * box_pad;
      if (int rc = make_view_map(&p.p_halo_map, d->p[d->seg[0].map_id], hw, hh)) {
        delete h;
        return rc;
      }
    }
  }
  if (!p.halo) {
    // smem: each stage holds the Q tile + 2 atoms per unit; keep at least 3 stages
    while (tmax > 1 && (qa + 2 * tmax) * kAtomBytes * 3 > 200 * 1024) --tmax;
    // balanced groups: e.g. 9 taps with room for 4 per CTA -> 3 + 3 + 3 instead of 4 + 4 + 1
    p.n_seg_groups = (n_units + tmax - 1) / tmax;
    p.taps_per_cta = (n_units + p.n_seg_groups - 1) / p.n_seg_groups;
    p.stage_bytes = (qa + 2 * p.taps_per_cta) * kAtomBytes;
  }
  p.PC = d->PC;
  p.QC = d->QC;
  p.p_tiles = (d->PC + 127) / 128;
  p.q_tiles = (d->QC + bq - 1) / bq;
  p.TW = d->TW;
  p.TH = d->TH;
  p.tiles_w = (d->OW + d->TW - 1) / d->TW;
  p.tiles_h = (d->OH + d->TH - 1) / d->TH;
  long long nblk = (long long)p.tiles_w * p.tiles_h * d->NB;
  if (nblk <= 0 || nblk > 0x7fffffffLL) {
    set_error("bad pixel block count %lld", nblk);
    delete h;
    return 2;
  }
  p.n_pix_blocks = (int)nblk;
  p.stages = (200 * 1024) / p.stage_bytes;
  if (p.stages > 8) p.stages = 8;
  int cols = p.taps_per_cta * bq;
  p.tmem_cols = 32;
  while (p.tmem_cols < cols) p.tmem_cols *= 2;
  int sms = max_ctas();      // urso_set_max_ctas: leave SMs to a concurrent kernel (overlapped NCCL all-reduce)
  if (sms <= 0) sms = 148;
  int items = p.n_seg_groups * p.p_tiles * p.q_tiles;
  int split = d->split_k;
  if (split <= 0) {
    // aim for just under 2 full waves (never a third, mostly empty one); at least 8 K-steps per CTA
    const int waves = 1;   // one wave: every CTA pays the prologue + atomic-reduction epilogue once (measured: 2 waves 4.4 ms, 1 wave 3.96 ms)
    split = (waves * sms) / items;
    if (split < 1) split = 1;
    if (items * split < sms && items * (split + 1) <= waves * sms) ++split;
    int max_split = p.n_pix_blocks / 8;
    if (max_split < 1) max_split = 1;
    if (split > max_split) split = max_split;
  }
  if (split > p.n_pix_blocks) split = p.n_pix_blocks;
  if (split < 1) split = 1;
  p.split_k = split;
  p.interleave = max_ctas() < num_sms() ? 1 : 0;   // planned on a subset of the SMs = one half of an L2-sharing pair
  p.order = 1;   // output tiles fastest (measured +3.5 % over split-fastest: co-resident CTAs share operand tiles through L2)
  p.g = d->g;
  p.g_seg_stride = d->g_seg_stride;
  p.g_sp = d->g_sp;
  p.g_sq = d->g_sq;
  // bulk-reduce epilogue: rows contiguous along q and 16-byte aligned, row buffers fit into the operand ring
  const long long row_buf = 128LL * (bq * 4 + 16);
  p.bulk_red = (d->g_sq == 1 && d->QC % 4 == 0 && d->g_sp % 4 == 0 && d->g_seg_stride % 4 == 0 &&
                (reinterpret_cast<uintptr_t>(d->g) & 15) == 0 && (long long)p.stages * p.stage_bytes >= row_buf) ? 1 : 0;
  p.red_bufs = (long long)p.stages * p.stage_bytes >= 2 * row_buf ? 2 : 1;
  h->grid = items * split;
  h->smem_bytes = p.stages * p.stage_bytes + 256 + 1024;
  *out = h;
  return 0;
}

extern "C" int urso_wgrad_launch(urso_wgrad_t* h, void* stream) {
  URSO_REQUIRE(h != nullptr, "null handle");
  URSO_REQUIRE(!urso::dry_run(), "dry run: nothing can be launched");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (h->block_q) {
    case 64: return launch_wgrad<64>(h, s);
    case 128: return launch_wgrad<128>(h, s);
    case 256: return launch_wgrad<256>(h, s);
  }
  urso::set_error("unsupported BLOCK_Q %d", h->block_q);
  return 2;
}

extern "C" void urso_wgrad_destroy(urso_wgrad_t* h) { delete h; }

/* plan introspection: writes {block_q, halo, halo_w, stages, stage_bytes, units per CTA, tap groups, split_k, grid, pair_mode} */
extern "C" int urso_wgrad_plan_info(const urso_wgrad_t* h, int32_t* out10) {
  URSO_REQUIRE(h != nullptr && out10 != nullptr, "null argument");
  const urso::WgradParams& p = h->params;
  const int32_t v[10] = {h->block_q, p.halo, p.halo_w, p.stages, p.stage_bytes, p.taps_per_cta, p.n_seg_groups, p.split_k,
                         h->grid, p.pair_mode};
  for (int i = 0; i < 10; ++i) out10[i] = v[i];
  return 0;
}
