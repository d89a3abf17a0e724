/* liburso_b200.so -- C-ABI of the B200-native UrsoNet hot path.
 *
 * The reference (pedropro/UrsoNet) has NO native/FFI boundary: its hot path is the Keras graph built in
 * net.py:85-352,639-643, the losses net.py:705-762 and the optimizer wired in net.py:973-1017, all executed by
 * TensorFlow library kernels.  Each entry point below replaces the TF op family named in its comment.  The
 * Python host (ursonet_b200/net.py, mirroring net.UrsoNet) binds them with ctypes; INTEGRATION.md shows the
 * stub a maintainer of the reference would add.
 *
 * Conventions: every pointer is a caller-owned DEVICE pointer unless stated; activations are NHWC; every launch
 * takes a cudaStream_t (passed as void*), never synchronises, never allocates device memory and is CUDA-graph
 * capturable.  Return value: 0 = ok, non-zero = error, message via urso_last_error() (thread-local).
 */
#ifndef URSO_B200_H
#define URSO_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define URSO_MAX_AMAPS 8
#define URSO_MAX_SEGS 32

int urso_version(void);
const char* urso_last_error(void);
int urso_num_sms(void);
/* Limit the persistent Engine-F grids planned from now on to n CTAs (0 = one per SM): leaves SMs free for kernels that
 * must run concurrently (an overlapped NCCL all-reduce), and lets tests exercise deep per-CTA tile queues at small sizes. */
void urso_set_max_ctas(int n);
/* Dry run (tests of the host-side planners on a machine without a GPU): while on, *_create calls plan tiles, shared
 * memory and pipelines for a 148-SM device but encode no tensor maps and touch no device memory; launches fail. */
void urso_set_dry_run(int on);
/* Programmatic dependent launch of the two tcgen05 engines (default on): a launch may run its prologue (barrier init, TMEM
 * allocation, tensor-map prefetch) while the previous kernel of the stream drains; it reads that kernel's results only after
 * griddepcontrol.wait.  0 = plain stream order. */
void urso_set_pdl(int on);
/* Engine F launches planned while this is on (default) accumulate their addend (residual / gradient fan-in) on the tensor
 * core -- per 64-channel chunk one extra K step "addend tile x 64x64 identity" into the chunk's TMEM columns -- instead of
 * loading, unpacking and adding it in the epilogue warps; bit-exact products, fp32 accumulation.  0 = epilogue add. */
void urso_set_residual_mma(int on);
/* Engine W halo mode (default on): weight-gradient launches of multi-tap filters on one stride-1 view (3x3 convolutions, the
 * stem's four row taps) load ONE 8x8-pixel block + halo per 64-channel atom and K step and address every tap as a shifted
 * window of it, instead of one operand atom per tap.  0 = one atom per tap (A/B). */
void urso_set_wgrad_halo(int on);
/* N-split tail of Engine F (default on).  A launch whose tile count is not a multiple of the CTA count ends with a partial
 * wave in which most SMs idle.  K-heavy BLOCK_N = 256 launches in stream mode cut the tiles of that last wave along N into 2
 * or 4 sub-tiles of 128 / 64 output channels, one per CTA: a sub-tile streams the same pixel tile but only its slice of the
 * weight tile and issues narrower MMAs, so the last wave costs about half a wave; no reduction is involved and every output
 * element is accumulated in the same order as before (bit-identical results).  0 = whole tiles only (A/B). */
void urso_set_tail_split(int on);
/* struct sizes, so that FFI bindings can verify their layout against this header */
int urso_sizeof_convgemm_desc(void);
int urso_sizeof_wgrad_desc(void);
int urso_sizeof_conv2d_fwd_desc(void);
int urso_sizeof_conv2d_dgrad_desc(void);
int urso_sizeof_conv2d_wgrad_desc(void);

/* A 4-D NHWC bf16 view (possibly strided): element (n,h,w,c) at base + n*stride_n + h*stride_h + w*stride_w + c.
 * C must be a multiple of 64 for MMA operands; strides are in ELEMENTS and must be multiples of 8. */
typedef struct {
  const void* base;
  int32_t C, W, H, N;
  int64_t stride_w, stride_h, stride_n;
} urso_view4;

/* One K-segment of an implicit GEMM: `c_chunks` 64-channel chunks read from view `map_id`, shifted by (dh,dw)
 * pixels relative to the output pixel (out-of-bounds pixels read as zero: this IS the conv padding). */
typedef struct {
  int32_t map_id, dh, dw, c_chunks;
} urso_seg;

/* Output / addend / mask pixel addressing: element (n,h,w,c) at ptr + n*sn + h*sh + w*sw + c. */
typedef struct {
  void* ptr;
  int64_t sn, sh, sw;
} urso_pix;

/* ---- Engine F: implicit-GEMM convolution on tcgen05 (replaces Conv2D fprop and dgrad: net.py:101-111,138-152,
 * 171,225-235,639 and their TF autodiff input-gradients) fused with the BatchNorm/bias/ReLU/residual epilogue
 * (net.py:103-116) or, for dgrad, with the ReLU-mask and gradient fan-in add.
 *   D[pixel, n] = sum_seg sum_chunk A_seg[pixel + (dh,dw), 64-ch chunk] . Bmat[n, k]      (fp32 accumulate in TMEM)
 *   out = mask( relu( D + shift[n] + addend[pixel, n] ) )                                 (each stage optional)
 * Bmat is bf16 [b_rows, b_k] row-major (K contiguous), k enumerating (segment, chunk, channel) in order. */
typedef struct {
  urso_view4 a[URSO_MAX_AMAPS];
  int32_t n_a;
  const void* b;
  int32_t b_rows, b_k;
  urso_seg seg[URSO_MAX_SEGS];
  int32_t n_seg;
  int32_t OW, OH, NB; /* output pixel grid the tiles cover */
  int32_t TW, TH;     /* pixel patch per 128-row tile, TW*TH == 128, powers of two */
  urso_pix out;
  int32_t out_fp32;   /* 0: bf16 output, 1: fp32 output */
  const float* shift; /* [b_rows] or NULL */
  urso_pix addend;    /* bf16, ptr NULL = none */
  urso_pix mask;      /* bf16, ptr NULL = none: output forced to 0 where mask <= 0 */
  int32_t relu;
  float* colsum;      /* [b_rows] fp32 or NULL: atomically accumulates per-channel sums of the stored output */
  int32_t block_n;    /* N tile: 0 = auto, else 32 / 64 / 128 / 256 */
  /* Bit-packed ReLU masks (1 bit per element instead of re-reading the bf16 activation in backward: 16x fewer bytes).
   * One 32-bit word per pixel and 32-channel group g: channel 32g + c (c = 0..31)  <->  bit (7 - (c >> 2)) + 8 (c & 3)
   * (the order the epilogues produce / consume with one byte-permute per channel pair).
   * ptr + strides in BYTES addressing the word group of pixel (n,h,w) of the OUTPUT grid: ptr + n*sn + h*sh + w*sw.
   * relu_bits (fprop, TMA epilogue only): written with (stored output != 0).  mask_bits (dgrad): output forced to 0
   * where the bit is clear; use it INSTEAD of `mask`. */
  urso_pix relu_bits;
  urso_pix mask_bits;
  int32_t halo;       /* 1: halo reuse -- all segments are taps of ONE stride-1 view (n_a == 1, TW == 8, TH == 16): each
                         64-channel chunk of the input patch (+ its halo) is fetched once and every tap reads a
                         row-shifted window of it from shared memory instead of re-fetching it from L2; when the weight
                         operand has one N tile and fits (<= 100 KB) it is kept resident in shared memory as well */
} urso_convgemm_desc;

typedef struct urso_convgemm urso_convgemm_t;
int urso_convgemm_create(const urso_convgemm_desc* d, urso_convgemm_t** out);
int urso_convgemm_launch(urso_convgemm_t* h, void* stream);
void urso_convgemm_destroy(urso_convgemm_t* h);
/* plan introspection (tests / profiling): out9 = {block_n, npipe, stages, kpack, halo, bres, a_stages, smem_bytes, grid} */
int urso_convgemm_plan_info(const urso_convgemm_t* h, int32_t* out9);
int urso_convgemm_tail_split(const urso_convgemm_t* h);   /* sub-tiles per tile of the last partial wave (1 = no N-split tail) */

/* ---- Engine W: weight-gradient GEMM on tcgen05 (replaces Conv2DBackpropFilter of TF autodiff).
 *   G[t][p, q] (+)= sum_pixels  P_t[pixel + (dh_t,dw_t), p] * Q[pixel, q]        t = 0..n_seg-1
 * P views are the layer inputs (tap-shifted), Q is the output gradient.  fp32 result accumulated with atomics
 * into g + t*g_seg_stride + p*g_sp + q*g_sq (caller zeroes g). */
typedef struct {
  urso_view4 p[URSO_MAX_AMAPS];
  int32_t n_p;
  urso_view4 q;
  urso_seg seg[URSO_MAX_SEGS]; /* c_chunks unused */
  int32_t n_seg;
  int32_t PC, QC;     /* valid channels of P (rows of G) and Q (cols of G) */
  int32_t OW, OH, NB; /* pixel grid of Q */
  int32_t TW, TH;     /* pixel patch per K block, TW*TH == 64 */
  float* g;
  int64_t g_seg_stride, g_sp, g_sq;
  int32_t split_k;    /* 0 = auto */
  int32_t block_q;    /* Q (N) tile: 0 = auto, else 64 / 128 / 256 */
} urso_wgrad_desc;

typedef struct urso_wgrad urso_wgrad_t;
int urso_wgrad_create(const urso_wgrad_desc* d, urso_wgrad_t** out);
int urso_wgrad_launch(urso_wgrad_t* h, void* stream);
void urso_wgrad_destroy(urso_wgrad_t* h);
/* plan introspection: out10 = {block_q, halo, halo_w, stages, stage_bytes, units per CTA, tap groups, split_k, grid, pair_mode} */
int urso_wgrad_plan_info(const urso_wgrad_t* h, int32_t* out10);

/* =====================================================================================================================
 * Conv2D OPERATORS (the level a non-Python host binds): a Keras Conv2D (net.py:101-111,138-152,171,225-235,639), its
 * input gradient and its weight gradient, described by tensor shapes, stride and explicit padding.  All planning --
 * K-segments, tap shifts, stride-2 parity views, pixel-patch choice, N tile, weight-operand layout -- happens inside
 * the library (csrc/conv_ops.cu); the Engine-F / Engine-W entry points above are what these operators are built from.
 *
 * Tensors: x bf16 NHWC dense [N,H,W,C]; w fp32 Keras HWIO master weights [k,k,C,K]; y / dy bf16 NHWC
 * [N,OH,OW,ceil64(K)] (fp32 when out_fp32); OH = (H + pad_t + pad_b - k) / stride + 1.
 * The frozen-BatchNorm scale[K] = gamma/sqrt(var+eps) is folded into the staged bf16 weight operand; shift[K] is added
 * in the epilogue (urso_bn_fold produces both).  Each operator owns no device memory: the caller passes a workspace of
 * *_workspace_bytes() bytes (staged operand + gather indices), which must stay valid while the handle lives.
 * `*_stage_weights` re-stages the operand from the fp32 masters (call it after every weight update, any stream order
 * before the launch); `*_launch` is graph-capturable and never synchronises.
 * The 7x7/stride-2 stem (net.py:170-171,254-255) is the same operator with ksize = 7: x is then the compact staged tensor
 * of urso_stem_stage ([N, H/2+3, W/2+3, 16]) and H, W are the IMAGE dimensions. */
typedef struct {
  int32_t N, H, W, C; /* input activation */
  int32_t K;          /* output channels */
  int32_t ksize;      /* 1, 3 (or 7 for the stem) */
  int32_t stride;     /* 1 or 2 */
  int32_t pad_t, pad_l, pad_b, pad_r;
} urso_conv2d_shape;

/* TF 'SAME' padding of one dimension (out = ceil(n/s), total = max((out-1)*s + k - n, 0), before = total/2). */
void urso_same_pad(int32_t n, int32_t k, int32_t s, int32_t* before, int32_t* after);

typedef struct {
  urso_conv2d_shape shape;
  const void* x;
  const float* w;      /* fp32 HWIO */
  const float* scale;  /* fp32 [K] or NULL (= 1) */
  const float* shift;  /* fp32 [K] or NULL */
  const void* addend;  /* bf16 [N,OH,OW,K] residual input or NULL */
  void* y;
  int32_t relu;
  int32_t out_fp32;
  void* workspace;
  void* relu_bits;     /* uint32 [N,OH,OW,K/32] or NULL: bit-packed (y != 0), the ReLU mask backward needs (see
                          urso_convgemm_desc.relu_bits for the bit order); K % 64 == 0, bf16 output */
} urso_conv2d_fwd_desc;

typedef struct urso_conv2d_fwd urso_conv2d_fwd_t;
int64_t urso_conv2d_fwd_workspace_bytes(const urso_conv2d_shape* s);
int urso_conv2d_fwd_create(const urso_conv2d_fwd_desc* d, urso_conv2d_fwd_t** out);
int urso_conv2d_fwd_stage_weights(urso_conv2d_fwd_t* h, void* stream);
int urso_conv2d_fwd_launch(urso_conv2d_fwd_t* h, void* stream);
void urso_conv2d_fwd_destroy(urso_conv2d_fwd_t* h);
int urso_conv2d_fwd_plan_info(const urso_conv2d_fwd_t* h, int32_t* out9); /* see urso_convgemm_plan_info */
int urso_conv2d_fwd_tail_split(const urso_conv2d_fwd_t* h);                  /* see urso_convgemm_tail_split */
/* the staging work of this operator as a job of urso_stage_weights_multi (declared below) */
int urso_conv2d_fwd_stage_job(const urso_conv2d_fwd_t* h, void* stage_job_out);

/* Input gradient with fused fan-in: dx = mask( sum_i dgrad_i(dy_i, w_i) + addend ), one launch per output phase
 * (stride^2 phases), the convolutions' reduction ranges concatenated.  All consumers share x's shape and stride.
 *   mask    bf16 [N,H,W,C]: the forward activation x (ReLU backward: dx = 0 where x <= 0) or NULL
 *   mask_bits: the same mask, bit-packed (preferred)
 *   addend  bf16 [N,H,W,C]: gradient arriving through an identity shortcut, or NULL
 *   colsum  fp32 [C]: += per-channel sums of dx (d beta / d bias of the layer that produced x), or NULL
 *   dy_sparse: every dy_i is non-zero on its even-even pixels only (it sits behind a 1x1/stride-2 convolution, the
 *           Keras-v1 bottleneck block); stride-1 consumers then run at stride 2 on the decimated gradient grid.
 * Phases that receive no filter tap (e.g. the odd pixels of a 1x1/stride-2 conv) are NOT written: dx must have been
 * zero-filled once at allocation (urso_conv2d_dgrad_untouched_phases reports them as a bit mask, bit = oph*stride+opw). */
#define URSO_MAX_FANIN 4
typedef struct {
  int32_t n_convs;
  urso_conv2d_shape shape[URSO_MAX_FANIN];
  const void* dy[URSO_MAX_FANIN];
  const float* w[URSO_MAX_FANIN];
  const float* scale[URSO_MAX_FANIN];
  int32_t dy_sparse;
  const void* mask;
  const void* addend;
  void* dx;
  float* colsum;
  void* workspace;
  const void* mask_bits; /* uint32 [N,H,W,C/32] or NULL: the ReLU mask of x as written by urso_conv2d_fwd's relu_bits;
                            replaces `mask` (16x fewer bytes, no shared-memory ring in the epilogue) */
} urso_conv2d_dgrad_desc;

typedef struct urso_conv2d_dgrad urso_conv2d_dgrad_t;
int64_t urso_conv2d_dgrad_workspace_bytes(const urso_conv2d_dgrad_desc* d);
int urso_conv2d_dgrad_create(const urso_conv2d_dgrad_desc* d, urso_conv2d_dgrad_t** out);
int urso_conv2d_dgrad_stage_weights(urso_conv2d_dgrad_t* h, void* stream);
int urso_conv2d_dgrad_launch(urso_conv2d_dgrad_t* h, void* stream);
int urso_conv2d_dgrad_untouched_phases(const urso_conv2d_dgrad_t* h);
int urso_conv2d_dgrad_num_launches(const urso_conv2d_dgrad_t* h);
int urso_conv2d_dgrad_plan_info(const urso_conv2d_dgrad_t* h, int32_t launch, int32_t* out9); /* see urso_convgemm_plan_info */
int urso_conv2d_dgrad_tail_split(const urso_conv2d_dgrad_t* h, int32_t launch);                /* see urso_convgemm_tail_split */
/* the staging work of this operator as jobs of urso_stage_weights_multi: writes up to max_jobs urso_stage_job structs,
 * returns how many the operator has (negative on error) */
int urso_conv2d_dgrad_stage_jobs(const urso_conv2d_dgrad_t* h, void* stage_jobs_out, int32_t max_jobs);
void urso_conv2d_dgrad_destroy(urso_conv2d_dgrad_t* h);

/* Raw weight gradient G[k*k*C, K] (fp32, HWIO order, accumulated with atomics: the caller zeroes it) of the UNSCALED
 * convolution; urso_conv_param_grads turns it into dW / dbias / dgamma / dbeta.  Stem (ksize 7): G is [4*64, K] in the
 * staged K order (urso_conv_param_grads takes the row map). */
typedef struct {
  urso_conv2d_shape shape;
  const void* x;
  const void* dy;
  int32_t dy_sparse;
  float* G;
} urso_conv2d_wgrad_desc;

typedef struct urso_conv2d_wgrad urso_conv2d_wgrad_t;
int urso_conv2d_wgrad_create(const urso_conv2d_wgrad_desc* d, urso_conv2d_wgrad_t** out);
int urso_conv2d_wgrad_launch(urso_conv2d_wgrad_t* h, void* stream);
int urso_conv2d_wgrad_plan_info(const urso_conv2d_wgrad_t* h, int32_t* out10); /* see urso_wgrad_plan_info */
void urso_conv2d_wgrad_destroy(urso_conv2d_wgrad_t* h);
/* HWIO row (r*7+s)*3+c -> row of the staged stem gradient G[4*64, K]; fills map[147]. */
void urso_stem_grad_row_map(int32_t* map147);

/* ---- Stem input staging (replaces mold_image net.py:1337-1348 + ZeroPadding2D(3) net.py:170,254 + the im2col TF does
 * internally): uint8 or fp32 RGB [B,H,W,3] -> compact bf16 space-to-depth tensor S[B, H/2+3, W/2+3, 16],
 * S[b,h2,w2,(ph,pw,c)] = (img[2*h2+ph-3, 2*w2+pw-3, c] - mean[c]) or 0 outside the image / for c==3.
 * The 7x7/s2 stem reads S through an overlapping 64-"channel" view E[b,h2,wo,k] = S_flat[(h2*(W/2+3) + wo)*16 + k]
 * (pixel stride 16 elements: k = s2*16 + ph*8 + pw*4 + c covers the staged pixels wo..wo+3, the four horizontal taps);
 * urso_conv2d_{fwd,wgrad} with ksize 7 build that view themselves from the pointer to S. */
int urso_stem_stage(const void* img, int32_t img_is_u8, int32_t subtract_mean, const float* mean3, void* e_out,
                    int32_t B, int32_t H, int32_t W, int32_t part, void* stream);
/* part: 0 = bf16(v) (normal); 1 = bf16(v - bf16(v)), the low half of a split-bf16 pair (parity mode, see below). */

/* ---- MaxPooling2D 3x3/s2 'same' on even maps (net.py:176,258): TF pads bottom/right only. bf16 NHWC.
 * argmax (uint8 [B,H/2,W/2,C], may be NULL for inference) records the FIRST maximum of each window (dr*3+ds). */
int urso_maxpool_fwd(const void* x, void* y, void* argmax, int32_t B, int32_t H, int32_t W, int32_t C, void* stream);
/* dx = scatter(dy to the recorded argmax) (TF MaxPoolGrad).  x (the pooled tensor's input) may be NULL when dy is
 * already masked by (pooled > 0) -- equivalent to the stem's ReLU mask, since a window's max is 0 only if all its
 * (post-ReLU) inputs are 0; otherwise dx is additionally masked by x > 0. */
int urso_maxpool_bwd(const void* x, const void* argmax, const void* dy, void* dx, float* colsum, int32_t B, int32_t H,
                     int32_t W, int32_t C, void* stream);
/* colsum (fp32 [C], may be NULL): atomically accumulates the per-channel sums of dx (d beta of the stem's BatchNorm). */

/* ---- Dense layers of the two heads (net.py:302,316,336,345,350): fp32, small batch, weight-bandwidth bound.
 * y[B,N] = act(x[B,K] @ w[K,N] + bias).  act: 0 linear, 1 relu. y must be zeroed by the caller (split-K atomics);
 * urso_dense_bias_act finishes it. */
int urso_dense_fwd(const float* x, const float* w, float* y, int32_t B, int32_t K, int32_t N, void* stream);
int urso_dense_bias_act(float* y, const float* bias, int32_t B, int32_t N, int32_t act, void* stream);
/* dy is masked in place by (y>0) when act==1; dw[K,N] = x^T dy (overwrites), db[N] = colsum(dy), dx[B,K] = dy w^T. */
int urso_dense_bwd(const float* x, const float* w, const float* y, float* dy, float* dx, float* dw, float* db,
                   int32_t B, int32_t K, int32_t N, int32_t act, void* stream);

/* ---- Losses (net.py:705-762) with their gradients; scalars are written to loss_out[0] (already weighted). */
/* softmax_loss_graph on ReLU'd logits: loss = w/B * sum_b -sum_k y_k log_softmax(z)_k ; dz = w/B (softmax(z)*sum_k y - y) */
int urso_softmax_xent(const float* z, const float* y, float* dz, float* loss_out, int32_t B, int32_t N, float weight,
                      void* stream);
/* rel_loss_graph: ||y-p||_F / ||y||_F over the whole [B,3] tensor. */
int urso_rel_loss(const float* pred, const float* gt, float* dpred, float* loss_out, int32_t B, int32_t N, float weight,
                  void* stream);
/* ori_q head: q = l2_normalize(raw) (net.py:346), loss = mean_b(1-|<gt,q>|) (net.py:724-733); draw = d loss / d raw. */
int urso_quat_head(const float* raw, const float* gt, float* q_out, float* draw, float* loss_out, int32_t B,
                   float weight, void* stream);

/* ---- Parameter plumbing --------------------------------------------------------------------------------- */
/* scale[c] = gamma/sqrt(var+eps), shift[c] = (bias-mean)*scale+beta  (BatchNorm(training=False), net.py:60-76).
 * gamma == NULL: layer without BN (scale = 1, shift = bias or 0). */
int urso_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, const float* bias,
                 float eps, float* scale, float* shift, int32_t C, void* stream);
/* Conv weight staging, Keras HWIO fp32 kernel viewed as w[R = KH*KW*CI][CO] -> bf16 K-major GEMM operands with the
 * frozen-BN scale folded in:
 *   rows (fprop):  out[row, k] = w[idx[k], row] * scale[row]                 row < rows_out (zero rows beyond CO)
 *   cols (dgrad):  out[ci, slot*COp + co] = w[tap[slot]*CI + ci, co] * scale[co]
 * idx / tap are device int32 arrays; -1 selects zero padding.  ld_out = row pitch of `out` in elements, so that
 * several convolutions can be staged side by side into one K-concatenated operand (fused fan-in dgrad). */
int urso_stage_weight_rows(const float* w, const float* scale, void* out, const int32_t* idx_dev, int32_t K, int32_t CO,
                           int32_t rows_out, int64_t ld_out, int32_t part, void* stream);
int urso_stage_weight_cols(const float* w, const float* scale, void* out, const int32_t* tap_dev, int32_t n_slots,
                           int32_t CI, int32_t CO, int32_t COp, int32_t rows_out, int64_t ld_out, void* stream);
/* From the raw wgrad G[R'][CO] (fp32; row r of the HWIO kernel lives at G row g_row_map[r], identity if NULL) and
 * colsum[c] = sum_pixels du: dW = scale*G, dbias = scale*colsum, dbeta = colsum,
 * dgamma = rstd*(sum_r W[r,c]*G[r,c] + (bias-mean)*colsum).  gamma == NULL: no BN; dbias == NULL: no bias.
 * s_scratch: zeroed fp32 [CO] accumulator for sum_r W*G (needed when gamma != NULL). */
int urso_conv_param_grads(const float* G, const int32_t* g_row_map_dev, const float* w, const float* colsum,
                          const float* scale, const float* gamma, const float* mean, const float* var,
                          const float* bias, float eps, float* dW, float* dbias, float* dgamma, float* dbeta,
                          float* s_scratch, int32_t R, int32_t CO, void* stream);

/* ---- Multi-tensor versions of the three small per-layer kernels above: ONE launch covers every layer of the network
 * through a device-resident job table (host code fills the structs, copies the table to the device once at plan time).
 * Next to the persistent convolution CTAs ~330 tiny launches per step effectively serialised with the convolutions
 * (1.4 ms of a 13.7 ms RN-50 step); the tables bring that to a handful of launches.
 *   urso_bn_fold_multi           one job per convolution (gamma == NULL: no BN)
 *   urso_stage_weights_multi     kind 0 = urso_stage_weight_rows (tile transpose through shared memory), kind 1 =
 *                                urso_stage_weight_cols; urso_conv2d_{fwd,dgrad}_stage_job(s) export the jobs of operators
 *   urso_conv_param_grads_multi  urso_conv_param_grads for many convolutions (S zeroed by the caller)
 * *_jobs_finalize assign each job its first block and return the grid size; begins[] is the same list as an array. */
typedef struct {
  const float *gamma, *beta, *mean, *var, *bias;
  float *scale, *shift;
  int32_t C, pad_;
} urso_bn_job;
typedef struct {
  const float* w;
  const float* scale;
  void* out;
  const int32_t* index; /* kind 0: idx[K]; kind 1: tap[n_slots] (device) */
  int32_t kind, K;      /* K: kind 0 = columns of the operand; kind 1 = number of tap slots */
  int32_t CI, CO, COp, rows_out;
  int64_t ld_out;
  int32_t part, block_begin;
} urso_stage_job;
typedef struct {
  const float* G;
  const int32_t* g_row_map;
  const float *w, *colsum, *scale, *gamma, *mean, *var, *bias;
  float *dW, *dbias, *dgamma, *dbeta, *S;
  int32_t R, CO, rows_per_slab, cblocks, block_begin, pad_;
} urso_pgrad_job;
int urso_sizeof_bn_job(void);
int urso_sizeof_stage_job(void);
int urso_sizeof_pgrad_job(void);
int urso_bn_fold_multi(const urso_bn_job* jobs_dev, int32_t n_jobs, int32_t max_c, float eps, void* stream);
int32_t urso_stage_jobs_finalize(urso_stage_job* jobs_host, int32_t n, int32_t* begins_out);
int urso_stage_weights_multi(const urso_stage_job* jobs_dev, const int32_t* begins_dev, int32_t n_jobs, int32_t total_blocks,
                             void* stream);
int32_t urso_pgrad_jobs_finalize(urso_pgrad_job* jobs_host, int32_t n, int32_t* begins_out);
int urso_conv_param_grads_multi(const urso_pgrad_job* jobs_dev, const int32_t* begins_dev, int32_t n_jobs,
                                int32_t total_blocks, int32_t max_co, float eps, void* stream);

/* ---- Optimizer (net.py:979-983,1008-1012; Keras-2 SGD / Adam(amsgrad) with global-norm clipnorm) over flat arenas.
 * chunk_coef[i] applies to elements [256 i, 256 i + 256): reg gradient 2*wd/size(w) (0 for gamma/beta);
 * chunk_lr[i] is 1 for trainable chunks, 0 for frozen ones (their gradient is zeroed and excluded from the norm).
 * hyper_dev (device fp32[8]): [0]=lr (Adam: lr_t), [1]=momentum|beta1, [2]=beta2, [3]=eps, [4]=clipnorm.        */
/* Gradient accumulation over micro-batches (BASELINE configs[4]: global batch 256 on fewer than 8 GPUs):
 * acc = beta * acc + grad (beta = 0 on the first micro-batch); when out != NULL the result alpha * acc is written to out
 * (the last micro-batch: out = the gradient arena, alpha = 1 / number of micro-batches) and acc is left untouched. */
int urso_grad_accumulate(float* acc, const float* grad, float* out, float beta, float alpha, int64_t n, void* stream);
/* sumsq_out: device fp32[URSO_SUMSQ_SCRATCH]; [0] receives sum(g^2), the rest is scratch for per-block partial sums, which
 * are combined in a fixed order: the norm (and with it the clip factor and every updated weight) is bit-reproducible, so
 * data-parallel ranks that hold the same all-reduced gradient apply the IDENTICAL update. */
#define URSO_SUMSQ_SCRATCH 2048
int urso_add_reg_sumsq(float* grad, const float* param, const float* chunk_coef, const float* chunk_lr,
                       float grad_scale, float* sumsq_out, int64_t n, void* stream);
int urso_sgd_step(float* param, float* vel, const float* grad, const float* chunk_lr, const float* sumsq,
                  const float* hyper_dev, int64_t n, void* stream);
int urso_amsgrad_step(float* param, float* m, float* v, float* vhat, const float* grad, const float* chunk_lr,
                      const float* sumsq, const float* hyper_dev, int64_t n, void* stream);

/* ---- split-bf16 parity mode (forward only): every fp32 value x is carried as hi = bf16(x), lo = bf16(x - hi) and a
 * convolution is evaluated as A_hi.W_hi + A_hi.W_lo + A_lo.W_hi on the tensor cores (three K-segments of Engine F, fp32
 * accumulation) -- ~2^-16 relative error per operand instead of 2^-9, which meets the 1e-3 forward-parity gate.
 * urso_split_f32: v = y (+ addend); optional ReLU; out32 = v (may be NULL); hi/lo = the bf16 pair of v. */
int urso_split_f32(const float* y, const float* addend, float* out32, void* hi, void* lo, int64_t n, int32_t relu,
                   void* stream);
int urso_maxpool_fwd_f32(const float* x, float* y, int32_t B, int32_t H, int32_t W, int32_t C, void* stream);

/* ---- orientation soft labels on the device (data format either side of the path).
 * urso_encode_ori: utils.encode_ori_fast (utils.py:319-346) for a batch: quats [B,4] fp32, hquat [nbins,4] fp32 (the
 * Euler-grid quaternions, 16-byte aligned), redundant [nbins] uint8 mask, var = (beta/n)^2/12 -> enc [B,nbins] fp32.
 * urso_decode_ori_moments: stable_softmax of the network's logits (utils.py:26-28) and the 4x4 moment matrix
 * A_b = sum_k p_k H_k H_k^T of se3lib.quat_weighted_avg (se3lib.py:241-248); the quaternion is the principal
 * eigenvector of A_b (a 4x4 eigen-solve, left to the host). */
int urso_encode_ori(const float* quats, const float* hquat, const uint8_t* redundant, float* enc, int32_t B,
                    int32_t nbins, float var, void* stream);
int urso_decode_ori_moments(const float* logits, const float* hquat, float* A, int32_t B, int32_t nbins, void* stream);

/* ---- sim2real augmentation on the device (replaces net.py:390-406 = luma + imgaug pipeline on the host).
 * One record per image, drawn on the host (ursonet_b200/augment.py):
 *   apply        0: luma only (the reference augments with p = 0.5, net.py:395)
 *   order[5]     the drawn order of the operations (iaa.Sequential(random_order=True)): 0 AdditiveGaussianNoise,
 *                1 GaussianBlur, 2 Add, 3 Multiply, 4 CoarseDropout
 *   noise_q      round(65536 * sigma / 147.8) (sigma = 0.01 * 255; Irwin-Hall(4) integer noise, see augment.cu)
 *   blur_w[5]    normalised 5-tap Gaussian weights (cv2.getGaussianKernel(5, sigma)), blur_sigma < 1e-3 skips the blur
 *   add, mul     Add / Multiply parameters;  drop_thresh = p * 2^32, drop_h x drop_w = low-resolution mask size
 *   win[4]       (y1, x1, y2, x2) of the un-padded image inside the pad64 frame: only these pixels are augmented and
 *                the blur reflects at its borders (the reference pads AFTER augmenting)
 * src, dst: uint8 [B,H,W,3], out of place.  Bit-exact against oracle/sim2real_oracle.py. */
typedef struct urso_aug_params {
  int32_t apply;
  int32_t order[5];
  int32_t noise_q;
  uint32_t noise_seed;
  float blur_sigma;
  float blur_w[5];
  int32_t add;
  float mul;
  uint32_t drop_thresh;
  int32_t drop_h, drop_w;
  uint32_t drop_seed;
  int32_t win[4];
} urso_aug_params;
int urso_sizeof_aug_params(void);
int urso_sim2real_aug(const uint8_t* src, uint8_t* dst, const urso_aug_params* params_dev, int32_t B, int32_t H,
                      int32_t W, void* stream);

/* ---- small elementwise helpers */
int urso_cast_f32_to_bf16(const float* x, void* y, int64_t n, void* stream);
int urso_cast_bf16_to_f32(const void* x, float* y, int64_t n, void* stream);
/* dst[r, 0:Cpad] (bf16) = src[r, 0:C] + src2[r, 0:C] (fp32; src2 may be NULL), zero padded to Cpad channels: sums the
 * two heads' input gradients and stages them as an Engine-F / Engine-W operand. */
int urso_pad_cast_rows(const float* src, const float* src2, void* dst, int64_t rows, int32_t C, int32_t Cpad,
                       void* stream);
int urso_colsum_bf16(const void* x, float* out, int64_t rows, int32_t C, void* stream);

#ifdef __cplusplus
}
#endif
#endif
