// Conv2D operators of the C-ABI (include/urso_b200.h, "Conv2D OPERATORS"): host-side planning that turns a Keras Conv2D
// (net.py:101-111,138-152,171,225-235,639), its input gradient and its weight gradient into Engine-F / Engine-W launches.
//
//   Engine F   D[pix, n]    = sum_seg sum_chunk A[seg.map][pix + (dh,dw), chunk*64 : +64] . Bmat[n, k]
//   Engine W   G[seg][p, q] = sum_pix P[seg.map][pix + (dh,dw), p] * Q[pix, q]
//
// Out-of-range pixels of a view read as zero (TMA OOB fill): that IS the convolution's zero padding, so explicit /
// TF-'SAME' asymmetric padding is only a different tap shift.  A stride-2 convolution addresses its input through the
// four parity views x[:, ph::2, pw::2, :]: a tap (r, s) with q = r - pad_t reads parity q mod 2 at offset floor(q / 2).
// The input gradient of a stride-s convolution is produced one output parity ("phase") at a time; several consumers of
// the same tensor are K-concatenated into one launch (fused gradient fan-in).
#include <vector>

#include "common.cuh"

namespace {

using urso::set_error;

inline int ceil64(int c) { return (c + 63) / 64 * 64; }
inline int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
inline int posmod(int a, int b) { return a - floordiv(a, b) * b; }

struct Geom {
  int k, stride, pad_t, pad_l, cin, cout, h, w, oh, ow;
};

int make_geom(const urso_conv2d_shape& s, Geom* g) {
  URSO_REQUIRE(s.N >= 1 && s.H >= 1 && s.W >= 1 && s.C >= 1 && s.K >= 1, "bad conv shape");
  URSO_REQUIRE(s.stride == 1 || s.stride == 2, "stride %d unsupported (1 or 2)", s.stride);
  URSO_REQUIRE(s.ksize >= 1 && s.ksize <= 7, "ksize %d unsupported", s.ksize);
  g->k = s.ksize; g->stride = s.stride; g->pad_t = s.pad_t; g->pad_l = s.pad_l;
  g->cin = s.C; g->cout = s.K; g->h = s.H; g->w = s.W;
  g->oh = (s.H + s.pad_t + s.pad_b - s.ksize) / s.stride + 1;
  g->ow = (s.W + s.pad_l + s.pad_r - s.ksize) / s.stride + 1;
  URSO_REQUIRE(g->oh >= 1 && g->ow >= 1, "empty conv output");
  return 0;
}

// A stride-1 conv whose OUTPUT GRADIENT is non-zero on the even-even pixels only acts, for its gradients, as the same
// filter at stride 2 on the decimated output grid dy[:, ::2, ::2, :].
Geom decimated(const Geom& g) {
  Geom d = g;
  d.stride = 2;
  d.oh = (g.oh + 1) / 2;
  d.ow = (g.ow + 1) / 2;
  return d;
}

// (TW, TH) with TW*TH == npix (powers of two <= 256) covering an oh x ow grid with least waste (ties: widest).
void pick_patch(int oh, int ow, int npix, int* tw_out, int* th_out) {
  long long best = -1;
  for (int tw = 1; tw <= npix && tw <= 256; tw *= 2) {
    const int th = npix / tw;
    if (th > 256) continue;
    const long long cover = (long long)((ow + tw - 1) / tw) * tw * ((oh + th - 1) / th) * th;
    if (best < 0 || cover <= best) {   // later (wider) candidates win ties
      best = cover;
      *tw_out = tw;
      *th_out = th;
    }
  }
}

// Halo mode of Engine F (one TMA box per channel chunk feeds every filter tap from shared memory) wants an 8 x 16 pixel
// patch; take it when that patch covers the map with at most 7 % more pixels than the best patch would.
bool halo_patch_ok(int oh, int ow) {
  int tw, th;
  pick_patch(oh, ow, 128, &tw, &th);
  const long long best = (long long)((ow + tw - 1) / tw) * tw * ((oh + th - 1) / th) * th;
  const long long halo = (long long)((ow + 7) / 8) * 8 * ((oh + 15) / 16) * 16;
  return halo * 100 <= best * 107;
}

// Engine W's halo mode wants 8 x 8 pixel blocks; take them when they cover the map with at most 7 % more pixels than the
// best 64-pixel patch would.
bool wgrad_block8_ok(int oh, int ow) {
  int tw, th;
  pick_patch(oh, ow, 64, &tw, &th);
  const long long best = (long long)((ow + tw - 1) / tw) * tw * ((oh + th - 1) / th) * th;
  const long long b8 = (long long)((ow + 7) / 8) * 8 * ((oh + 7) / 8) * 8;
  return b8 * 100 <= best * 107;
}

struct Tap {
  int tap, map, dh, dw;
};

std::vector<Tap> fwd_taps(const Geom& g) {
  std::vector<Tap> out;
  for (int r = 0; r < g.k; ++r)
    for (int s = 0; s < g.k; ++s) {
      const int qh = r - g.pad_t, qw = s - g.pad_l;
      if (g.stride == 1) out.push_back({r * g.k + s, 0, qh, qw});
      else out.push_back({r * g.k + s, posmod(qh, 2) * 2 + posmod(qw, 2), floordiv(qh, 2), floordiv(qw, 2)});
    }
  return out;
}

urso_view4 dense_view(const void* base, int n, int h, int w, int c) {
  urso_view4 v;
  v.base = base; v.C = c; v.W = w; v.H = h; v.N = n;
  v.stride_w = c; v.stride_h = (int64_t)w * c; v.stride_n = (int64_t)h * w * c;
  return v;
}
// x[:, ph::step, pw::step, :] of a dense bf16 NHWC tensor
urso_view4 strided_view(const void* base, int n, int h, int w, int c, int ph, int pw, int step) {
  urso_view4 v;
  v.base = static_cast<const char*>(base) + ((int64_t)ph * w + pw) * c * 2;
  v.C = c; v.W = (w - pw + step - 1) / step; v.H = (h - ph + step - 1) / step; v.N = n;
  v.stride_w = (int64_t)step * c; v.stride_h = (int64_t)step * w * c; v.stride_n = (int64_t)h * w * c;
  return v;
}
// The stem's operand: an OVERLAPPING view over the compact staged tensor S[N, H/2+3, W/2+3, 16] of urso_stem_stage: view pixel
// (h2, wo) has 64 "channels" = the staged pixels wo .. wo+3 (pixel stride 16 elements), the four horizontal filter taps.
urso_view4 stem_view(const void* base, int n, int H, int W) {
  const int64_t h2 = H / 2 + 3, w2 = W / 2 + 3;
  urso_view4 v;
  v.base = base; v.C = 64; v.W = W / 2; v.H = (int32_t)h2; v.N = n;
  v.stride_w = 16; v.stride_h = w2 * 16; v.stride_n = h2 * w2 * 16;
  return v;
}
urso_view4 flat_view(const void* base, int64_t m, int c) {
  urso_view4 v;
  v.base = base; v.C = c; v.W = (int32_t)m; v.H = 1; v.N = 1;
  v.stride_w = c; v.stride_h = m * c; v.stride_n = m * c;
  return v;
}
urso_pix dense_pix(const void* base, int h, int w, int c) {
  urso_pix p;
  p.ptr = const_cast<void*>(base); p.sn = (int64_t)h * w * c; p.sh = (int64_t)w * c; p.sw = c;
  return p;
}
urso_pix strided_pix(const void* base, int h, int w, int c, int ph, int pw, int step, int esize) {
  urso_pix p;
  p.ptr = base ? const_cast<char*>(static_cast<const char*>(base)) + ((int64_t)ph * w + pw) * c * esize : nullptr;
  p.sn = (int64_t)h * w * c; p.sh = (int64_t)step * w * c; p.sw = (int64_t)step * c;
  return p;
}
urso_pix flat_pix(const void* base, int64_t m, int c) {
  urso_pix p;
  p.ptr = const_cast<void*>(base); p.sn = m * c; p.sh = m * c; p.sw = c;
  return p;
}
const urso_pix kNoPix = {nullptr, 0, 0, 0};
// bit-packed mask tensor [N,H,W,C/32] uint32 (C/8 bytes per pixel): BYTE strides of the (phase-strided / flattened) pixel grid
urso_pix bits_strided_pix(const void* base, int h, int w, int c, int ph, int pw, int step) {
  const int64_t pb = c / 8;
  urso_pix p;
  p.ptr = const_cast<char*>(static_cast<const char*>(base)) + ((int64_t)ph * w + pw) * pb;
  p.sn = (int64_t)h * w * pb; p.sh = (int64_t)step * w * pb; p.sw = (int64_t)step * pb;
  return p;
}
urso_pix bits_flat_pix(const void* base, int64_t m, int c) {
  urso_pix p;
  p.ptr = const_cast<void*>(base); p.sn = m * (c / 8); p.sh = m * (c / 8); p.sw = c / 8;
  return p;
}

// K index -> row of the 7x7x3 HWIO kernel for the space-to-depth staged stem (urso_stem_stage layout):
// k = r2*64 + s2*16 + ph*8 + pw*4 + c  <->  tap (r, s) = (2*r2 + ph, 2*s2 + pw), channel c;  -1 = zero padding.
std::vector<int32_t> stem_weight_index(int cin) {
  std::vector<int32_t> idx;
  for (int r2 = 0; r2 < 4; ++r2)
    for (int s2 = 0; s2 < 4; ++s2)
      for (int ph = 0; ph < 2; ++ph)
        for (int pw = 0; pw < 2; ++pw)
          for (int c = 0; c < 4; ++c) {
            const int r = 2 * r2 + ph, s = 2 * s2 + pw;
            idx.push_back((r < 7 && s < 7 && c < cin) ? (r * 7 + s) * cin + c : -1);
          }
  return idx;
}

int upload_i32(void* dst, const std::vector<int32_t>& v) {
  if (urso::dry_run()) return 0;
  URSO_CUDA_OK(cudaMemcpy(dst, v.data(), v.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  return 0;
}
inline int64_t align256(int64_t n) { return (n + 255) / 256 * 256; }

bool is_stem(const urso_conv2d_shape& s) { return s.ksize == 7; }
int check_stem(const urso_conv2d_shape& s) {
  URSO_REQUIRE(s.stride == 2 && s.pad_t == 3 && s.pad_l == 3 && s.pad_b == 3 && s.pad_r == 3 && s.C == 3 &&
                   s.H % 2 == 0 && s.W % 2 == 0,
               "ksize 7 is the stem: 3 channels, stride 2, padding 3, even image size");
  return 0;
}

}  // namespace

// =========================================================================================================== forward
struct urso_conv2d_fwd {
  urso_convgemm_t* plan = nullptr;
  const float* w = nullptr;
  const float* scale = nullptr;
  void* bmat = nullptr;
  int32_t* idx = nullptr;
  int K = 0, cout = 0;
};

extern "C" void urso_same_pad(int32_t n, int32_t k, int32_t s, int32_t* before, int32_t* after) {
  const int out = (n + s - 1) / s;
  int total = (out - 1) * s + k - n;
  if (total < 0) total = 0;
  *before = total / 2;
  *after = total - total / 2;
}

static int fwd_k(const urso_conv2d_shape& s) { return is_stem(s) ? 256 : s.ksize * s.ksize * ceil64(s.C); }

extern "C" int64_t urso_conv2d_fwd_workspace_bytes(const urso_conv2d_shape* s) {
  if (s == nullptr) return -1;
  const int64_t K = fwd_k(*s);
  return align256((int64_t)s->K * K * 2) + align256(K * 4);
}

extern "C" int urso_conv2d_fwd_create(const urso_conv2d_fwd_desc* d, urso_conv2d_fwd_t** out) {
  URSO_REQUIRE(d != nullptr && out != nullptr, "null argument");
  URSO_REQUIRE(d->x && d->w && d->y && d->workspace, "null tensor / workspace");
  const urso_conv2d_shape& s = d->shape;
  Geom g;
  if (int rc = make_geom(s, &g)) return rc;
  urso_convgemm_desc cd;
  memset(&cd, 0, sizeof(cd));
  std::vector<int32_t> idx;
  if (is_stem(s)) {
    if (int rc = check_stem(s)) return rc;
    URSO_REQUIRE(d->addend == nullptr, "the stem has no residual input");
    cd.a[0] = stem_view(d->x, s.N, s.H, s.W);
    cd.n_a = 1;
    for (int r2 = 0; r2 < 4; ++r2) cd.seg[r2] = urso_seg{0, r2, 0, 1};
    cd.n_seg = 4;
    idx = stem_weight_index(3);
    cd.OW = g.ow; cd.OH = g.oh; cd.NB = s.N;
    pick_patch(g.oh, g.ow, 128, &cd.TW, &cd.TH);
    if (halo_patch_ok(g.oh, g.ow)) {    // the 4 row taps read one (16+3) x 8 box
      cd.TW = 8; cd.TH = 16; cd.halo = 1;
    }
    cd.out = dense_pix(d->y, g.oh, g.ow, s.K);
    if (d->relu_bits) cd.relu_bits = bits_strided_pix(d->relu_bits, g.oh, g.ow, s.K, 0, 0, 1);
  } else {
    URSO_REQUIRE(s.C % 64 == 0, "input channels %d must be a multiple of 64", s.C);
    const std::vector<Tap> taps = fwd_taps(g);
    URSO_REQUIRE((int)taps.size() <= URSO_MAX_SEGS, "too many filter taps");
    const int cp = ceil64(s.C);
    for (size_t i = 0; i < taps.size(); ++i) {
      cd.seg[i] = urso_seg{taps[i].map, taps[i].dh, taps[i].dw, cp / 64};
      for (int c = 0; c < cp; ++c) idx.push_back(c < s.C ? taps[i].tap * s.C + c : -1);
    }
    cd.n_seg = (int)taps.size();
    const int yc = s.K;   // channels of y as allocated
    if (s.ksize == 1 && s.stride == 1 && s.pad_t == 0 && s.pad_l == 0 && s.pad_b == 0 && s.pad_r == 0) {
      // pointwise: the whole batch is one long row of pixels
      const int64_t M = (int64_t)s.N * g.oh * g.ow;
      URSO_REQUIRE(M <= 0x7fffffffLL, "too many pixels");
      cd.a[0] = flat_view(d->x, M, s.C);
      cd.n_a = 1;
      cd.OW = (int32_t)M; cd.OH = 1; cd.NB = 1; cd.TW = 128; cd.TH = 1;
      cd.out = flat_pix(d->y, M, yc);
      cd.addend = d->addend ? flat_pix(d->addend, M, yc) : kNoPix;
      if (d->relu_bits) cd.relu_bits = bits_flat_pix(d->relu_bits, M, yc);
    } else {
      if (s.stride == 1) {
        cd.a[0] = dense_view(d->x, s.N, s.H, s.W, s.C);
        cd.n_a = 1;
      } else {
        for (int ph = 0; ph < 2; ++ph)
          for (int pw = 0; pw < 2; ++pw) cd.a[ph * 2 + pw] = strided_view(d->x, s.N, s.H, s.W, s.C, ph, pw, 2);
        cd.n_a = 4;
      }
      cd.OW = g.ow; cd.OH = g.oh; cd.NB = s.N;
      pick_patch(g.oh, g.ow, 128, &cd.TW, &cd.TH);
      if (s.stride == 1 && s.ksize == 3 && halo_patch_ok(g.oh, g.ow)) {
        cd.TW = 8; cd.TH = 16; cd.halo = 1;
      }
      cd.out = dense_pix(d->y, g.oh, g.ow, yc);
      cd.addend = d->addend ? dense_pix(d->addend, g.oh, g.ow, yc) : kNoPix;
      if (d->relu_bits) cd.relu_bits = bits_strided_pix(d->relu_bits, g.oh, g.ow, yc, 0, 0, 1);
    }
  }
  const int K = (int)idx.size();
  auto* h = new urso_conv2d_fwd();
  h->w = d->w; h->scale = d->scale; h->K = K; h->cout = s.K;
  h->bmat = d->workspace;
  h->idx = reinterpret_cast<int32_t*>(static_cast<char*>(d->workspace) + align256((int64_t)s.K * K * 2));
  if (int rc = upload_i32(h->idx, idx)) {
    delete h;
    return rc;
  }
  cd.b = h->bmat; cd.b_rows = s.K; cd.b_k = K;
  cd.out_fp32 = d->out_fp32;
  cd.shift = d->shift;
  cd.relu = d->relu;
  if (int rc = urso_convgemm_create(&cd, &h->plan)) {
    delete h;
    return rc;
  }
  *out = h;
  return 0;
}

extern "C" int urso_conv2d_fwd_stage_weights(urso_conv2d_fwd_t* h, void* stream) {
  URSO_REQUIRE(h != nullptr, "null handle");
  return urso_stage_weight_rows(h->w, h->scale, h->bmat, h->idx, h->K, h->cout, h->cout, h->K, 0, stream);
}
extern "C" int urso_conv2d_fwd_launch(urso_conv2d_fwd_t* h, void* stream) {
  URSO_REQUIRE(h != nullptr, "null handle");
  return urso_convgemm_launch(h->plan, stream);
}
extern "C" int urso_conv2d_fwd_stage_job(const urso_conv2d_fwd_t* h, void* out_) {
  URSO_REQUIRE(h != nullptr && out_ != nullptr, "null argument");
  urso_stage_job* j = static_cast<urso_stage_job*>(out_);
  memset(j, 0, sizeof(*j));
  j->w = h->w; j->scale = h->scale; j->out = h->bmat; j->index = h->idx;
  j->kind = 0; j->K = h->K; j->CO = h->cout; j->rows_out = h->cout; j->ld_out = h->K;
  return 0;
}
extern "C" int urso_conv2d_fwd_tail_split(const urso_conv2d_fwd_t* h) { return h != nullptr ? urso_convgemm_tail_split(h->plan) : 0; }
extern "C" int urso_conv2d_fwd_plan_info(const urso_conv2d_fwd_t* h, int32_t* out9) {
  URSO_REQUIRE(h != nullptr, "null handle");
  return urso_convgemm_plan_info(h->plan, out9);
}
extern "C" void urso_conv2d_fwd_destroy(urso_conv2d_fwd_t* h) {
  if (h == nullptr) return;
  urso_convgemm_destroy(h->plan);
  delete h;
}

// =========================================================================================================== dgrad
namespace {
struct DgradPart {     // one consumer's slice of one phase operand
  int conv, koff, cop;
  std::vector<int32_t> tap_map;
  int32_t* tap_dev = nullptr;
};
struct DgradPhase {
  int oph, opw, ktot;
  std::vector<urso_seg> segs;
  std::vector<DgradPart> parts;
  void* bmat = nullptr;
  urso_convgemm_t* plan = nullptr;
};

// Taps of conv g that contribute to output phase (oph, opw) of dx: dx[phase][pix] += dy[pix + (dh,dw)] . W[tap]
void phase_taps(const Geom& g, int oph, int opw, int map_id, std::vector<urso_seg>* segs, std::vector<int32_t>* tap_map) {
  const int s = g.stride, cop = ceil64(g.cout);
  for (int r = 0; r < g.k; ++r) {
    if (posmod(oph + g.pad_t - r, s)) continue;
    for (int c = 0; c < g.k; ++c) {
      if (posmod(opw + g.pad_l - c, s)) continue;
      segs->push_back(urso_seg{map_id, floordiv(oph + g.pad_t - r, s), floordiv(opw + g.pad_l - c, s), cop / 64});
      tap_map->push_back(r * g.k + c);
    }
  }
}

int dgrad_plan(const urso_conv2d_dgrad_desc* d, std::vector<Geom>* geoms, std::vector<DgradPhase>* phases, int* stride_out,
               int* untouched) {
  URSO_REQUIRE(d->n_convs >= 1 && d->n_convs <= URSO_MAX_FANIN, "n_convs=%d out of range", d->n_convs);
  const urso_conv2d_shape& s0 = d->shape[0];
  for (int i = 0; i < d->n_convs; ++i) {
    const urso_conv2d_shape& s = d->shape[i];
    URSO_REQUIRE(!is_stem(s), "the stem has no input gradient");
    URSO_REQUIRE(s.N == s0.N && s.H == s0.H && s.W == s0.W && s.C == s0.C && s.stride == s0.stride,
                 "fan-in consumers must share the input tensor and the stride");
    URSO_REQUIRE(!d->dy_sparse || s.stride == 1, "dy_sparse needs stride-1 consumers");
    Geom g;
    if (int rc = make_geom(s, &g)) return rc;
    geoms->push_back(d->dy_sparse ? decimated(g) : g);
  }
  const int stride = (*geoms)[0].stride;
  URSO_REQUIRE(!(d->addend != nullptr && stride != 1), "a gradient addend needs stride-1 consumers");
  *stride_out = stride;
  *untouched = 0;
  for (int oph = 0; oph < stride; ++oph)
    for (int opw = 0; opw < stride; ++opw) {
      DgradPhase ph;
      ph.oph = oph; ph.opw = opw; ph.ktot = 0;
      for (int i = 0; i < d->n_convs; ++i) {
        DgradPart part;
        part.conv = i; part.koff = ph.ktot; part.cop = ceil64((*geoms)[i].cout);
        phase_taps((*geoms)[i], oph, opw, i, &ph.segs, &part.tap_map);
        if (part.tap_map.empty()) continue;
        ph.ktot += (int)part.tap_map.size() * part.cop;
        ph.parts.push_back(part);
      }
      if (ph.segs.empty()) {
        *untouched |= 1 << (oph * stride + opw);
        continue;
      }
      URSO_REQUIRE((int)ph.segs.size() <= URSO_MAX_SEGS, "too many K segments in one dgrad phase");
      phases->push_back(ph);
    }
  return 0;
}
}  // namespace

struct urso_conv2d_dgrad {
  std::vector<DgradPhase> phases;
  std::vector<Geom> geoms;
  const float* w[URSO_MAX_FANIN];
  const float* scale[URSO_MAX_FANIN];
  int cin = 0, untouched = 0;
};

extern "C" int64_t urso_conv2d_dgrad_workspace_bytes(const urso_conv2d_dgrad_desc* d) {
  if (d == nullptr) return -1;
  std::vector<Geom> geoms;
  std::vector<DgradPhase> phases;
  int stride, untouched;
  if (dgrad_plan(d, &geoms, &phases, &stride, &untouched)) return -1;
  int64_t total = 0;
  for (const DgradPhase& ph : phases) {
    total += align256((int64_t)d->shape[0].C * ph.ktot * 2);
    for (const DgradPart& pt : ph.parts) total += align256((int64_t)pt.tap_map.size() * 4);
  }
  return total;
}

extern "C" int urso_conv2d_dgrad_create(const urso_conv2d_dgrad_desc* d, urso_conv2d_dgrad_t** out) {
  URSO_REQUIRE(d != nullptr && out != nullptr, "null argument");
  URSO_REQUIRE(d->dx && d->workspace, "null tensor / workspace");
  auto* h = new urso_conv2d_dgrad();
  int stride;
  if (int rc = dgrad_plan(d, &h->geoms, &h->phases, &stride, &h->untouched)) {
    delete h;
    return rc;
  }
  const urso_conv2d_shape& s0 = d->shape[0];
  const int N = s0.N, H = s0.H, W = s0.W, cin = s0.C;
  h->cin = cin;
  bool flat_ok = stride == 1;
  for (int i = 0; i < d->n_convs; ++i) {
    h->w[i] = d->w[i];
    h->scale[i] = d->scale[i];
    const urso_conv2d_shape& s = d->shape[i];
    flat_ok = flat_ok && s.ksize == 1 && s.pad_t == 0 && s.pad_l == 0;
    if (d->dy[i] == nullptr || d->w[i] == nullptr) {
      set_error("null dy / w for consumer %d", i);
      delete h;
      return 2;
    }
  }
  char* ws = static_cast<char*>(d->workspace);
  int rc = 0;
  for (DgradPhase& ph : h->phases) {
    ph.bmat = ws;
    ws += align256((int64_t)cin * ph.ktot * 2);
    for (DgradPart& pt : ph.parts) {
      pt.tap_dev = reinterpret_cast<int32_t*>(ws);
      ws += align256((int64_t)pt.tap_map.size() * 4);
      if ((rc = upload_i32(pt.tap_dev, pt.tap_map))) break;
    }
    if (rc) break;
    urso_convgemm_desc cd;
    memset(&cd, 0, sizeof(cd));
    cd.n_a = d->n_convs;
    for (size_t i = 0; i < ph.segs.size(); ++i) cd.seg[i] = ph.segs[i];
    cd.n_seg = (int)ph.segs.size();
    cd.b = ph.bmat; cd.b_rows = cin; cd.b_k = ph.ktot;
    cd.colsum = d->colsum;
    if (flat_ok) {
      const int64_t M = (int64_t)N * H * W;
      for (int i = 0; i < d->n_convs; ++i) cd.a[i] = flat_view(d->dy[i], M, ceil64(d->shape[i].K));
      cd.OW = (int32_t)M; cd.OH = 1; cd.NB = 1; cd.TW = 128; cd.TH = 1;
      cd.out = flat_pix(d->dx, M, cin);
      cd.addend = d->addend ? flat_pix(d->addend, M, cin) : kNoPix;
      cd.mask = d->mask ? flat_pix(d->mask, M, cin) : kNoPix;
      if (d->mask_bits) cd.mask_bits = bits_flat_pix(d->mask_bits, M, cin);
    } else {
      for (int i = 0; i < d->n_convs; ++i) {
        const Geom& g = h->geoms[i];    // (decimated) output grid of consumer i
        const int kc = ceil64(d->shape[i].K);
        if (d->dy_sparse) {
          Geom full;
          make_geom(d->shape[i], &full);
          cd.a[i] = strided_view(d->dy[i], N, full.oh, full.ow, kc, 0, 0, 2);
        } else {
          cd.a[i] = dense_view(d->dy[i], N, g.oh, g.ow, kc);
        }
      }
      const int th_ = (H - ph.oph + stride - 1) / stride, tw_ = (W - ph.opw + stride - 1) / stride;
      cd.OW = tw_; cd.OH = th_; cd.NB = N;
      pick_patch(th_, tw_, 128, &cd.TW, &cd.TH);
      if (stride == 1 && d->n_convs == 1 && d->shape[0].ksize == 3 && halo_patch_ok(th_, tw_)) {
        cd.TW = 8; cd.TH = 16; cd.halo = 1;
      }
      cd.out = strided_pix(d->dx, H, W, cin, ph.oph, ph.opw, stride, 2);
      cd.mask = d->mask ? strided_pix(d->mask, H, W, cin, ph.oph, ph.opw, stride, 2) : kNoPix;
      cd.addend = d->addend ? strided_pix(d->addend, H, W, cin, 0, 0, 1, 2) : kNoPix;
      if (d->mask_bits) cd.mask_bits = bits_strided_pix(d->mask_bits, H, W, cin, ph.oph, ph.opw, stride);
    }
    if ((rc = urso_convgemm_create(&cd, &ph.plan))) break;
  }
  if (rc) {
    urso_conv2d_dgrad_destroy(h);
    return rc;
  }
  *out = h;
  return 0;
}

extern "C" int urso_conv2d_dgrad_stage_weights(urso_conv2d_dgrad_t* h, void* stream) {
  URSO_REQUIRE(h != nullptr, "null handle");
  for (const DgradPhase& ph : h->phases)
    for (const DgradPart& pt : ph.parts) {
      const Geom& g = h->geoms[pt.conv];
      void* dst = static_cast<char*>(ph.bmat) + (int64_t)pt.koff * 2;
      if (int rc = urso_stage_weight_cols(h->w[pt.conv], h->scale[pt.conv], dst, pt.tap_dev, (int)pt.tap_map.size(), g.cin,
                                          g.cout, pt.cop, g.cin, ph.ktot, stream))
        return rc;
    }
  return 0;
}
extern "C" int urso_conv2d_dgrad_stage_jobs(const urso_conv2d_dgrad_t* h, void* out_, int32_t max_jobs) {
  if (h == nullptr || out_ == nullptr) return -1;
  urso_stage_job* jobs = static_cast<urso_stage_job*>(out_);
  int n = 0;
  for (const DgradPhase& ph : h->phases)
    for (const DgradPart& pt : ph.parts) {
      if (n < max_jobs) {
        const Geom& g = h->geoms[pt.conv];
        urso_stage_job* j = &jobs[n];
        memset(j, 0, sizeof(*j));
        j->w = h->w[pt.conv]; j->scale = h->scale[pt.conv];
        j->out = static_cast<char*>(ph.bmat) + (int64_t)pt.koff * 2;
        j->index = pt.tap_dev;
        j->kind = 1; j->K = (int32_t)pt.tap_map.size(); j->CI = g.cin; j->CO = g.cout; j->COp = pt.cop; j->rows_out = g.cin;
        j->ld_out = ph.ktot;
      }
      ++n;
    }
  return n;
}
extern "C" int urso_conv2d_dgrad_launch(urso_conv2d_dgrad_t* h, void* stream) {
  URSO_REQUIRE(h != nullptr, "null handle");
  for (const DgradPhase& ph : h->phases)
    if (int rc = urso_convgemm_launch(ph.plan, stream)) return rc;
  return 0;
}
extern "C" int urso_conv2d_dgrad_plan_info(const urso_conv2d_dgrad_t* h, int32_t launch, int32_t* out9) {
  URSO_REQUIRE(h != nullptr && launch >= 0 && launch < (int)h->phases.size(), "bad handle / launch index");
  return urso_convgemm_plan_info(h->phases[launch].plan, out9);
}
extern "C" int urso_conv2d_dgrad_tail_split(const urso_conv2d_dgrad_t* h, int32_t launch) {
  return (h != nullptr && launch >= 0 && launch < (int)h->phases.size()) ? urso_convgemm_tail_split(h->phases[launch].plan) : 0;
}
extern "C" int urso_conv2d_dgrad_untouched_phases(const urso_conv2d_dgrad_t* h) { return h ? h->untouched : -1; }
extern "C" int urso_conv2d_dgrad_num_launches(const urso_conv2d_dgrad_t* h) { return h ? (int)h->phases.size() : -1; }
extern "C" void urso_conv2d_dgrad_destroy(urso_conv2d_dgrad_t* h) {
  if (h == nullptr) return;
  for (DgradPhase& ph : h->phases)
    if (ph.plan) urso_convgemm_destroy(ph.plan);
  delete h;
}

// =========================================================================================================== wgrad
struct urso_conv2d_wgrad {
  urso_wgrad_t* plan = nullptr;
};

extern "C" int urso_conv2d_wgrad_create(const urso_conv2d_wgrad_desc* d, urso_conv2d_wgrad_t** out) {
  URSO_REQUIRE(d != nullptr && out != nullptr, "null argument");
  URSO_REQUIRE(d->x && d->dy && d->G, "null tensor");
  const urso_conv2d_shape& s = d->shape;
  Geom g;
  if (int rc = make_geom(s, &g)) return rc;
  urso_wgrad_desc wd;
  memset(&wd, 0, sizeof(wd));
  wd.g = d->G;
  if (is_stem(s)) {
    if (int rc = check_stem(s)) return rc;
    wd.p[0] = stem_view(d->x, s.N, s.H, s.W);
    wd.n_p = 1;
    for (int r2 = 0; r2 < 4; ++r2) wd.seg[r2] = urso_seg{0, r2, 0, 0};
    wd.n_seg = 4;
    wd.q = dense_view(d->dy, s.N, g.oh, g.ow, ceil64(s.K));
    wd.PC = 64; wd.QC = s.K; wd.OW = g.ow; wd.OH = g.oh; wd.NB = s.N;
    pick_patch(g.oh, g.ow, 64, &wd.TW, &wd.TH);
    if (urso::wgrad_halo_enabled() && wgrad_block8_ok(g.oh, g.ow)) wd.TW = wd.TH = 8;     // the 4 row taps read one 8 x 11 box
    wd.g_seg_stride = (int64_t)64 * s.K; wd.g_sp = s.K; wd.g_sq = 1;
  } else {
    URSO_REQUIRE(s.C % 64 == 0, "input channels %d must be a multiple of 64", s.C);
    URSO_REQUIRE(!d->dy_sparse || s.stride == 1, "dy_sparse needs a stride-1 convolution");
    const int kc = ceil64(s.K);
    const bool pointwise = s.ksize == 1 && s.stride == 1 && !d->dy_sparse && s.pad_t == 0 && s.pad_l == 0;
    if (pointwise) {
      const int64_t M = (int64_t)s.N * g.oh * g.ow;
      URSO_REQUIRE(M <= 0x7fffffffLL, "too many pixels");
      wd.seg[0] = urso_seg{0, 0, 0, 0};
      wd.n_seg = 1; wd.n_p = 1;
      wd.OW = (int32_t)M; wd.OH = 1; wd.NB = 1; wd.TW = 64; wd.TH = 1;
      wd.g_seg_stride = (int64_t)s.C * s.K;
      if (s.C < 128 && s.K >= 128) {   // wide side on the 128-row MMA M dimension; transposed accumulation into HWIO
        wd.p[0] = flat_view(d->dy, M, kc);
        wd.q = flat_view(d->x, M, s.C);
        wd.PC = s.K; wd.QC = s.C; wd.g_sp = 1; wd.g_sq = s.K;
      } else {
        wd.p[0] = flat_view(d->x, M, s.C);
        wd.q = flat_view(d->dy, M, kc);
        wd.PC = s.C; wd.QC = s.K; wd.g_sp = s.K; wd.g_sq = 1;
      }
    } else {
      const Geom ge = d->dy_sparse ? decimated(g) : g;
      const std::vector<Tap> taps = fwd_taps(ge);
      URSO_REQUIRE((int)taps.size() <= URSO_MAX_SEGS, "too many filter taps");
      for (size_t i = 0; i < taps.size(); ++i) wd.seg[i] = urso_seg{taps[i].map, taps[i].dh, taps[i].dw, 0};
      wd.n_seg = (int)taps.size();
      if (ge.stride == 1) {
        wd.p[0] = dense_view(d->x, s.N, s.H, s.W, s.C);
        wd.n_p = 1;
      } else {
        for (int ph = 0; ph < 2; ++ph)
          for (int pw = 0; pw < 2; ++pw) wd.p[ph * 2 + pw] = strided_view(d->x, s.N, s.H, s.W, s.C, ph, pw, 2);
        wd.n_p = 4;
      }
      wd.q = d->dy_sparse ? strided_view(d->dy, s.N, g.oh, g.ow, kc, 0, 0, 2) : dense_view(d->dy, s.N, g.oh, g.ow, kc);
      wd.PC = s.C; wd.QC = s.K; wd.OW = ge.ow; wd.OH = ge.oh; wd.NB = s.N;
      pick_patch(ge.oh, ge.ow, 64, &wd.TW, &wd.TH);
      if (urso::wgrad_halo_enabled() && ge.stride == 1 && wd.n_seg >= 2 && wgrad_block8_ok(ge.oh, ge.ow)) wd.TW = wd.TH = 8;
      wd.g_seg_stride = (int64_t)s.C * s.K; wd.g_sp = s.K; wd.g_sq = 1;
    }
  }
  auto* h = new urso_conv2d_wgrad();
  if (int rc = urso_wgrad_create(&wd, &h->plan)) {
    delete h;
    return rc;
  }
  *out = h;
  return 0;
}
extern "C" int urso_conv2d_wgrad_launch(urso_conv2d_wgrad_t* h, void* stream) {
  URSO_REQUIRE(h != nullptr, "null handle");
  return urso_wgrad_launch(h->plan, stream);
}
extern "C" int urso_conv2d_wgrad_plan_info(const urso_conv2d_wgrad_t* h, int32_t* out10) {
  URSO_REQUIRE(h != nullptr, "null handle");
  return urso_wgrad_plan_info(h->plan, out10);
}
extern "C" void urso_conv2d_wgrad_destroy(urso_conv2d_wgrad_t* h) {
  if (h == nullptr) return;
  urso_wgrad_destroy(h->plan);
  delete h;
}
extern "C" void urso_stem_grad_row_map(int32_t* map147) {
  const std::vector<int32_t> idx = stem_weight_index(3);
  for (int i = 0; i < 147; ++i) map147[i] = 0;
  for (size_t k = 0; k < idx.size(); ++k)
    if (idx[k] >= 0) map147[idx[k]] = (int32_t)k;
}

extern "C" int urso_sizeof_conv2d_fwd_desc(void) { return (int)sizeof(urso_conv2d_fwd_desc); }
extern "C" int urso_sizeof_conv2d_dgrad_desc(void) { return (int)sizeof(urso_conv2d_dgrad_desc); }
extern "C" int urso_sizeof_conv2d_wgrad_desc(void) { return (int)sizeof(urso_conv2d_wgrad_desc); }
