"""Host-side geometry of the implicit-GEMM engines: how a Keras Conv2D (net.py:101-152,171,225-235,639) and its
two gradients map onto K-segments, tap shifts, strided "phase" views and staged weight matrices.

Pure Python / index arithmetic -- no GPU needed, unit-tested on CPU against the oracle by emulating the engines
(tests/test_convplan.py).  The conventions (see include/urso_b200.h):

  Engine F   D[pix, n] = sum_seg sum_chunk  A[seg.map][pix + (dh,dw), chunk*64 : +64] . Bmat[n, k]
  Engine W   G[seg][p, q] = sum_pix  P[seg.map][pix + (dh,dw), p] * Q[pix, q]

Out-of-range pixels of a view read as zero (TMA OOB fill) -- that is the convolution's zero padding.
Stride-2 convolutions address their input through the four parity ("phase") views x[:, ph::2, pw::2, :], so a tap
(r, s) with q = r - pad_t reads phase q mod 2 at offset floor(q / 2).
"""
from dataclasses import dataclass
from typing import List, Tuple


def same_pad(n: int, k: int, s: int) -> Tuple[int, int]:
    """TF 'SAME' padding (before, after) -- SURVEY App. C."""
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def ceil64(c: int) -> int:
    return (c + 63) // 64 * 64


def pick_patch(oh: int, ow: int, npix: int) -> Tuple[int, int]:
    """(TW, TH) with TW*TH == npix (powers of two, TW, TH <= 256) covering an oh x ow grid with least waste."""
    best = None
    tw = 1
    while tw <= min(npix, 256):
        th = npix // tw
        if th <= 256:
            cover = -(-ow // tw) * tw * (-(-oh // th) * th)
            key = (cover, -tw)
            if best is None or key < best[0]:
                best = (key, (tw, th))
        tw *= 2
    return best[1]


@dataclass
class ConvGeom:
    """One Conv2D: kernel kh x kw, stride (1 or 2), explicit top/left padding, channel counts, input size."""
    kh: int
    kw: int
    stride: int
    pad_t: int
    pad_l: int
    cin: int
    cout: int
    h: int
    w: int
    oh: int
    ow: int


def make_geom(kh, stride, padding, cin, cout, h, w) -> ConvGeom:
    """padding: 'same' (TF), 'valid', or an int (explicit symmetric ZeroPadding2D followed by a valid conv)."""
    if padding == "same":
        pt, pb = same_pad(h, kh, stride)
        pl, pr = same_pad(w, kh, stride)
    elif padding == "valid":
        pt = pb = pl = pr = 0
    else:
        pt = pb = pl = pr = int(padding)
    oh = (h + pt + pb - kh) // stride + 1
    ow = (w + pl + pr - kh) // stride + 1
    assert stride in (1, 2)
    return ConvGeom(kh, kh, stride, pt, pl, cin, cout, h, w, oh, ow)


def decimated_geom(g: ConvGeom) -> ConvGeom:
    """Geometry of a stride-1 conv whose OUTPUT GRADIENT is non-zero on the even-even pixels only (it sits behind a
    1x1/stride-2 conv): for its dgrad / wgrad it acts as the same filter at stride 2 on the decimated output grid
    du[:, ::2, ::2, :]."""
    assert g.stride == 1
    return ConvGeom(g.kh, g.kw, 2, g.pad_t, g.pad_l, g.cin, g.cout, g.h, g.w, (g.oh + 1) // 2, (g.ow + 1) // 2)


# ------------------------------------------------------------------------------------------------ forward / wgrad
def n_phase_views(stride: int) -> int:
    return 1 if stride == 1 else 4


def input_views(x, stride: int):
    """The A / P operand views of a conv input tensor x [N,H,W,C] (torch tensor or anything sliceable)."""
    if stride == 1:
        return [x]
    return [x[:, ph::2, pw::2, :] for ph in (0, 1) for pw in (0, 1)]


def fwd_taps(g: ConvGeom) -> List[Tuple[int, int, int, int]]:
    """[(tap index r*kw+s, map_id, dh, dw)] in K order."""
    out = []
    for r in range(g.kh):
        for s in range(g.kw):
            qh, qw = r - g.pad_t, s - g.pad_l
            if g.stride == 1:
                out.append((r * g.kw + s, 0, qh, qw))
            else:
                out.append((r * g.kw + s, (qh % 2) * 2 + (qw % 2), qh // 2, qw // 2))
    return out


def fwd_segments(g: ConvGeom):
    """Engine-F segments [(map_id, dh, dw, c_chunks)] and the weight gather index (len K) for stage_weight_rows:
    idx[k] = row of the HWIO kernel viewed as [kh*kw*cin, cout], -1 for channel padding."""
    cp = ceil64(g.cin)
    segs, idx = [], []
    for tap, m, dh, dw in fwd_taps(g):
        segs.append((m, dh, dw, cp // 64))
        idx.extend([tap * g.cin + c if c < g.cin else -1 for c in range(cp)])
    return segs, idx


def wgrad_segments(g: ConvGeom):
    """Engine-W segments [(map_id, dh, dw)], one per filter tap, in HWIO tap order (so G is [taps, cin, cout])."""
    return [(m, dh, dw) for _tap, m, dh, dw in fwd_taps(g)]


# ------------------------------------------------------------------------------------------------ dgrad
def dgrad_phases(g: ConvGeom):
    """Input-gradient launches.  Returns [(oph, opw, segs, tap_map)]:
      dx[:, oph::s, opw::s, :][pix] = sum_seg du[pix + (dh,dw)] . Wt[ci, slot*COp + co]
    segs are Engine-F segments over the single view du (map 0); tap_map[slot] is the HWIO tap feeding that slot.
    A phase with no contributing tap gets segs == [] (its dx is identically zero)."""
    s = g.stride
    cop = ceil64(g.cout)
    out = []
    for oph in range(s):
        for opw in range(s):
            segs, tap_map = [], []
            for r in range(g.kh):
                if (oph + g.pad_t - r) % s:
                    continue
                for c in range(g.kw):
                    if (opw + g.pad_l - c) % s:
                        continue
                    segs.append((0, (oph + g.pad_t - r) // s, (opw + g.pad_l - c) // s, cop // 64))
                    tap_map.append(r * g.kw + c)
            out.append((oph, opw, segs, tap_map))
    return out


# ------------------------------------------------------------------------------------------------ stem (7x7 / s2, C=3)
STEM_K = 256  # 4 row taps x 64 packed values


def stem_weight_index(cin: int = 3) -> List[int]:
    """K index -> row of the 7x7xcin HWIO kernel for the space-to-depth staged stem (urso_stem_stage layout):
    k = r2*64 + s2*16 + ph*8 + pw*4 + c  <->  tap (r, s) = (2*r2 + ph, 2*s2 + pw), channel c."""
    idx = []
    for r2 in range(4):
        for s2 in range(4):
            for ph in range(2):
                for pw in range(2):
                    for c in range(4):
                        r, s = 2 * r2 + ph, 2 * s2 + pw
                        idx.append((r * 7 + s) * cin + c if (r < 7 and s < 7 and c < cin) else -1)
    return idx


def stem_segments():
    """Engine-F / Engine-W segments over the staged tensor E [B, H/2+3, W/2, 64]: 4 row taps."""
    return [(0, r2, 0, 1) for r2 in range(4)]


def stem_grad_row_map(cin: int = 3) -> List[int]:
    """HWIO row (r*7+s)*cin+c -> row of the staged wgrad G[4*64, cout]."""
    inv = [0] * (49 * cin)
    for k, src in enumerate(stem_weight_index(cin)):
        if src >= 0:
            inv[src] = k
    return inv
