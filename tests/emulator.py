"""CPU emulation of the two tcgen05 engines' *semantics* (include/urso_b200.h) with plain torch indexing.
Test infrastructure: lets the host-side segment / phase-view / weight-staging logic (ursonet_b200/convplan.py) be
checked against the oracle without a GPU, and gives the GPU tests an independent expected value."""
import torch


def shifted(view, dh, dw, OH, OW, cpad):
    """out[n,h,w,c] = view[n,h+dh,w+dw,c] for in-range pixels/channels, else 0 (TMA out-of-bounds fill)."""
    N, H, W, C = view.shape
    out = torch.zeros(N, OH, OW, cpad, dtype=view.dtype)
    h_lo, h_hi = max(0, -dh), min(OH, H - dh)
    w_lo, w_hi = max(0, -dw), min(OW, W - dw)
    if h_lo < h_hi and w_lo < w_hi:
        c = min(C, cpad)
        out[:, h_lo:h_hi, w_lo:w_hi, :c] = view[:, h_lo + dh:h_hi + dh, w_lo + dw:w_hi + dw, :c]
    return out


def emu_convgemm(a_views, bmat, segs, OH, OW):
    """D[n,h,w,row] = sum_seg A_seg[pix+(dh,dw), :] . bmat[row, kslice]."""
    N = a_views[0].shape[0]
    D = torch.zeros(N, OH, OW, bmat.shape[0], dtype=bmat.dtype)
    k0 = 0
    for m, dh, dw, chunks in segs:
        A = shifted(a_views[m], dh, dw, OH, OW, chunks * 64)
        D += A @ bmat[:, k0:k0 + chunks * 64].T
        k0 += chunks * 64
    assert k0 == bmat.shape[1]
    return D


def emu_wgrad(p_views, q, segs, PC, QC):
    """G[seg, p, q] = sum_pix P_seg[pix+(dh,dw), p] * Q[pix, q]."""
    N, OH, OW, _ = q.shape
    G = torch.zeros(len(segs), PC, QC, dtype=q.dtype)
    for i, (m, dh, dw) in enumerate(segs):
        P = shifted(p_views[m], dh, dw, OH, OW, PC)
        G[i] = torch.einsum("nhwp,nhwq->pq", P, q[..., :QC])
    return G


def stage_rows(w_hwio, scale, idx, rows_out=None):
    """urso_stage_weight_rows: out[row,k] = w[idx[k],row]*scale[row]."""
    R = w_hwio.shape[0] * w_hwio.shape[1] * w_hwio.shape[2]
    CO = w_hwio.shape[3]
    w2 = w_hwio.reshape(R, CO)
    rows_out = rows_out or CO
    out = torch.zeros(rows_out, len(idx), dtype=w_hwio.dtype)
    idx_t = torch.tensor(idx)
    ok = idx_t >= 0
    out[:CO, ok] = (w2[idx_t[ok], :] * (scale if scale is not None else 1.0)).T
    return out


def stage_cols(w_hwio, scale, tap_map, rows_out=None):
    """urso_stage_weight_cols: out[ci, slot*COp+co] = w[tap,ci,co]*scale[co]."""
    kh, kw, CI, CO = w_hwio.shape
    cop = (CO + 63) // 64 * 64
    rows_out = rows_out or CI
    w3 = w_hwio.reshape(kh * kw, CI, CO) * (scale if scale is not None else 1.0)
    out = torch.zeros(rows_out, len(tap_map) * cop, dtype=w_hwio.dtype)
    for slot, t in enumerate(tap_map):
        if t >= 0:
            out[:CI, slot * cop:slot * cop + CO] = w3[t]
    return out


def stem_stage(img, mean=None):
    """urso_stem_stage: E[b,h2,wo,s2*16+ph*8+pw*4+c] = img[2*h2+ph-3, 2*(wo+s2)+pw-3, c] - mean (0 outside / c==3)."""
    B, H, W, _ = img.shape
    x = img.clone()
    if mean is not None:
        x = x - mean
    xp = torch.zeros(B, H + 8, W + 8, 4, dtype=img.dtype)      # pad 3 before, 5 after (taps up to index 7)
    xp[:, 3:3 + H, 3:3 + W, :3] = x
    H2, WO = H // 2 + 3, W // 2
    E = torch.zeros(B, H2, WO, 64, dtype=img.dtype)
    for s2 in range(4):
        for ph in range(2):
            for pw in range(2):
                k = s2 * 16 + ph * 8 + pw * 4
                rows = xp[:, ph:ph + 2 * H2:2]
                E[..., k:k + 4] = rows[:, :H2, 2 * s2 + pw:2 * s2 + pw + 2 * WO:2][:, :, :WO]
    return E
