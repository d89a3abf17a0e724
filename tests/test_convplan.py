"""Host logic of the implicit-GEMM engines (segments, phase views, weight staging, stem packing) checked on CPU
against the oracle's conv2d by emulating the engines' documented semantics."""
import pytest
import torch

from oracle import ursonet_oracle as O
from tests import emulator as E
from ursonet_b200 import convplan as P

torch.manual_seed(0)
DT = torch.float64

CASES = [  # kh, stride, padding, cin, cout, h, w
    (1, 1, "valid", 64, 64, 8, 12),
    (1, 2, "valid", 128, 64, 8, 12),        # conv_block 2a / shortcut: strided 1x1 samples 0,2,4..
    (3, 1, "same", 64, 128, 8, 12),         # deep 2b
    (3, 1, 1, 64, 64, 6, 10),               # shallow conv1/conv2: ZeroPadding2D(1)+valid
    (3, 2, 1, 64, 128, 8, 12),              # shallow stage entry: symmetric pad, stride 2
    (3, 2, "same", 128, 32, 8, 12),         # bottleneck_layer: TF SAME on even maps pads bottom/right only
    (3, 2, "same", 64, 32, 7, 9),           # odd maps: SAME pads 1 each side
]


@pytest.mark.parametrize("kh,stride,padding,cin,cout,h,w", CASES)
def test_forward_segments_match_oracle_conv(kh, stride, padding, cin, cout, h, w):
    x = torch.randn(2, h, w, cin, dtype=DT)
    wk = torch.randn(kh, kh, cin, cout, dtype=DT)
    scale = torch.rand(cout, dtype=DT) + 0.5
    g = P.make_geom(kh, stride, padding, cin, cout, h, w)
    segs, idx = P.fwd_segments(g)
    bmat = E.stage_rows(wk, scale, idx)
    D = E.emu_convgemm(P.input_views(x, stride), bmat, segs, g.oh, g.ow)
    ref = O.conv2d(x, wk, None, stride, padding) * scale
    assert D.shape == ref.shape
    assert torch.allclose(D, ref, atol=1e-10)


@pytest.mark.parametrize("kh,stride,padding,cin,cout,h,w", CASES)
def test_dgrad_phases_match_autograd(kh, stride, padding, cin, cout, h, w):
    x = torch.randn(2, h, w, cin, dtype=DT, requires_grad=True)
    wk = torch.randn(kh, kh, cin, cout, dtype=DT)
    scale = torch.rand(cout, dtype=DT) + 0.5
    y = O.conv2d(x, wk, None, stride, padding) * scale
    du = torch.randn_like(y)
    (ref,) = torch.autograd.grad(y, x, du)
    g = P.make_geom(kh, stride, padding, cin, cout, h, w)
    dx = torch.zeros_like(ref)
    cop = P.ceil64(cout)
    du_p = torch.zeros(*du.shape[:3], cop, dtype=DT)
    du_p[..., :cout] = du
    for oph, opw, segs, tap_map in P.dgrad_phases(g):
        tgt = dx[:, oph::stride, opw::stride, :]
        if not segs:
            continue
        bmat = E.stage_cols(wk, scale, tap_map)
        tgt.copy_(E.emu_convgemm([du_p], bmat, segs, tgt.shape[1], tgt.shape[2]))
    assert torch.allclose(dx, ref, atol=1e-10)


@pytest.mark.parametrize("kh,stride,padding,cin,cout,h,w", CASES)
def test_wgrad_segments_match_autograd(kh, stride, padding, cin, cout, h, w):
    x = torch.randn(2, h, w, cin, dtype=DT)
    wk = torch.randn(kh, kh, cin, cout, dtype=DT, requires_grad=True)
    y = O.conv2d(x, wk, None, stride, padding)
    du = torch.randn_like(y)
    (ref,) = torch.autograd.grad(y, wk, du)
    g = P.make_geom(kh, stride, padding, cin, cout, h, w)
    G = E.emu_wgrad(P.input_views(x, stride), du, P.wgrad_segments(g), cin, cout)
    assert torch.allclose(G.reshape(kh, kh, cin, cout), ref, atol=1e-9)


@pytest.mark.parametrize("kh,padding,cin,cout,h,w", [(1, "valid", 64, 128, 8, 12), (3, "same", 64, 64, 8, 12),
                                                     (3, 1, 128, 64, 6, 10)])
def test_sparse_output_gradient_uses_decimated_geometry(kh, padding, cin, cout, h, w):
    """Backward of a stride-1 conv whose output gradient lives on the even-even pixels only (engine.py, structural
    sparsity behind 1x1/stride-2 convs): dgrad / wgrad on the decimated grid with the stride-2 geometry equal autograd
    with the zero-filled dense gradient; for a 1x1 conv only phase (0,0) of dx is written (it is sparse again)."""
    x = torch.randn(2, h, w, cin, dtype=DT, requires_grad=True)
    wk = torch.randn(kh, kh, cin, cout, dtype=DT, requires_grad=True)
    scale = torch.rand(cout, dtype=DT) + 0.5
    y = O.conv2d(x, wk, None, 1, padding)
    du = torch.zeros_like(y)
    du[:, ::2, ::2, :] = torch.randn_like(du[:, ::2, ::2, :])          # what a 1x1/s2 consumer sends back
    ref_dx, ref_dw = torch.autograd.grad(y * scale, (x, wk), du)
    g = P.decimated_geom(P.make_geom(kh, 1, padding, cin, cout, h, w))
    du_dec = du[:, ::2, ::2, :]
    assert (g.oh, g.ow) == tuple(du_dec.shape[1:3])
    # dgrad: four phase launches over the decimated gradient
    dx = torch.zeros_like(ref_dx)
    du_p = torch.zeros(*du_dec.shape[:3], P.ceil64(cout), dtype=DT)
    du_p[..., :cout] = du_dec
    written = []
    for oph, opw, segs, tap_map in P.dgrad_phases(g):
        tgt = dx[:, oph::2, opw::2, :]
        if not segs:
            continue
        written.append((oph, opw))
        tgt.copy_(E.emu_convgemm([du_p], E.stage_cols(wk.detach(), scale, tap_map), segs, tgt.shape[1], tgt.shape[2]))
    assert torch.allclose(dx, ref_dx, atol=1e-10)
    assert written == ([(0, 0)] if kh == 1 else [(0, 0), (0, 1), (1, 0), (1, 1)])
    # wgrad: the four parity views of x against the decimated gradient
    G = E.emu_wgrad(P.input_views(x.detach(), 2), du_dec, P.wgrad_segments(g), cin, cout)
    assert torch.allclose(G.reshape(kh, kh, cin, cout) * scale, ref_dw, atol=1e-9)


def test_stem_space_to_depth_matches_7x7_s2():
    B, H, W = 2, 16, 24
    img = torch.rand(B, H, W, 3, dtype=DT) * 255
    mean = torch.tensor(O.MEAN_PIXEL, dtype=DT)
    wk = torch.randn(7, 7, 3, 64, dtype=DT)
    ref = O.conv2d(img - mean, wk, None, 2, 3)             # ZeroPadding2D(3) + 7x7/s2 valid (net.py:170-171)
    Est = E.stem_stage(img, mean)
    assert Est.shape == (B, H // 2 + 3, W // 2, 64)
    bmat = E.stage_rows(wk, None, P.stem_weight_index(3))
    assert bmat.shape == (64, P.STEM_K)
    D = E.emu_convgemm([Est], bmat, P.stem_segments(), H // 2, W // 2)
    assert torch.allclose(D, ref, atol=1e-8)
    # wgrad of the staged form maps back onto the 7x7x3 kernel through stem_grad_row_map
    du = torch.randn_like(ref)
    wk2 = wk.clone().requires_grad_(True)
    (gref,) = torch.autograd.grad(O.conv2d(img - mean, wk2, None, 2, 3), wk2, du)
    G = E.emu_wgrad([Est], du, [(m, dh, dw) for m, dh, dw, _ in P.stem_segments()], 64, 64).reshape(256, 64)
    rows = torch.tensor(P.stem_grad_row_map(3))
    assert torch.allclose(G[rows].reshape(7, 7, 3, 64), gref, atol=1e-8)


def test_pick_patch():
    assert P.pick_patch(160, 240, 128) == (16, 8) or P.pick_patch(160, 240, 128)[0] * P.pick_patch(160, 240, 128)[1] == 128
    for oh, ow in [(160, 240), (80, 120), (40, 60), (20, 30), (10, 15), (1, 4800)]:
        for npix in (64, 128):
            tw, th = P.pick_patch(oh, ow, npix)
            assert tw * th == npix and tw <= 256 and th <= 256
            cover = -(-ow // tw) * tw * (-(-oh // th) * th)
            assert cover <= 1.35 * oh * ow or oh * ow < npix * 4
    assert P.pick_patch(1, 100000, 128) == (128, 1)


def test_same_pad_rule():
    assert P.same_pad(320, 3, 2) == (0, 1)     # even map: bottom/right only
    assert P.same_pad(7, 3, 2) == (1, 1)
    assert P.same_pad(20, 3, 1) == (1, 1)
    assert O.same_pad(320, 3, 2) == (0, 1)
