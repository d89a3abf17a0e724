"""GPU parity of the Conv2D OPERATORS of the C-ABI (urso_conv2d_{fwd,dgrad,wgrad}_*: include/urso_b200.h): the level a
non-Python host binds.  Each case passes only shapes, stride, padding and device pointers through ctypes -- all
planning (segments, parity views, tiles, operand layout) happens in csrc/conv_ops.cu -- and is compared with the fp64
oracle convolution / its autograd on bf16-exact inputs.  Tolerance: fp32 accumulation order + bf16 rounding of the
folded weights (exact in the reference here, see `staged`) and of the stored output (2^-8 relative)."""
import pytest
import torch

from oracle import ursonet_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(params=[0, 3], autouse=True, ids=["all_sms", "3_ctas"])
def cta_limit(request):
    """Every case also runs with the persistent grids limited to 3 CTAs (urso_set_max_ctas): each CTA then works through a
    long queue of tiles, so the two operand pipelines, their ring wrap-arounds and the accumulator-stage phase flips are
    exercised even at these small shapes."""
    from ursonet_b200 import lib
    lib.load().urso_set_max_ctas(request.param)
    yield
    lib.load().urso_set_max_ctas(0)


@pytest.fixture(params=[1, 0], ids=["addend_mma", "addend_epilogue"])
def residual_mma(request):
    """Launches with an addend run both ways: accumulated on the tensor core as an extra identity K step (default), or
    loaded and added by the epilogue warps (urso_set_residual_mma)."""
    from ursonet_b200 import lib
    lib.load().urso_set_residual_mma(request.param)
    yield request.param
    lib.load().urso_set_residual_mma(1)


def bf16_exact(*shape, scale=1.0, seed=0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(torch.bfloat16).to(torch.float64)


def staged(wk, scale):
    """the bf16-rounded, scale-folded kernel the engine multiplies with"""
    return (wk * scale).to(torch.bfloat16).to(torch.float64)


def relerr(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


CASES = [  # k, stride, padding, cin, cout, h, w
    (1, 1, "valid", 64, 64, 16, 24),
    (1, 1, "valid", 256, 128, 12, 20),
    (1, 1, "valid", 64, 256, 16, 24),
    (1, 2, "valid", 128, 64, 16, 24),        # Keras-v1 block: stride on the 1x1 (net.py:138,152)
    (3, 1, "same", 64, 64, 16, 24),
    (3, 1, "same", 128, 128, 20, 30),
    (3, 2, "same", 128, 32, 20, 30),         # bottleneck_layer: TF SAME on an even map pads bottom/right only (net.py:639)
    (3, 2, 1, 64, 128, 16, 24),              # shallow block: ZeroPadding2D(1) + VALID stride 2 (net.py:225-226)
    (3, 1, 1, 64, 64, 10, 14),               # shallow block conv2
]


@pytest.mark.parametrize("k,stride,padding,cin,cout,h,w", CASES)
def test_conv2d_fwd_operator(k, stride, padding, cin, cout, h, w, residual_mma):
    from ursonet_b200 import lib
    N = 2
    x = bf16_exact(N, h, w, cin, seed=1)
    wk = bf16_exact(k, k, cin, cout, scale=0.05, seed=2)
    scale = 0.5 + torch.rand(cout, dtype=torch.float64)
    shift = torch.randn(cout, dtype=torch.float64)
    shape = lib.conv_shape(N, h, w, cin, cout, k, stride, padding)
    oh, ow = lib.out_hw(shape)
    out_fp32 = cout % 64 != 0
    addend = bf16_exact(N, oh, ow, cout, seed=3) if (k == 1 and not out_fp32) else None
    relu = k != 3 or stride == 1
    xd = x.to(torch.bfloat16).to(DEV)
    wd, sd, hd = wk.float().to(DEV), scale.float().to(DEV), shift.float().to(DEV)
    ad = addend.to(torch.bfloat16).to(DEV) if addend is not None else None
    y = torch.full((N, oh, ow, cout), float("nan"), dtype=torch.float32 if out_fp32 else torch.bfloat16, device=DEV)
    op = lib.Conv2dFwd(shape, xd, wd, sd, hd, y, addend=ad, relu=relu and not out_fp32)
    op.stage()
    op.launch()
    torch.cuda.synchronize()
    ref = O.conv2d(x, staged(wk, scale.float().double()), None, stride, padding) + shift.float().double()
    assert tuple(ref.shape) == (N, oh, ow, cout)
    if addend is not None:
        ref = ref + addend
    if relu and not out_fp32:
        ref = torch.relu(ref)
    assert relerr(y.double().cpu(), ref) <= (1e-5 if out_fp32 else 6e-3)


def test_conv2d_fwd_stem_operator():
    """ksize 7: the stem over the staged tensor E of urso_stem_stage (net.py:170-171: ZeroPadding2D(3) + 7x7/s2 VALID)."""
    from ursonet_b200 import lib
    N, H, W = 2, 32, 48
    g = torch.Generator().manual_seed(4)
    img = torch.randint(0, 256, (N, H, W, 3), generator=g, dtype=torch.uint8)
    mean = torch.tensor(O.MEAN_PIXEL, dtype=torch.float32)
    wk = bf16_exact(7, 7, 3, 64, scale=0.05, seed=5)
    scale = 0.5 + torch.rand(64, dtype=torch.float64)
    shift = torch.randn(64, dtype=torch.float64)
    E = torch.zeros(N, H // 2 + 3, W // 2 + 3, 16, dtype=torch.bfloat16, device=DEV)      # compact staged tensor
    imgd = img.to(DEV)
    lib.call("urso_stem_stage", imgd.data_ptr(), 1, 1, mean.to(DEV).data_ptr(), E.data_ptr(), N, H, W, 0, lib.stream_ptr())
    shape = lib.conv_shape(N, H, W, 3, 64, 7, 2, 3)
    y = torch.zeros(N, H // 2, W // 2, 64, dtype=torch.bfloat16, device=DEV)
    op = lib.Conv2dFwd(shape, E, wk.float().to(DEV), scale.float().to(DEV), shift.float().to(DEV), y, relu=True)
    op.stage()
    op.launch()
    torch.cuda.synchronize()
    x = (img.double() - mean.double()).to(torch.bfloat16).double()
    ref = torch.relu(O.conv2d(x, staged(wk, scale.float().double()), None, 2, 3) + shift.float().double())
    assert relerr(y.double().cpu(), ref) <= 6e-3
    # weight gradient of the stem through the operator + the row map
    dy = bf16_exact(N, H // 2, W // 2, 64, seed=6)
    G = torch.zeros(4 * 64, 64, dtype=torch.float32, device=DEV)
    wg = lib.Conv2dWgrad(shape, E, dy.to(torch.bfloat16).to(DEV), G)
    wg.launch()
    torch.cuda.synchronize()
    wr = wk.clone().requires_grad_(True)
    (gref,) = torch.autograd.grad(O.conv2d(x, wr, None, 2, 3), wr, dy)
    got = G.double().cpu()[torch.tensor(lib.stem_grad_row_map())].reshape(7, 7, 3, 64)
    assert relerr(got, gref) <= 2e-3


DGRAD_CASES = [  # list of consumers (k, stride, padding, cout), cin, h, w, sparse, mask, addend
    ([(1, 1, "valid", 64)], 256, 16, 24, False, True, True),             # identity block 2a + shortcut fan-in
    ([(3, 1, "same", 64)], 64, 16, 24, False, True, False),
    ([(1, 2, "valid", 128), (1, 2, "valid", 512)], 256, 16, 24, False, True, False),   # conv block: 2a + shortcut, stride 2
    ([(3, 2, "same", 32)], 128, 20, 30, False, True, False),             # bottleneck conv
    ([(3, 2, 1, 128), (1, 2, "valid", 128)], 64, 16, 24, False, True, False),          # shallow block: conv1 + 'post' shortcut
    ([(1, 1, "valid", 256)], 64, 16, 24, True, True, False),             # sparse gradient behind a 1x1/s2
    ([(3, 1, "same", 64)], 64, 16, 24, True, True, False),               # sparse gradient through a 3x3
    ([(1, 1, "valid", 64), (1, 1, "valid", 256)], 64, 12, 20, False, False, False),    # stage-2 conv block on pool1
]


@pytest.mark.parametrize("consumers,cin,h,w,sparse,use_mask,use_addend", DGRAD_CASES)
def test_conv2d_dgrad_operator(consumers, cin, h, w, sparse, use_mask, use_addend, residual_mma):
    from ursonet_b200 import lib
    N = 2
    shapes, dys, ws, scs, ref_terms = [], [], [], [], []
    x = bf16_exact(N, h, w, cin, seed=10)
    xr = x.clone().requires_grad_(True)
    total = torch.zeros_like(x)
    for i, (k, stride, padding, cout) in enumerate(consumers):
        shape = lib.conv_shape(N, h, w, cin, cout, k, stride, padding)
        oh, ow = lib.out_hw(shape)
        wk = bf16_exact(k, k, cin, cout, scale=0.05, seed=20 + i)
        scale = 0.5 + torch.rand(cout, dtype=torch.float64)
        dy = bf16_exact(N, oh, ow, cout, seed=30 + i)
        if sparse:
            keep = torch.zeros(1, oh, ow, 1, dtype=torch.float64)
            keep[:, ::2, ::2, :] = 1
            dy = dy * keep
        y = O.conv2d(xr, staged(wk, scale.float().double()), None, stride, padding)
        total = total + torch.autograd.grad(y, xr, dy)[0]
        cp = (cout + 63) // 64 * 64
        dyp = torch.zeros(N, oh, ow, cp, dtype=torch.bfloat16)
        dyp[..., :cout] = dy.to(torch.bfloat16)
        shapes.append(shape)
        dys.append(dyp.to(DEV))
        ws.append(wk.float().to(DEV))
        scs.append(scale.float().to(DEV))
    mask = bf16_exact(N, h, w, cin, seed=40) if use_mask else None
    addend = bf16_exact(N, h, w, cin, seed=41) if use_addend else None
    if addend is not None:
        total = total + addend
    if mask is not None:
        total = total * (mask > 0)
    dx = torch.zeros(N, h, w, cin, dtype=torch.bfloat16, device=DEV)       # zero-filled once (untouched phases)
    cs = torch.zeros(cin, dtype=torch.float32, device=DEV)
    op = lib.Conv2dDgrad(shapes, dys, ws, scs, dx, mask=mask.to(torch.bfloat16).to(DEV) if use_mask else None,
                         addend=addend.to(torch.bfloat16).to(DEV) if use_addend else None, colsum=cs, dy_sparse=sparse)
    op.stage()
    op.launch()
    torch.cuda.synchronize()
    got = dx.double().cpu()
    assert relerr(got, total) <= 6e-3, (op.n_launches, op.untouched)
    assert torch.allclose(cs.double().cpu(), got.sum((0, 1, 2)), rtol=1e-3, atol=1e-2 * got.abs().max().item())
    eff_stride = consumers[0][1] * (2 if sparse else 1)
    if eff_stride == 2 and all(c[0] == 1 for c in consumers):
        assert op.untouched == 0b1110 and op.n_launches == 1
    else:
        assert op.untouched == 0 and op.n_launches == eff_stride ** 2


WGRAD_CASES = [  # k, stride, padding, cin, cout, h, w, sparse
    (1, 1, "valid", 64, 256, 16, 24, False),      # swapped roles (wide side on M)
    (1, 1, "valid", 256, 64, 16, 24, False),
    (1, 2, "valid", 128, 64, 16, 24, False),
    (3, 1, "same", 64, 64, 16, 24, False),        # tap pairing
    (3, 1, "same", 128, 128, 12, 20, False),
    (3, 2, "same", 128, 32, 20, 30, False),
    (3, 2, 1, 64, 128, 16, 24, False),
    (1, 1, "valid", 64, 256, 16, 24, True),
    (3, 1, "same", 64, 64, 16, 24, True),
    (3, 1, "same", 128, 128, 16, 24, False),      # halo mode, two 64-channel atoms per box (LBO = box pitch), 3 tap groups
    (3, 1, "same", 128, 256, 24, 40, False),      # halo mode, block_q 256 (2 taps per CTA): not taken (boxes >= 70 % of the atoms)
    (3, 1, "same", 64, 64, 23, 31, False),        # halo mode on a map that 8 x 8 blocks do not tile (clipped boxes)
    (3, 1, 1, 64, 64, 10, 14, False),             # explicit padding 1 (shallow block conv2)
]


@pytest.fixture(params=[1, 0], ids=["wgrad_halo", "wgrad_atom_per_tap"])
def wgrad_halo(request):
    """Engine W with and without its halo mode (urso_set_wgrad_halo): one box + halo per K step for all taps of a CTA, or
    one operand atom per tap."""
    from ursonet_b200 import lib
    lib.load().urso_set_wgrad_halo(request.param)
    yield request.param
    lib.load().urso_set_wgrad_halo(1)


@pytest.mark.parametrize("k,stride,padding,cin,cout,h,w,sparse", WGRAD_CASES)
def test_conv2d_wgrad_operator(k, stride, padding, cin, cout, h, w, sparse, wgrad_halo):
    from ursonet_b200 import lib
    N = 2
    shape = lib.conv_shape(N, h, w, cin, cout, k, stride, padding)
    oh, ow = lib.out_hw(shape)
    x = bf16_exact(N, h, w, cin, seed=50)
    dy = bf16_exact(N, oh, ow, cout, seed=51)
    if sparse:
        keep = torch.zeros(1, oh, ow, 1, dtype=torch.float64)
        keep[:, ::2, ::2, :] = 1
        dy = dy * keep
    wr = torch.zeros(k, k, cin, cout, dtype=torch.float64, requires_grad=True)
    (gref,) = torch.autograd.grad(O.conv2d(x, wr, None, stride, padding), wr, dy)
    cp = (cout + 63) // 64 * 64
    dyp = torch.zeros(N, oh, ow, cp, dtype=torch.bfloat16)
    dyp[..., :cout] = dy.to(torch.bfloat16)
    G = torch.zeros(k * k * cin, cout, dtype=torch.float32, device=DEV)
    op = lib.Conv2dWgrad(shape, x.to(torch.bfloat16).to(DEV), dyp.to(DEV), G, dy_sparse=sparse)
    op.launch()
    torch.cuda.synchronize()
    assert relerr(G.double().cpu().reshape(k, k, cin, cout), gref) <= 2e-3


def test_operator_errors_are_reported():
    from ursonet_b200 import lib
    shape = lib.conv_shape(1, 16, 16, 48, 64, 3, 1, "same")         # 48 input channels: not a multiple of 64
    x = torch.zeros(1, 16, 16, 48, dtype=torch.bfloat16, device=DEV)
    y = torch.zeros(1, 16, 16, 64, dtype=torch.bfloat16, device=DEV)
    w = torch.zeros(3, 3, 48, 64, device=DEV)
    with pytest.raises(lib.UrsoError):
        lib.Conv2dFwd(shape, x, w, None, None, y)
    assert b"multiple of 64" in lib.load().urso_last_error()


# ------------------------------------------------------------------------------------------------ pipelines at depth
# Shapes with MANY tiles per CTA (>= 8 on 148 SMs), so that both operand pipelines wrap their shared-memory rings and
# reuse each of their two accumulator stages several times (barrier phase flips), in every planning mode:
# halo + resident weights (N = 64), halo + streamed weights (N = 128, two channel chunks), stream mode with two K steps
# per barrier round, one pipeline with BLOCK_N = 256.
BIG_CASES = [  # k, stride, padding, cin, cout, N, h, w, expect (subset of plan_info)
    (3, 1, "same", 64, 64, 4, 160, 240, dict(halo=1, bres=1, npipe=2)),
    (3, 1, "same", 128, 128, 16, 80, 120, dict(halo=1, bres=0, npipe=2)),
    (1, 1, "valid", 256, 64, 4, 160, 240, dict(halo=0, npipe=2)),
    (1, 1, "valid", 512, 128, 8, 80, 120, dict(halo=0, npipe=2)),
    (3, 1, "same", 256, 256, 8, 40, 60, dict(halo=0, npipe=1, block_n=256)),
    (1, 2, "valid", 256, 128, 8, 160, 240, dict(halo=0, npipe=2)),
]


@pytest.mark.parametrize("k,stride,padding,cin,cout,N,h,w,expect", BIG_CASES)
def test_conv2d_fwd_many_tiles_per_cta(k, stride, padding, cin, cout, N, h, w, expect):
    from ursonet_b200 import lib
    x = bf16_exact(N, h, w, cin, seed=61)
    wk = bf16_exact(k, k, cin, cout, scale=0.05, seed=62)
    scale = 0.5 + torch.rand(cout, dtype=torch.float64)
    shift = torch.randn(cout, dtype=torch.float64)
    shape = lib.conv_shape(N, h, w, cin, cout, k, stride, padding)
    oh, ow = lib.out_hw(shape)
    y = torch.full((N, oh, ow, cout), float("nan"), dtype=torch.bfloat16, device=DEV)
    op = lib.Conv2dFwd(shape, x.to(torch.bfloat16).to(DEV), wk.float().to(DEV), scale.float().to(DEV),
                       shift.float().to(DEV), y, relu=True)
    info = op.plan_info()
    for key, val in expect.items():
        assert info[key] == val, (key, info)
    op.stage()
    op.launch()
    op.launch()          # a second launch must not depend on state left by the first
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = torch.relu(O.conv2d(x, staged(wk, scale.float().double()), None, stride, padding) + shift.float().double())
    assert relerr(y.double().cpu(), ref) <= 6e-3, info


def test_conv2d_dgrad_halo_resident_with_mask_many_tiles():
    """The stage-2 3x3 input gradient of the bench workload (64 channels, 160x240): halo + resident weights + two
    pipelines + the TMA-prefetched ReLU mask, 8 tiles per CTA."""
    from ursonet_b200 import lib
    N, h, w, c = 4, 160, 240, 64
    shape = lib.conv_shape(N, h, w, c, c, 3, 1, "same")
    wk = bf16_exact(3, 3, c, c, scale=0.05, seed=71)
    scale = 0.5 + torch.rand(c, dtype=torch.float64)
    dy = bf16_exact(N, h, w, c, seed=72)
    mask = bf16_exact(N, h, w, c, seed=73)
    xr = torch.zeros(N, h, w, c, dtype=torch.float64, requires_grad=True)
    (gref,) = torch.autograd.grad(O.conv2d(xr, staged(wk, scale.float().double()), None, 1, "same"), xr, dy)
    gref = gref * (mask > 0)
    dx = torch.zeros(N, h, w, c, dtype=torch.bfloat16, device=DEV)
    cs = torch.zeros(c, dtype=torch.float32, device=DEV)
    op = lib.Conv2dDgrad([shape], [dy.to(torch.bfloat16).to(DEV)], [wk.float().to(DEV)], [scale.float().to(DEV)], dx,
                         mask=mask.to(torch.bfloat16).to(DEV), colsum=cs)
    op.stage()
    op.launch()
    torch.cuda.synchronize()
    got = dx.double().cpu()
    assert relerr(got, gref) <= 6e-3
    assert torch.allclose(cs.double().cpu(), got.sum((0, 1, 2)), rtol=1e-3, atol=1e-2 * got.abs().max().item())


# ------------------------------------------------------------------------------------------------ bit-packed ReLU masks
def unpack_bits(bits, C):
    """int32 [N,H,W,C/32] -> bool [N,H,W,C] following the documented order (include/urso_b200.h, urso_convgemm_desc):
    channel 32g + c  <->  bit (7 - (c >> 2)) + 8 (c & 3) of word g."""
    b = bits.cpu().to(torch.int64) & 0xFFFFFFFF
    out = torch.zeros(*bits.shape[:3], C, dtype=torch.bool)
    for g in range(C // 32):
        for c in range(32):
            out[..., 32 * g + c] = ((b[..., g] >> ((7 - (c >> 2)) + 8 * (c & 3))) & 1).bool()
    return out


@pytest.mark.parametrize("k,stride,padding,cin,cout,N,h,w", [
    (1, 1, "valid", 64, 256, 2, 16, 24),        # flat pixels, 4 chunks per tile
    (3, 1, "same", 64, 64, 2, 32, 24),          # halo patch
    (3, 1, "same", 128, 128, 2, 20, 30),        # partial tiles
    (1, 2, "valid", 128, 64, 2, 16, 24),
    (3, 1, "same", 64, 64, 4, 160, 240),        # two pipelines, many tiles per CTA
])
def test_relu_bits_written_by_fwd_and_consumed_by_dgrad(k, stride, padding, cin, cout, N, h, w):
    from ursonet_b200 import lib
    x = bf16_exact(N, h, w, cin, seed=81)
    wk = bf16_exact(k, k, cin, cout, scale=0.05, seed=82)
    scale = 0.5 + torch.rand(cout, dtype=torch.float64)
    shift = torch.randn(cout, dtype=torch.float64) * 0.3
    shape = lib.conv_shape(N, h, w, cin, cout, k, stride, padding)
    oh, ow = lib.out_hw(shape)
    y = torch.full((N, oh, ow, cout), float("nan"), dtype=torch.bfloat16, device=DEV)
    bits = torch.full((N, oh, ow, cout // 32), -1, dtype=torch.int32, device=DEV)
    op = lib.Conv2dFwd(shape, x.to(torch.bfloat16).to(DEV), wk.float().to(DEV), scale.float().to(DEV),
                       shift.float().to(DEV), y, relu=True, relu_bits=bits)
    op.stage()
    op.launch()
    torch.cuda.synchronize()
    # the bits are exactly (stored output > 0)
    assert torch.equal(unpack_bits(bits, cout), (y.float().cpu() > 0))
    frac = (y.float() > 0).float().mean().item()
    assert 0.2 < frac < 0.8          # a meaningful mask
    # dgrad of a 3x3 consumer of y with the bit mask == dgrad with the bf16 activation as the mask (bit-exact)
    c2 = 64
    shape2 = lib.conv_shape(N, oh, ow, cout, c2, 3, 1, "same")
    w2 = bf16_exact(3, 3, cout, c2, scale=0.05, seed=83).float().to(DEV)
    s2 = (0.5 + torch.rand(c2)).to(DEV)
    dy = bf16_exact(N, oh, ow, c2, seed=84).to(torch.bfloat16).to(DEV)
    outs = []
    for use_bits in (False, True):
        dx = torch.zeros(N, oh, ow, cout, dtype=torch.bfloat16, device=DEV)
        cs = torch.zeros(cout, dtype=torch.float32, device=DEV)
        dop = lib.Conv2dDgrad([shape2], [dy], [w2], [s2], dx, mask=None if use_bits else y,
                              mask_bits=bits if use_bits else None, colsum=cs)
        dop.stage()
        dop.launch()
        torch.cuda.synchronize()
        outs.append((dx.clone(), cs.clone()))
    assert torch.equal(outs[0][0], outs[1][0])
    assert torch.allclose(outs[0][1], outs[1][1], rtol=1e-4, atol=1e-3)
    assert (outs[1][0].float().abs().sum() > 0)


def test_mask_bits_on_a_strided_dgrad_phase_grid():
    """Stride-2 consumers write dx one parity phase at a time: the bit mask is then addressed through the phase-strided
    pixel grid (conv block: 1x1/s2 '2a' + 1x1/s2 shortcut on a 256-channel block output)."""
    from ursonet_b200 import lib
    N, h, w, cin = 2, 16, 24, 256
    xact = bf16_exact(N, h, w, cin, seed=91)
    mask_ref = xact > 0
    # build the bits from the documented order on the host
    words = torch.zeros(N, h, w, cin // 32, dtype=torch.int64)
    for g in range(cin // 32):
        for c in range(32):
            words[..., g] |= mask_ref[..., 32 * g + c].to(torch.int64) << ((7 - (c >> 2)) + 8 * (c & 3))
    words = torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32).to(DEV)
    shapes, dys, ws, scs = [], [], [], []
    for i, cout in enumerate((128, 512)):
        sh = lib.conv_shape(N, h, w, cin, cout, 1, 2, "valid")
        oh, ow = lib.out_hw(sh)
        shapes.append(sh)
        dys.append(bf16_exact(N, oh, ow, cout, seed=92 + i).to(torch.bfloat16).to(DEV))
        ws.append(bf16_exact(1, 1, cin, cout, scale=0.05, seed=94 + i).float().to(DEV))
        scs.append((0.5 + torch.rand(cout)).to(DEV))
    outs = []
    for use_bits in (False, True):
        dx = torch.zeros(N, h, w, cin, dtype=torch.bfloat16, device=DEV)
        dop = lib.Conv2dDgrad(shapes, dys, ws, scs, dx, mask=None if use_bits else xact.to(torch.bfloat16).to(DEV),
                              mask_bits=words if use_bits else None)
        dop.stage()
        dop.launch()
        torch.cuda.synchronize()
        outs.append(dx.clone())
    assert torch.equal(outs[0], outs[1]) and outs[0].float().abs().sum() > 0


# ------------------------------------------------------------------------------------------------ N-split tail
TAIL_CASES = [  # k, cin, cout, N, h, w, max_ctas (0 = all SMs), addend
    (1, 512, 256, 4, 80, 120, 0, True),       # 300 tiles on 148 CTAs: 4 tail tiles x 4 sub-tiles; residual through the epilogue ring
    (3, 256, 256, 8, 40, 60, 0, False),       # the stage-4 3x3 of the bench workload in small: 150 tiles, 2 tail tiles x 4
    (1, 1024, 512, 8, 40, 60, 0, False),      # two N tiles per pixel tile: the sub-tiles of tail tiles with n_tile = 0 and 1
    (1, 512, 256, 4, 80, 120, 13, False),     # 300 % 13 = 1 tail tile x 4 sub-tiles, 23 whole tiles per CTA before it
    (1, 512, 256, 4, 80, 120, 11, True),      # 300 % 11 = 3 tail tiles: 4 x 3 > 11 -> 2 sub-tiles of 128 channels each
    (1, 512, 256, 1, 37, 53, 0, False),       # 16 tiles: fewer tiles than CTAs -> no full wave, no tail split (plan check only)
]


@pytest.mark.parametrize("k,cin,cout,N,h,w,max_ctas,with_addend", TAIL_CASES)
def test_tail_split_is_bit_identical_and_matches_oracle(k, cin, cout, N, h, w, max_ctas, with_addend):
    """urso_set_tail_split(0 / 1): K-heavy BLOCK_N = 256 launches cut the tiles of their last partial wave along N into 2 or 4
    sub-tiles (one per CTA).  No reduction is involved and every output element keeps its accumulation order, so outputs
    and ReLU mask bits must be bit-identical to the whole-tile launch -- and match the fp64 oracle."""
    from ursonet_b200 import lib
    L = lib.load()
    x = bf16_exact(N, h, w, cin, seed=171)
    wk = bf16_exact(k, k, cin, cout, scale=0.03, seed=172)
    scale = 0.5 + torch.rand(cout, dtype=torch.float64)
    shift = torch.randn(cout, dtype=torch.float64)
    addend = bf16_exact(N, h, w, cout, seed=173) if with_addend else None
    shape = lib.conv_shape(N, h, w, cin, cout, k, 1, "same" if k == 3 else "valid")
    xd, wd = x.to(torch.bfloat16).to(DEV), wk.float().to(DEV)
    sd, hd = scale.float().to(DEV), shift.float().to(DEV)
    ad = addend.to(torch.bfloat16).to(DEV) if with_addend else None
    outs = []
    try:
        L.urso_set_max_ctas(max_ctas)
        for on in (0, 1):
            L.urso_set_tail_split(on)
            y = torch.full((N, h, w, cout), float("nan"), dtype=torch.bfloat16, device=DEV)
            bits = torch.full((N, h, w, cout // 32), -1, dtype=torch.int32, device=DEV)
            op = lib.Conv2dFwd(shape, xd, wd, sd, hd, y, addend=ad, relu=True, relu_bits=bits)
            op.stage()
            op.launch()
            op.launch()
            torch.cuda.synchronize()
            outs.append((y, bits, op.plan_info()))
    finally:
        L.urso_set_tail_split(1)
        L.urso_set_max_ctas(0)
    info = outs[1][2]
    tiles = -(-(N * h * w) // 128) * (cout // 256) if k == 1 else None
    assert outs[0][2]["tail_split"] == 1
    if tiles is not None:
        grid = info["grid"]
        R = tiles % grid if tiles > grid else 0
        expect = 1 if R == 0 else (4 if 4 * R <= grid else (2 if 2 * R <= grid else 1))
        assert info["tail_split"] == expect, (info, tiles)
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]), info
    with torch.no_grad():
        ref = O.conv2d(x, staged(wk, scale.float().double()), None, 1, "same" if k == 3 else "valid") + shift.float().double()
        if with_addend:
            ref = ref + addend
        ref = torch.relu(ref)
    assert relerr(outs[1][0].double().cpu(), ref) <= 6e-3, info


def test_tail_split_dgrad_with_mask_bits_and_colsum(request):
    """The input gradient of a 1x1 conv 256 -> 1024 (K = 1024, 16 K steps) with bit-packed ReLU mask and fused column sums,
    300 tiles: the tail sub-tiles address their own 64-channel slice of the mask words and of the column sums."""
    from ursonet_b200 import lib
    L = lib.load()
    N, h, w, cin, cout = 4, 80, 120, 256, 1024
    shape = lib.conv_shape(N, h, w, cin, cout, 1, 1, "valid")
    wk = bf16_exact(1, 1, cin, cout, scale=0.03, seed=181)
    scale = 0.5 + torch.rand(cout, dtype=torch.float64)
    dy = bf16_exact(N, h, w, cout, seed=182)
    act = bf16_exact(N, h, w, cin, seed=183)
    bits = torch.zeros(N, h, w, cin // 32, dtype=torch.int32)
    pos = (act > 0)
    for c in range(32):        # channel 32 g + c  <->  bit (7 - (c >> 2)) + 8 (c & 3) of word g   (see unpack_bits)
        bit = (7 - (c >> 2)) + 8 * (c & 3)
        word = pos.reshape(N, h, w, cin // 32, 32)[..., c].to(torch.int64) << bit
        bits += torch.where(word >= 2 ** 31, word - 2 ** 32, word).to(torch.int32)
    xr = torch.zeros(N, h, w, cin, dtype=torch.float64, requires_grad=True)
    (gref,) = torch.autograd.grad(O.conv2d(xr, staged(wk, scale.float().double()), None, 1, "valid"), xr, dy)
    gref = gref * pos
    outs = []
    try:
        for on in (0, 1):
            L.urso_set_tail_split(on)
            dx = torch.zeros(N, h, w, cin, dtype=torch.bfloat16, device=DEV)
            cs = torch.zeros(cin, dtype=torch.float32, device=DEV)
            op = lib.Conv2dDgrad([shape], [dy.to(torch.bfloat16).to(DEV)], [wk.float().to(DEV)], [scale.float().to(DEV)], dx,
                                 mask_bits=bits.to(DEV), colsum=cs)
            op.stage()
            op.launch()
            torch.cuda.synchronize()
            outs.append((dx, cs, L.urso_conv2d_dgrad_tail_split(op._h, 0)))
    finally:
        L.urso_set_tail_split(1)
    grid = min(L.urso_num_sms(), 300) if request.node.callspec.params["cta_limit"] == 0 else request.node.callspec.params["cta_limit"]
    R = 300 % grid
    expect = 1 if R == 0 else (4 if 4 * R <= grid else (2 if 2 * R <= grid else 1))      # 148 CTAs: 4 tail tiles x 4 sub-tiles
    assert outs[0][2] == 1 and outs[1][2] == expect, (outs[0][2], outs[1][2], grid)
    assert torch.equal(outs[0][0], outs[1][0])
    assert torch.allclose(outs[0][1], outs[1][1], rtol=1e-4, atol=1e-2)
    assert relerr(outs[1][0].double().cpu(), gref) <= 6e-3
