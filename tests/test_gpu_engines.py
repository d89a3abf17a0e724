"""GPU parity of the two tcgen05 engines (through the C-ABI) against the oracle conv / autograd on identical
bf16-representable inputs.  Accumulation is fp32 in TMEM, so with bf16-exact operands the only error is fp32
summation order: tolerance 2e-3 relative to the output scale (bf16 output rounding: 2^-8 relative, stated per test)."""
import pytest
import torch

from oracle import ursonet_oracle as O
from tests import emulator as E
from ursonet_b200 import convplan as P

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


def run_fwd(x, wk, scale, kh, stride, padding, shift=None, addend=None, mask=None, relu=False, out_fp32=False,
            colsum=False, block_n=0, flatten=False, halo=False):
    from ursonet_b200 import lib
    N, H, W, CI = x.shape
    CO = wk.shape[3]
    g = P.make_geom(kh, stride, padding, CI, CO, H, W)
    segs, idx = P.fwd_segments(g)
    bmat = E.stage_rows(wk, scale, idx).to(torch.bfloat16).to(DEV).contiguous()
    xd = x.to(torch.bfloat16).to(DEV).contiguous()
    out = torch.full((N, g.oh, g.ow, CO), float("nan"), dtype=torch.float32 if out_fp32 else torch.bfloat16, device=DEV)
    views = P.input_views(xd, stride)
    ad = addend.to(torch.bfloat16).to(DEV).contiguous() if addend is not None else None
    md = mask.to(torch.bfloat16).to(DEV).contiguous() if mask is not None else None
    sd = shift.to(torch.float32).to(DEV) if shift is not None else None
    cs = torch.zeros(CO, dtype=torch.float32, device=DEV) if colsum else None
    if flatten:
        assert kh == 1 and stride == 1
        M = N * H * W
        views = [xd.view(1, 1, M, CI)]
        o = out.view(1, 1, M, CO)
        a2 = ad.view(1, 1, M, CO) if ad is not None else None
        m2 = md.view(1, 1, M, CO) if md is not None else None
        plan = lib.ConvGemm(views, bmat, segs, o, M, 1, 1, 128, 1, shift=sd, addend=a2, mask=m2, relu=relu, colsum=cs,
                            block_n=block_n)
    else:
        tw, th = (8, 16) if halo else P.pick_patch(g.oh, g.ow, 128)
        plan = lib.ConvGemm(views, bmat, segs, out, g.ow, g.oh, N, tw, th, shift=sd, addend=ad, mask=md, relu=relu,
                            colsum=cs, block_n=block_n, halo=halo)
    plan.launch()
    torch.cuda.synchronize()
    return out.double().cpu(), (cs.double().cpu() if cs is not None else None), bmat.double().cpu()


def ref_fwd(x, wk_staged_scale, kh, stride, padding, shift=None, addend=None, mask=None, relu=False):
    y = O.conv2d(x, wk_staged_scale, None, stride, padding)
    if shift is not None:
        y = y + shift
    if addend is not None:
        y = y + addend
    if relu:
        y = torch.relu(y)
    if mask is not None:
        y = torch.where(mask > 0, y, torch.zeros_like(y))
    return y


def staged_kernel(wk, scale):
    """the bf16-rounded, scale-folded kernel the engine actually multiplies with"""
    return (wk * scale).to(torch.bfloat16).to(torch.float64)


FWD_CASES = [  # kh, stride, padding, cin, cout, h, w, flatten, block_n
    (1, 1, "valid", 64, 64, 16, 24, True, 0),
    (1, 1, "valid", 256, 128, 16, 24, True, 0),
    (1, 1, "valid", 128, 256, 16, 24, True, 0),      # BLOCK_N = 256
    (1, 1, "valid", 128, 256, 16, 24, True, 128),
    (1, 1, "valid", 64, 64, 20, 30, False, 0),       # 4-D patches with partial tiles
    (3, 1, "same", 64, 64, 16, 24, False, 0),
    (3, 1, "same", 128, 128, 20, 30, False, 0),      # partial tiles + halo
    (1, 2, "valid", 128, 64, 16, 24, False, 0),      # strided views
    (3, 2, "same", 128, 32, 20, 30, False, 0),       # bottleneck: BLOCK_N 32, phase views, SAME on even map
    (3, 2, 1, 64, 128, 16, 24, False, 0),            # shallow stage entry
    (3, 1, "same", 64, 64, 40, 60, False, 0),        # many tiles per CTA? (> 148 tiles): persistence + phases
]


@pytest.mark.parametrize("out_fp32", [True, False], ids=["fp32out_regs", "bf16out_tma"])
@pytest.mark.parametrize("kh,stride,padding,cin,cout,h,w,flatten,block_n", FWD_CASES)
def test_engine_f_forward(kh, stride, padding, cin, cout, h, w, flatten, block_n, out_fp32):
    """fp32 output exercises the register epilogue, bf16 output the TMA-store epilogue (cout >= 64)."""
    nb = 3 if h <= 20 else 8
    x = bf16_exact(nb, h, w, cin, seed=1)
    wk = bf16_exact(kh, kh, cin, cout, scale=0.05, seed=2)
    scale = torch.ones(cout, dtype=torch.float64)
    got, _, _ = run_fwd(x, wk, scale, kh, stride, padding, out_fp32=out_fp32, flatten=flatten, block_n=block_n)
    ref = ref_fwd(x, staged_kernel(wk, scale), kh, stride, padding)
    assert got.shape == ref.shape
    assert torch.isfinite(got).all(), "output has NaN: some tile was never written"
    err = (got - ref).abs().max().item()
    tol = 2e-3 if out_fp32 else 2 ** -8 + 2e-3          # bf16 output rounding
    assert err <= tol * ref.abs().max().item(), f"max abs err {err} vs scale {ref.abs().max().item()}"


@pytest.mark.parametrize("which", ["addend", "mask", "both"])
@pytest.mark.parametrize("flatten", [True, False])
def test_engine_f_tma_epilogue_inputs_many_tiles(which, flatten, residual_mma):
    """more tiles than SMs x prefetch depth, N = 256 (4 chunks/tile): exercises the per-warp input ring phases"""
    N, H, W, CI, CO = 8, 40, 64, 64, 256
    x = bf16_exact(N, H, W, CI, seed=21)
    wk = bf16_exact(1, 1, CI, CO, scale=0.05, seed=22)
    scale = torch.ones(CO, dtype=torch.float64)
    addend = bf16_exact(N, H, W, CO, seed=23) if which in ("addend", "both") else None
    mask = bf16_exact(N, H, W, CO, seed=24) if which in ("mask", "both") else None
    got, cs, _ = run_fwd(x, wk, scale, 1, 1, "valid", addend=addend, mask=mask, relu=(which != "mask"), colsum=True,
                         flatten=flatten)
    ref = ref_fwd(x, staged_kernel(wk, scale), 1, 1, "valid", None, addend, mask, which != "mask")
    assert (got - ref).abs().max().item() <= (2 ** -8 + 2e-3) * ref.abs().max().item()
    # column sums are taken over the bf16-rounded stored values
    assert torch.allclose(cs, got.sum((0, 1, 2)), rtol=2e-3, atol=1e-2 * ref.abs().max().item())


@pytest.mark.parametrize("cin,cout,h,w,out_fp32", [(64, 64, 32, 24, True), (64, 64, 32, 24, False),
                                                  (128, 128, 20, 30, False), (256, 64, 48, 40, False),
                                                  (64, 128, 160, 240, False)])
def test_engine_f_halo_mode(cin, cout, h, w, out_fp32):
    """3x3 taps served from ONE halo tile per channel chunk (row-shifted UMMA descriptors) == per-tap TMA loads."""
    nb = 2
    x = bf16_exact(nb, h, w, cin, seed=31)
    wk = bf16_exact(3, 3, cin, cout, scale=0.05, seed=32)
    scale = torch.ones(cout, dtype=torch.float64)
    mask = bf16_exact(nb, h, w, cout, seed=33) if not out_fp32 else None
    got, cs, _ = run_fwd(x, wk, scale, 3, 1, "same", mask=mask, out_fp32=out_fp32, colsum=not out_fp32, halo=True)
    ref = ref_fwd(x, staged_kernel(wk, scale), 3, 1, "same", mask=mask)
    assert torch.isfinite(got).all()
    tol = 2e-3 if out_fp32 else 2 ** -8 + 2e-3
    assert (got - ref).abs().max().item() <= tol * ref.abs().max().item()
    if cs is not None:
        assert torch.allclose(cs, got.sum((0, 1, 2)), rtol=2e-3, atol=1e-2 * ref.abs().max().item())
    base, _, _ = run_fwd(x, wk, scale, 3, 1, "same", mask=mask, out_fp32=out_fp32, halo=False)
    assert (got - base).abs().max().item() <= 1e-2 * ref.abs().max().item()


def test_engine_f_epilogue_all_stages():
    N, H, W, CI, CO = 2, 16, 24, 64, 128
    x = bf16_exact(N, H, W, CI, seed=3)
    wk = bf16_exact(3, 3, CI, CO, scale=0.05, seed=4)
    scale = (torch.rand(CO, dtype=torch.float64) + 0.5)
    shift = torch.randn(CO, dtype=torch.float64).float().double()
    addend = bf16_exact(N, H, W, CO, seed=5)
    mask = bf16_exact(N, H, W, CO, seed=6)
    got, cs, bmat = run_fwd(x, wk, scale, 3, 1, "same", shift=shift, addend=addend, mask=mask, relu=True, colsum=True)
    ref = ref_fwd(x, staged_kernel(wk, scale), 3, 1, "same", shift, addend, mask, True)
    tol = 2 ** -8 * ref.abs().max().item() + 1e-3       # bf16 output rounding
    assert (got - ref).abs().max().item() <= tol
    assert torch.allclose(cs, got.sum((0, 1, 2)), rtol=2e-3, atol=1e-2 * ref.abs().max().item())


DGRAD_CASES = [(1, 1, "valid", 64, 128, 16, 24), (3, 1, "same", 64, 64, 16, 24), (1, 2, "valid", 128, 64, 16, 24),
               (3, 2, "same", 128, 32, 20, 30), (3, 2, 1, 64, 128, 16, 24)]


@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16], ids=["fp32out_regs", "bf16out_tma"])
@pytest.mark.parametrize("kh,stride,padding,cin,cout,h,w", DGRAD_CASES)
def test_engine_f_dgrad(kh, stride, padding, cin, cout, h, w, out_dtype):
    from ursonet_b200 import lib
    N = 2
    wk = bf16_exact(kh, kh, cin, cout, scale=0.05, seed=7)
    g = P.make_geom(kh, stride, padding, cin, cout, h, w)
    du = bf16_exact(N, g.oh, g.ow, cout, seed=8)
    x = torch.zeros(N, h, w, cin, dtype=torch.float64, requires_grad=True)
    (ref,) = torch.autograd.grad(O.conv2d(x, wk, None, stride, padding), x, du)
    cop = P.ceil64(cout)
    du_d = torch.zeros(N, g.oh, g.ow, cop, dtype=torch.bfloat16, device=DEV)
    du_d[..., :cout] = du.to(torch.bfloat16).to(DEV)
    dx = torch.zeros(N, h, w, cin, dtype=out_dtype, device=DEV)
    for oph, opw, segs, tap_map in P.dgrad_phases(g):
        tgt = dx[:, oph::stride, opw::stride, :]
        if not segs:
            continue
        bmat = E.stage_cols(wk, None, tap_map).to(torch.bfloat16).to(DEV).contiguous()
        tw, th = P.pick_patch(tgt.shape[1], tgt.shape[2], 128)
        plan = lib.ConvGemm([du_d], bmat, segs, tgt, tgt.shape[2], tgt.shape[1], N, tw, th)
        plan.launch()
    torch.cuda.synchronize()
    got = dx.double().cpu()
    tol = 2e-3 if out_dtype == torch.float32 else 2 ** -8 + 2e-3
    assert (got - ref).abs().max().item() <= tol * ref.abs().max().item()


WGRAD_CASES = [  # kh, stride, padding, cin, cout, h, w, swap
    (1, 1, "valid", 128, 128, 16, 24, False),
    (1, 1, "valid", 64, 256, 16, 24, False),
    (1, 1, "valid", 64, 256, 16, 24, True),      # P = du (256 ch), Q = x (64 ch): transposed store
    (3, 1, "same", 128, 128, 16, 24, False),     # 9 taps -> 3 tap groups at BLOCK_Q 128
    (3, 1, "same", 64, 64, 20, 30, False),       # P has 64 valid channels: second atom is all OOB
    (1, 2, "valid", 256, 128, 16, 24, False),
    (3, 2, "same", 128, 64, 20, 30, False),
]


@pytest.mark.parametrize("kh,stride,padding,cin,cout,h,w,swap", WGRAD_CASES)
def test_engine_w(kh, stride, padding, cin, cout, h, w, swap):
    from ursonet_b200 import lib
    N = 3
    x = bf16_exact(N, h, w, cin, seed=9)
    g = P.make_geom(kh, stride, padding, cin, cout, h, w)
    du = bf16_exact(N, g.oh, g.ow, cout, seed=10)
    wk = torch.zeros(kh, kh, cin, cout, dtype=torch.float64, requires_grad=True)
    (ref,) = torch.autograd.grad(O.conv2d(x, wk, None, stride, padding), wk, du)
    xd = x.to(torch.bfloat16).to(DEV).contiguous()
    dud = du.to(torch.bfloat16).to(DEV).contiguous()
    G = torch.zeros(kh * kh, cin, cout, dtype=torch.float32, device=DEV)
    tw, th = P.pick_patch(g.oh, g.ow, 64)
    segs = P.wgrad_segments(g)
    if not swap:
        plan = lib.Wgrad(P.input_views(xd, stride), dud, segs, cin, cout, g.ow, g.oh, N, tw, th, G, cin * cout, cout, 1)
    else:
        assert kh == 1 and stride == 1
        plan = lib.Wgrad([dud], xd, [(0, 0, 0)], cout, cin, g.ow, g.oh, N, tw, th, G, cin * cout, 1, cout)
    plan.launch()
    torch.cuda.synchronize()
    got = G.double().cpu().reshape(kh, kh, cin, cout)
    assert (got - ref).abs().max().item() <= 2e-3 * ref.abs().max().item()


def test_engine_w_split_k_and_accumulate():
    """explicit split-K = 7 and a second launch accumulating on top (atomics add into g)"""
    from ursonet_b200 import lib
    N, H, W, CI, CO = 4, 16, 32, 128, 64
    x = bf16_exact(N, H, W, CI, seed=11)
    du = bf16_exact(N, H, W, CO, seed=12)
    ref = torch.einsum("nhwp,nhwq->pq", x, du)
    G = torch.zeros(1, CI, CO, dtype=torch.float32, device=DEV)
    xd, dud = x.to(torch.bfloat16).to(DEV), du.to(torch.bfloat16).to(DEV)
    plan = lib.Wgrad([xd], dud, [(0, 0, 0)], CI, CO, W, H, N, 8, 8,
                     G, CI * CO, CO, 1, split_k=7)
    plan.launch()
    plan.launch()
    torch.cuda.synchronize()
    assert (G[0].double().cpu() - 2 * ref).abs().max().item() <= 2e-3 * 2 * ref.abs().max().item()


def test_stem_stage_and_conv():
    from ursonet_b200 import lib
    B, H, W = 2, 64, 128
    g = torch.Generator().manual_seed(13)
    img = torch.randint(0, 256, (B, H, W, 3), generator=g, dtype=torch.uint8)
    mean = torch.tensor(O.MEAN_PIXEL, dtype=torch.float32)
    e_ref = E.stem_stage(img.double(), mean.double())
    s_out = torch.empty(B, H // 2 + 3, W // 2 + 3, 16, dtype=torch.bfloat16, device=DEV)   # compact staging ...
    e_out = lib.stem_view(s_out)              # ... read through the overlapping 64-"channel" view
    img_d, img_f, mean_d = img.to(DEV), img.float().to(DEV), mean.to(DEV)   # keep device buffers alive across calls
    lib.call("urso_stem_stage", img_d.data_ptr(), 1, 1, mean_d.data_ptr(), s_out.data_ptr(), B, H, W,
             0, lib.stream_ptr())
    torch.cuda.synchronize()
    assert (e_out.double().cpu() - e_ref.to(torch.bfloat16).double()).abs().max().item() <= 1.0  # 1 bf16 ulp at 255
    # fp32 image input path gives the same staging
    s2 = torch.empty_like(s_out)
    lib.call("urso_stem_stage", img_f.data_ptr(), 0, 1, mean_d.data_ptr(), s2.data_ptr(), B, H, W,
             0, lib.stream_ptr())
    torch.cuda.synchronize()
    assert torch.equal(s2, s_out)
    # 7x7/s2 conv through Engine F on the staged tensor
    wk = bf16_exact(7, 7, 3, 64, scale=0.05, seed=14)
    bmat = E.stage_rows(wk, None, P.stem_weight_index(3)).to(torch.bfloat16).to(DEV).contiguous()
    out = torch.empty(B, H // 2, W // 2, 64, dtype=torch.float32, device=DEV)
    tw, th = P.pick_patch(H // 2, W // 2, 128)
    lib.ConvGemm([e_out], bmat, P.stem_segments(), out, W // 2, H // 2, B, tw, th).launch()
    torch.cuda.synchronize()
    x_used = e_out.double().cpu()   # the engine multiplies the bf16-staged pixels
    ref = E.emu_convgemm([x_used], bmat.double().cpu(), P.stem_segments(), H // 2, W // 2)
    assert (out.double().cpu() - ref).abs().max().item() <= 2e-3 * ref.abs().max().item()
    # and the staged form agrees with the oracle's 7x7/s2 conv up to bf16 rounding of the pixels
    ref2 = O.conv2d(img.double() - mean.double(), wk, None, 2, 3)
    assert (out.double().cpu() - ref2).abs().max().item() <= 2e-2 * ref2.abs().max().item()
