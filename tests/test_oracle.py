"""The oracle restatement checked against itself and against hand-computed semantics of the traps listed in
SURVEY.md App. A (the reference ships no golden vectors for the network: 'parity unpinned')."""
import math

import pytest
import torch

from oracle import ursonet_oracle as O
from ursonet_b200.config import Config


def small_cfg(backbone="resnet18", classify=True, optimizer="SGD"):
    c = Config()
    c.BACKBONE, c.BOTTLENECK_WIDTH, c.BRANCH_SIZE, c.NR_DENSE_LAYERS = backbone, 32, 64, 1
    c.ORI_BINS_PER_DIM, c.REGRESS_ORI, c.OPTIMIZER = 4, not classify, optimizer
    c.IMAGE_MIN_DIM, c.IMAGE_MAX_DIM = 64, 128
    c.update()
    return c


def batch(cfg, B=2, seed=0):
    g = torch.Generator().manual_seed(seed)
    H, W = int(cfg.IMAGE_SHAPE[0]), int(cfg.IMAGE_SHAPE[1])
    img = torch.randint(0, 256, (B, H, W, 3), generator=g, dtype=torch.uint8)
    loc = torch.randn(B, 3, generator=g, dtype=torch.float64) * 3
    if cfg.REGRESS_ORI:
        ori = torch.nn.functional.normalize(torch.randn(B, 4, generator=g, dtype=torch.float64), dim=-1)
    else:
        ori = torch.softmax(torch.randn(B, cfg.ORI_BINS_PER_DIM ** 3, generator=g, dtype=torch.float64), -1)
    return O.mold_image(img), loc, ori


def test_tf_same_padding_is_bottom_right_on_even_maps():
    x = torch.arange(16, dtype=torch.float64).reshape(1, 4, 4, 1)
    y = O.maxpool3x3s2_same(x)                       # windows rows {0,1,2},{2,3}; never the PyTorch padding=1 result
    assert y.reshape(-1).tolist() == [10.0, 11.0, 14.0, 15.0]
    w = torch.ones(3, 3, 1, 1, dtype=torch.float64)
    c = O.conv2d(x, w, None, 2, "same")              # bottleneck_layer rule (net.py:639)
    assert c.reshape(-1).tolist() == [sum([0, 1, 2, 4, 5, 6, 8, 9, 10]), sum([2, 3, 6, 7, 10, 11]),
                                      sum([8, 9, 10, 12, 13, 14]), sum([10, 11, 14, 15])]


def test_strided_1x1_samples_even_pixels():
    x = torch.arange(16, dtype=torch.float64).reshape(1, 4, 4, 1)
    y = O.conv2d(x, torch.ones(1, 1, 1, 1, dtype=torch.float64), None, 2, "valid")
    assert y.reshape(-1).tolist() == [0.0, 2.0, 8.0, 10.0]


def test_losses_hand_values():
    y = torch.tensor([[0.5, 0.5, 0.0], [0.0, 1.0, 0.0]], dtype=torch.float64)
    z = torch.tensor([[0.0, 0.0, 0.0], [0.0, 2.0, 0.0]], dtype=torch.float64)
    want = (math.log(3) + (-(2 - math.log(2 + math.e ** 2)))) / 2
    assert math.isclose(O.softmax_loss(y, z).item(), want, rel_tol=1e-12)
    gt = torch.tensor([[3.0, 0.0, 4.0], [0.0, 0.0, 0.0]], dtype=torch.float64)
    pr = torch.tensor([[3.0, 0.0, 0.0], [0.0, 3.0, 0.0]], dtype=torch.float64)
    assert math.isclose(O.rel_loss(gt, pr).item(), 5.0 / 5.0, rel_tol=1e-12)      # batch-global norms (net.py:757)
    q = torch.tensor([[0.0, 0.0, 0.0, 1.0]], dtype=torch.float64)
    assert math.isclose(O.one_minus_dot_prod(q, -q).item(), 0.0, abs_tol=1e-15)    # |dot|: q and -q are the same pose


def test_keras_clipnorm_is_global():
    g = {"a": torch.full((4,), 3.0, dtype=torch.float64), "b": torch.full((9,), 4.0, dtype=torch.float64)}
    clipped, norm = O.clip_by_global_norm(g, 5.0)
    assert math.isclose(norm.item(), math.sqrt(4 * 9 + 9 * 16))
    tot = math.sqrt(sum((v * v).sum().item() for v in clipped.values()))
    assert math.isclose(tot, 5.0, rel_tol=1e-12)
    same, _ = O.clip_by_global_norm(g, 100.0)
    assert same is g


def test_sgd_and_amsgrad_first_steps():
    p = {"w": torch.tensor([1.0, -2.0], dtype=torch.float64)}
    g = {"w": torch.tensor([0.5, 0.25], dtype=torch.float64)}
    st = {}
    p1, _ = O.sgd_step(p, st, g, 0.1, 0.9, 5.0)
    assert torch.allclose(p1["w"], torch.tensor([0.95, -2.025], dtype=torch.float64))
    p2, _ = O.sgd_step(p1, st, g, 0.1, 0.9, 5.0)             # v = 0.9*(-0.05) - 0.05
    assert torch.allclose(p2["w"], torch.tensor([0.95 - 0.095, -2.025 - 0.0475], dtype=torch.float64))
    st = {}
    a1, _ = O.amsgrad_step(p, st, g, 0.1, 5.0)               # first Adam step moves by ~lr * sign(g)
    assert torch.allclose(p["w"] - a1["w"], torch.tensor([0.1, 0.1], dtype=torch.float64), atol=1e-5)


@pytest.mark.parametrize("backbone", ["resnet18", "resnet50"])
def test_folded_form_equals_bn_form(backbone):
    """conv(x, W*scale)+shift == BN(conv(x,W)+b): the engine's folding is exact algebra (quant path without rounding
    is checked by monkeypatching the rounding to identity)."""
    cfg = small_cfg(backbone)
    p = O.init_weights(cfg, seed=3, pretrained_like=True)
    img, loc, ori = batch(cfg)
    a = O.forward(p, img, cfg)
    q_saved = O._q
    O._q = lambda x: x
    try:
        b = O.forward(p, img, cfg, quant=True)
    finally:
        O._q = q_saved
    assert torch.allclose(a[0], b[0], rtol=1e-9, atol=1e-9) and torch.allclose(a[1], b[1], rtol=1e-9, atol=1e-9)


def test_regulariser_gradient_formula():
    cfg = small_cfg()
    p = O.init_weights(cfg, seed=4)
    name = "bottleneck_layer/kernel"
    w = p[name].clone().requires_grad_(True)
    q = dict(p); q[name] = w
    (g,) = torch.autograd.grad(O.reg_loss(q, cfg), w)
    assert torch.allclose(g, 2 * cfg.WEIGHT_DECAY * p[name] / p[name].numel())
    assert not O.is_regularised("bn_conv0/gamma") and O.is_regularised("conv1/bias")


@pytest.mark.parametrize("classify,optimizer", [(True, "SGD"), (False, "ADAM")])
def test_train_step_decreases_loss(classify, optimizer):
    cfg = small_cfg("resnet18", classify, optimizer)
    p = O.init_weights(cfg, seed=5)
    b = batch(cfg)
    st = {}
    l0 = O.total_loss(p, b, cfg)[0].item()
    for _ in range(3):
        p, info = O.train_step(p, st, b, cfg, lr=1e-4)
    assert O.total_loss(p, b, cfg)[0].item() < l0


def test_clr_triangular():
    assert math.isclose(O.clr_triangular(0, 1e-4, 5e-4, 4000), 1e-4)
    assert math.isclose(O.clr_triangular(4000, 1e-4, 5e-4, 4000), 5e-4)
    assert math.isclose(O.clr_triangular(6000, 1e-4, 5e-4, 4000), 3e-4)
    assert math.isclose(O.clr_triangular(8000, 1e-4, 5e-4, 4000), 1e-4)


# ------------------------------------------------------------------------------------------------ oracle hedge
# The network oracle is unpinned (no TF/Keras here).  oracle/direct_oracle.py is a second restatement written
# independently (numpy, direct loops over output pixels, its own walk through net.py's builder functions); both must
# agree on the same random weights.  A semantic slip (TF-SAME offsets, Keras-v1 stride placement, the one-BN shallow
# block, ReLU'd logits, flatten order) would have to be made twice to survive.
@pytest.mark.parametrize("backbone,classify", [("resnet18", True), ("resnet34", False), ("resnet50", True),
                                               ("resnet101", False)])
def test_two_independent_restatements_agree(backbone, classify):
    import numpy as np
    from oracle import direct_oracle as D2
    cfg = small_cfg(backbone, classify)
    cfg.IMAGE_MIN_DIM, cfg.IMAGE_MAX_DIM = 64, 128
    cfg.update()
    p = O.init_weights(cfg, seed=3, pretrained_like=True)
    img, gt_loc, gt_ori = batch(cfg, B=2, seed=4)
    with torch.no_grad():
        loc, ori = O.forward(p, img, cfg)
    loc2, ori2 = D2.forward({k: v.numpy() for k, v in p.items()}, img.numpy(), cfg)
    assert np.abs(loc.numpy() - loc2).max() <= 1e-10 * max(1.0, np.abs(loc2).max())
    assert np.abs(ori.numpy() - ori2).max() <= 1e-10 * max(1.0, np.abs(ori2).max())
    # losses
    if classify:
        assert math.isclose(O.softmax_loss(gt_ori, ori).item(), D2.softmax_loss(gt_ori.numpy(), ori2), rel_tol=1e-10)
    else:
        assert math.isclose(O.one_minus_dot_prod(gt_ori, ori).item(), D2.one_minus_dot_prod(gt_ori.numpy(), ori2),
                            rel_tol=1e-10, abs_tol=1e-12)
    assert math.isclose(O.rel_loss(gt_loc, loc).item(), D2.rel_loss(gt_loc.numpy(), loc2), rel_tol=1e-10)


def test_direct_oracle_primitives_against_hand_values():
    import numpy as np
    from oracle import direct_oracle as D2
    x = np.arange(16, dtype=np.float64).reshape(1, 4, 4, 1)
    assert D2.maxpool_3x3_s2_same_direct(x).reshape(-1).tolist() == [10.0, 11.0, 14.0, 15.0]
    y = D2.conv_direct(x, np.ones((3, 3, 1, 1)), None, 2, "same")
    assert y.reshape(-1).tolist() == [sum([0, 1, 2, 4, 5, 6, 8, 9, 10]), sum([2, 3, 6, 7, 10, 11]),
                                      sum([8, 9, 10, 12, 13, 14]), sum([10, 11, 14, 15])]
    assert D2.conv_direct(x, np.ones((1, 1, 1, 1)), None, 2, "valid").reshape(-1).tolist() == [0.0, 2.0, 8.0, 10.0]
    # odd map: SAME pads symmetrically (1 before, 1 after for k=3, s=2, n=5)
    assert D2.tf_same_pads(5, 3, 2) == (3, 1, 1) and D2.tf_same_pads(4, 3, 2) == (2, 0, 1)
    assert D2.tf_same_pads(6, 3, 1) == (6, 1, 1) and D2.tf_same_pads(8, 7, 2) == (4, 2, 3)


def test_fp32_oracle_gap_to_fp64_is_far_below_the_parity_gate():
    """SURVEY 8c (last row): how far the SAME restatement run in fp32 (what TF-CPU computes in) sits from fp64 -- this
    anchors the 1e-3 forward tolerance.  Measured here at 128x192 on RN-50: ~1e-6 relative."""
    cfg = small_cfg("resnet50", True)
    cfg.IMAGE_MIN_DIM, cfg.IMAGE_MAX_DIM = 128, 192
    cfg.update()
    p = O.init_weights(cfg, seed=5, pretrained_like=True)
    img, _, _ = batch(cfg, B=1, seed=6)
    with torch.no_grad():
        loc64, ori64 = O.forward(p, img, cfg)
        loc32, ori32 = O.forward({k: v.float() for k, v in p.items()}, img.float(), cfg)
    e_loc = (loc32.double() - loc64).abs().max().item() / loc64.abs().max().item()
    e_ori = (ori32.double() - ori64).abs().max().item() / ori64.abs().max().item()
    print("fp32 vs fp64 oracle gap: loc %.2e ori %.2e" % (e_loc, e_ori))
    assert e_loc < 1e-4 and e_ori < 1e-4
