"""End-to-end GPU parity of the engine (all launches through the C-ABI) against the fp64 oracle on the same seeded
weights / inputs: forward outputs, losses, every parameter gradient, and the weights after one optimizer step.

Two oracles are used (both fp64 arithmetic):
  * quant=True  -- the restatement with the engine's bf16 ROUNDING POINTS inserted (activations after each fused
    epilogue, BN-folded staged weights, gradient buffers after each ReLU mask).  Remaining differences are fp32
    accumulation order and the occasional 1-ulp rounding flip it causes, so the gates are tight:
      stem/pool exact to 2e-3; end-to-end outputs 3e-2; gradient direction cos >= 0.95 (a random, un-normalised deep
      net amplifies 1-ulp rounding flips chaotically).  The BUG-CATCHING comparison is layer-local: each layer's
      output / gradients recomputed in fp64 from the engine's own stored tensors must agree to 4e-3.
  * quant=False -- the plain fp64 graph.  It measures what bf16 storage costs on a randomly initialised (chaotic,
    un-normalised-input) network: forward <= 3e-2, gradient direction cos >= 0.93.  The 1e-3 forward gate of
    BASELINE.json is met by the split-bf16 parity mode (test_gpu_parity_mode.py), not by bf16 storage.
"""
import numpy as np
import pytest
import torch

from oracle import ursonet_oracle as O
from ursonet_b200.config import Config

pytestmark = pytest.mark.gpu


def make_cfg(backbone, classify=True, h=128, w=192, optimizer="SGD", ori_bins=8):
    cfg = Config()
    cfg.BACKBONE = backbone
    cfg.BOTTLENECK_WIDTH = 32
    cfg.BRANCH_SIZE = 256
    cfg.NR_DENSE_LAYERS = 1
    cfg.ORI_BINS_PER_DIM = ori_bins
    cfg.REGRESS_ORI = not classify
    cfg.REGRESS_LOC = True
    cfg.IMAGE_RESIZE_MODE = "pad64"
    cfg.IMAGE_MIN_DIM, cfg.IMAGE_MAX_DIM = h, w
    cfg.OPTIMIZER = optimizer
    cfg.LOSS_WEIGHTS = {"loc_loss": 1.0, "ori_loss": 1.0}
    cfg.NAME = "test"
    cfg.update()
    return cfg


def make_batch(cfg, B, seed=0):
    g = torch.Generator().manual_seed(seed)
    H, W = int(cfg.IMAGE_SHAPE[0]), int(cfg.IMAGE_SHAPE[1])
    img = torch.randint(0, 256, (B, H, W, 3), generator=g, dtype=torch.uint8)
    gt_loc = torch.stack([torch.rand(B, generator=g) * 4 - 2, torch.rand(B, generator=g) * 4 - 2,
                          torch.rand(B, generator=g) * 35 + 5], 1)
    if cfg.REGRESS_ORI:
        q = torch.nn.functional.normalize(torch.randn(B, 4, generator=g), dim=-1)
        gt_ori = q * torch.sign(q[:, 3:4])
    else:
        gt_ori = torch.softmax(torch.randn(B, cfg.ORI_BINS_PER_DIM ** 3, generator=g) * 4, -1)
    return img, gt_loc, gt_ori


def load_oracle_weights(engine, p64):
    engine.params.load_state_dict({k: v.numpy() for k, v in p64.items()})


def rel(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


@pytest.mark.parametrize("backbone,classify", [("resnet18", True), ("resnet50", True), ("resnet50", False),
                                               ("resnet34", False)])
def test_forward_matches_oracle(backbone, classify):
    from ursonet_b200.engine import Engine
    cfg = make_cfg(backbone, classify)
    B = 2
    p64 = O.init_weights(cfg, seed=1, pretrained_like=True)
    eng = Engine(cfg, B, training=False)
    load_oracle_weights(eng, p64)
    img, _, _ = make_batch(cfg, B)
    eng.img_u8.copy_(img)
    loc, ori = eng.forward(use_graph=False)
    torch.cuda.synchronize()
    # layer-local exactness: every conv output recomputed in fp64 from the engine's OWN bf16 input must agree to
    # bf16 output rounding (rms 2^-9, max 2^-8 + fp32 accumulation noise)
    P64 = {k: v.double() for k, v in p64.items()}
    mean = torch.tensor(O.MEAN_PIXEL, dtype=torch.float64)
    with torch.no_grad():
        for c in eng.graph.convs:
            x = (img.double() - mean).to(torch.bfloat16).double() if c.stem else eng.act[c.src].double().cpu()
            y = O.conv_bn(x, P64, c.name, c.bn, c.stride, c.padding, quant=True)
            if c.addend:
                y = y + eng.act[c.addend].double().cpu()
            if c.relu:
                y = torch.relu(y)
            got = eng.act[c.dst].double().cpu()
            rms = (got - y).norm().item() / max(y.norm().item(), 1e-30)
            assert rms <= 3e-3 and rel(got, y) <= 1e-2, (c.name, rms, rel(got, y))
    qt = {}
    qloc, qori = O.forward(p64, O.mold_image(img), cfg, qt, quant=True)
    assert rel(eng.act["pool1"].double().cpu(), qt["pool1"].detach()) <= 2e-3
    assert rel(loc.double().cpu(), qloc.detach()) <= 3e-2
    assert rel(ori.double().cpu(), qori.detach()) <= 3e-2
    taps = {}
    rloc, rori = O.forward(p64, O.mold_image(img), cfg, taps)
    # intermediate activations first: localises a failure to a layer
    for name in ["pool1", "bottleneck_layer"]:
        got = eng.act[name].double().cpu()
        assert rel(got, taps[name]) <= 3e-2, (name, rel(got, taps[name]))
    assert rel(loc.double().cpu(), rloc) <= 3e-2
    assert rel(ori.double().cpu(), rori) <= 3e-2
    # CUDA-graph replay gives the same result as eager launches
    loc2, ori2 = eng.forward(use_graph=True)
    loc3, ori3 = eng.forward(use_graph=True)
    torch.cuda.synchronize()
    assert torch.equal(loc2, loc3) and torch.equal(ori2, ori3)
    assert rel(loc2.double().cpu(), rloc) <= 3e-2


def local_backward_check(eng, p64, cfg):
    """Layer-local exactness of the backward plan.  A randomly initialised deep net amplifies 1-ulp bf16 rounding
    flips chaotically, so END-TO-END gradients can only be compared loosely; instead every layer's gradients are
    recomputed in fp64 FROM THE ENGINE'S OWN stored tensors (its bf16 input activation and its bf16 output gradient)
    and must match tightly.  This pins the wiring: which buffer feeds which launch, masks, fan-in, strides, phases."""
    g = eng.graph
    P64 = {k: v.double() for k, v in p64.items()}
    cons = {}
    for c in g.convs:
        cons.setdefault(c.src, []).append(c)
    adds = {}
    for c in g.convs:
        if c.addend:
            adds.setdefault(c.addend, []).append(c)
    mean = torch.tensor(O.MEAN_PIXEL, dtype=torch.float64)
    report = []
    for c in g.convs:
        du = eng.dact[c.dst].double().cpu()[..., :c.cout]
        if c.stem:
            x = (eng.img_u8.double().cpu() - mean).to(torch.bfloat16).double()
        else:
            x = eng.act[c.src].double().cpu()
        names = [c.name + "/kernel"] + ([c.name + "/bias"] if c.bias else []) + \
                ([c.bn + "/gamma", c.bn + "/beta"] if c.bn else [])
        leaves = {n: P64[n].clone().requires_grad_(True) for n in names}
        q = dict(P64); q.update(leaves)
        y = O.conv_bn(x, q, c.name, c.bn, c.stride, c.padding, quant=True)
        grads = torch.autograd.grad(y, [leaves[n] for n in names], du)
        for n, gref in zip(names, grads):
            got = eng.params.view(n, eng.grads).double().cpu()
            e = (got - gref).norm().item() / max(gref.norm().item(), 1e-30)
            report.append((n, e))
    for X, convs in cons.items():
        if X == "image":
            continue
        x = eng.act[X].double().cpu().requires_grad_(True)
        tot = torch.zeros_like(x)
        for c in convs:
            du = eng.dact[c.dst].double().cpu()[..., :c.cout]
            y = O.conv_bn(x, P64, c.name, c.bn, c.stride, c.padding, quant=True)
            tot = tot + torch.autograd.grad(y, x, du)[0]
        for c in adds.get(X, []):
            tot = tot + eng.dact[c.dst].double().cpu()
        # pool1 is masked too: (pool1 > 0) is exactly the stem's ReLU mask seen through the max (engine.py), which lets
        # the pool backward skip re-reading the stem output
        if X in g.relu_buffers or X == "pool1":
            tot = tot * (x.detach() > 0)
        got = eng.dact[X].double().cpu()
        e = (got - tot).norm().item() / max(tot.norm().item(), 1e-30)
        report.append(("d:" + X, e))
    return report


@pytest.mark.parametrize("backbone,classify,optimizer", [("resnet18", True, "SGD"), ("resnet50", True, "SGD"),
                                                         ("resnet50", False, "ADAM")])
def test_train_step_matches_oracle(backbone, classify, optimizer):
    from ursonet_b200.engine import Engine
    cfg = make_cfg(backbone, classify, optimizer=optimizer)
    B, lr = 2, 1e-3
    p64 = O.init_weights(cfg, seed=2, pretrained_like=True)
    eng = Engine(cfg, B, training=True)
    load_oracle_weights(eng, p64)
    img, gt_loc, gt_ori = make_batch(cfg, B, seed=3)
    eng.img_u8.copy_(img)
    eng.gt_loc.copy_(gt_loc)
    eng.gt_ori.copy_(gt_ori)
    # ---- 1. forward + backward only (no update): layer-local exactness, tolerance 4e-3 (bf16 rounding of du: 2^-9 rms)
    eng._phase_train()
    torch.cuda.synchronize()
    rep = local_backward_check(eng, p64, cfg)
    bad = [(n, round(e, 5)) for n, e in rep if e > 4e-3]
    assert not bad, bad[:12]
    # ---- 2. the full step against the oracle with the engine's rounding points (quant=True)
    eng.train_step(lr, use_graph=False)
    torch.cuda.synchronize()
    batch = (O.mold_image(img), gt_loc.double(), gt_ori.double())
    newp, info = O.train_step(p64, {}, batch, cfg, lr=lr, quant=True)
    losses = eng.losses.double().cpu()
    assert abs(losses[0].item() - info["loc_loss"].item()) <= 1e-2 * abs(info["loc_loss"].item()) + 1e-4
    assert abs(losses[1].item() - info["ori_loss"].item()) <= 1e-2 * abs(info["ori_loss"].item()) + 1e-4
    # end-to-end gradients (eng.grads now includes the regulariser): chaotic amplification of rounding flips through
    # the random net (53 layers for RN-50) limits this to direction -- the tight check is the layer-local one above
    gmax = max(g.norm().item() for g in info["grads"].values())
    for name, gref in info["grads"].items():
        n = gref.norm().item()
        if n < 1e-3 * gmax:
            continue
        got = eng.params.view(name, eng.grads).double().cpu()
        cos = (got * gref).sum().item() / (got.norm().item() * n)
        assert cos >= 0.95 and (got - gref).norm().item() / n <= 0.3, (name, cos)
    norm = float(torch.sqrt(eng.sumsq.double().cpu())[0])
    assert abs(norm - info["grad_norm"].item()) <= 3e-2 * info["grad_norm"].item()
    # the optimizer update itself is exact given the engine's own gradient: recompute it on the host in fp64
    gflat = eng.grads.double().cpu()
    gn = float(torch.sqrt((gflat * gflat).sum()))
    assert abs(gn - norm) <= 1e-4 * gn
    cf = cfg.GRADIENT_CLIP_NORM / gn if gn >= cfg.GRADIENT_CLIP_NORM else 1.0
    for name in info["grads"]:
        w0 = p64[name].float().double()
        gq = eng.params.view(name, eng.grads).double().cpu() * cf
        if optimizer == "SGD":
            ref = w0 - lr * gq                       # first step: v = -lr*g ; p += v
        else:
            m, v = 0.1 * gq, 0.001 * gq * gq
            lr_t = lr * (1 - 0.999) ** 0.5 / (1 - 0.9)
            ref = w0 - lr_t * m / (torch.sqrt(v) + 1e-7)
        got = eng.params.view(name).double().cpu()
        assert torch.allclose(got, ref, rtol=1e-5, atol=1e-6), name
    # ---- 3. against the un-quantised fp64 graph: direction of every sizeable gradient is preserved
    _, info64 = O.train_step(p64, {}, batch, cfg, lr=lr)
    for name, gref in info64["grads"].items():
        if gref.norm().item() < 1e-3 * gmax:
            continue
        got = eng.params.view(name, eng.grads).double().cpu()
        cos = (got * gref).sum().item() / (got.norm().item() * gref.norm().item())
        assert cos >= 0.93, (name, cos)
    # ---- 4. further steps through CUDA graphs run and change the weights
    before = eng.params.flat.clone()
    eng.train_step(lr, use_graph=True)
    eng.train_step(lr, use_graph=True)
    torch.cuda.synchronize()
    assert torch.isfinite(eng.params.flat).all()
    assert not torch.equal(before, eng.params.flat)


def test_set_trainable_freezes_layers():
    from ursonet_b200.engine import Engine
    cfg = make_cfg("resnet18")
    eng = Engine(cfg, 1, training=True)
    img, gt_loc, gt_ori = make_batch(cfg, 1)
    eng.img_u8.copy_(img); eng.gt_loc.copy_(gt_loc); eng.gt_ori.copy_(gt_ori)
    heads = r"(ori\_.*)|(loc\_.*)|(fpn\_.*)|(bottleneck_layer)"          # net.py:1088
    n = eng.params.set_trainable(heads)
    assert n > 0
    before = {k: v.copy() for k, v in eng.params.state_dict().items()}
    eng.train_step(1e-2, use_graph=False)
    torch.cuda.synchronize()
    after = eng.params.state_dict()
    for k in before:
        changed = not np.array_equal(before[k], after[k])
        layer = k.split("/")[0]
        is_head = layer.startswith(("ori_", "loc_")) or layer == "bottleneck_layer"
        assert changed == is_head or (is_head and k.endswith("bias")), k


class _Done:
    def wait(self):
        pass


@pytest.mark.parametrize("use_graph", [False, True])
def test_split_backward_for_overlapped_allreduce(use_graph):
    """The overlapped all-reduce schedule replays backward in two graphs (arena tail first).  With an identity
    'all-reduce' it must give the same step as the single-graph schedule (up to fp32 atomic summation order)."""
    from ursonet_b200.engine import Engine
    cfg = make_cfg("resnet50", True)
    B, lr = 2, 1e-3
    p64 = O.init_weights(cfg, seed=5, pretrained_like=True)
    img, gt_loc, gt_ori = make_batch(cfg, B, seed=6)
    outs = []
    for split in (False, True):
        eng = Engine(cfg, B, training=True)
        load_oracle_weights(eng, p64)
        eng.img_u8.copy_(img); eng.gt_loc.copy_(gt_loc); eng.gt_ori.copy_(gt_ori)
        assert eng._bwd_split is not None and 0 < eng._bwd_split[1] < eng.params.n_train
        seen = []
        ar = (lambda g: (seen.append(g.numel()), _Done())[1]) if split else None
        eng.train_step(lr, use_graph=use_graph, allreduce_async=ar)
        torch.cuda.synchronize()
        first = eng.params.flat.clone()
        eng.train_step(lr, use_graph=use_graph, allreduce_async=ar)
        torch.cuda.synchronize()
        if split:   # two pieces per step that tile the arena, tail (>= 90 % of it) first
            assert len(seen) == 4 and seen[0] + seen[1] == eng.params.n_train and seen[0] >= 0.9 * eng.params.n_train
        outs.append((first, eng.losses.clone()))
    (p0, l0), (p1, l1) = outs
    # after ONE step the weights agree to fp32 atomic-order noise; the second step's losses (a chaotic random net
    # amplifies that noise) only have to agree loosely -- a double-counted accumulation would be off by far more
    assert (p0 - p1).abs().max().item() <= 2e-6 * p0.abs().max().item() + 1e-7
    assert torch.allclose(l0, l1, rtol=3e-2, atol=1e-3)
