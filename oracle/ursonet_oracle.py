"""CPU oracle for the UrsoNet hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product path (ursonet_b200/) never does.

PARITY UNPINNED: the reference (pedropro/UrsoNet @ 8e59d9b) ships no tests, golden
vectors or seeds for this path, and its arithmetic lives in un-vendored, un-pinned
third-party code (tensorflow>=1.9, keras>=2.1.6; requirements.txt:9-10) that cannot be
installed for Python 3.12.  This file therefore restates
  * the graph            net.py:85-199 (deep), 208-282 (shallow), 288-352 (heads), 639-643
  * the losses           net.py:705-762
  * compile()            net.py:973-1017 (loss weights, L2 regulariser, optimizer)
and the published Keras-2 / TF-1 semantics of the library ops it calls:
TF 'SAME' padding, inference-mode BatchNormalization(eps=1e-3), Dense, softmax_cross_entropy
(mean over batch), tf.norm (Frobenius), K.l2_normalize, Keras-2 global-norm `clipnorm`,
SGD-momentum and Adam(amsgrad=True).  The reference functions that DO run here
(se3lib.*, utils.encode_ori*, utils.stable_softmax) are pinned by tests/golden/ fixtures.

Weights are a flat dict  "<keras layer name>/<weight name>" -> torch tensor in KERAS layout
(Conv2D kernel HWIO, Dense kernel [in,out]).  All maths runs in the dtype of the weights
(float64 for the gold oracle, float32 for the "what TF-CPU would give" leg).
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

BN_EPS = 1e-3  # keras.layers.BatchNormalization default epsilon (net.py:60)


# --------------------------------------------------------------------------------------
# architecture description (layer names follow net.py exactly)
# --------------------------------------------------------------------------------------
def deep_blocks(backbone: str) -> List[Tuple[int, str, bool, int, Tuple[int, int, int]]]:
    """(stage, block letter, has conv shortcut, stride, filters) per net.py:178-196."""
    n4 = {"resnet50": 5, "resnet101": 22}[backbone]
    out = [(2, "a", True, 1, (64, 64, 256)), (2, "b", False, 1, (64, 64, 256)), (2, "c", False, 1, (64, 64, 256))]
    out += [(3, "a", True, 2, (128, 128, 512))] + [(3, c, False, 1, (128, 128, 512)) for c in "bcd"]
    out += [(4, "a", True, 2, (256, 256, 1024))]
    out += [(4, chr(98 + i), False, 1, (256, 256, 1024)) for i in range(n4)]
    out += [(5, "a", True, 2, (512, 512, 2048)), (5, "b", False, 1, (512, 512, 2048)), (5, "c", False, 1, (512, 512, 2048))]
    return out


def shallow_blocks(backbone: str) -> List[Tuple[int, int, int, int, str]]:
    """(stage, block, filters, stride, cut) per net.py:261-280."""
    reps = [2, 2, 2, 2] if backbone == "resnet18" else [3, 4, 6, 3]
    out = []
    for stage, rep in enumerate(reps):
        for block in range(rep):
            filt = 64 * 2 ** stage
            if block == 0:
                out.append((stage, block, filt, 1 if stage == 0 else 2, "post"))
            else:
                out.append((stage, block, filt, 1, "pre"))
    return out


def weight_shapes(cfg) -> Dict[str, Tuple[int, ...]]:
    """Every weight of the model (trainable + BN moving stats), Keras layout, creation order."""
    s: Dict[str, Tuple[int, ...]] = {}

    def conv(name, kh, cin, cout, bias):
        s[name + "/kernel"] = (kh, kh, cin, cout)
        if bias:
            s[name + "/bias"] = (cout,)

    def bn(name, c):
        for w in ("gamma", "beta", "moving_mean", "moving_variance"):
            s[f"{name}/{w}"] = (c,)

    def dense(name, cin, cout):
        s[name + "/kernel"] = (cin, cout)
        s[name + "/bias"] = (cout,)

    cin = cfg.NR_IMAGE_CHANNELS
    if cfg.BACKBONE in ("resnet50", "resnet101"):
        conv("conv1", 7, cin, 64, True); bn("bn_conv1", 64)
        c = 64
        for stage, blk, has_sc, _stride, (f1, f2, f3) in deep_blocks(cfg.BACKBONE):
            cb, bb = f"res{stage}{blk}_branch", f"bn{stage}{blk}_branch"
            conv(cb + "2a", 1, c, f1, True); bn(bb + "2a", f1)
            conv(cb + "2b", 3, f1, f2, True); bn(bb + "2b", f2)
            conv(cb + "2c", 1, f2, f3, True); bn(bb + "2c", f3)
            if has_sc:
                conv(cb + "1", 1, c, f3, True); bn(bb + "1", f3)
            c = f3
    else:
        conv("conv0", 7, cin, 64, False); bn("bn_conv0", 64)
        c = 64
        for stage, block, filt, _stride, cut in shallow_blocks(cfg.BACKBONE):
            base = f"stage{stage + 1}_unit{block + 1}_"
            if cut == "post":
                conv(base + "sc", 1, c, filt, False)
            conv(base + "conv1", 3, c, filt, False); bn(base + "bn2", filt)
            conv(base + "conv2", 3, filt, filt, False)
            c = filt
    conv("bottleneck_layer", 3, c, cfg.BOTTLENECK_WIDTH, True)
    h, w = int(cfg.IMAGE_SHAPE[0]), int(cfg.IMAGE_SHAPE[1])
    nr_features = int(cfg.BOTTLENECK_WIDTH * h * w / 64 ** 2)  # net.py:640
    for branch in ("loc", "ori"):
        f = nr_features
        for i in range(cfg.NR_DENSE_LAYERS):
            dense(f"{branch}_dense_{i}", f, cfg.BRANCH_SIZE)
            f = cfg.BRANCH_SIZE
        if branch == "loc":
            dense("loc_final", f, 3 if cfg.REGRESS_LOC else cfg.LOC_BINS_PER_DIM ** 3)
        elif cfg.REGRESS_ORI:
            if cfg.ORIENTATION_PARAM == "quaternion":
                dense("ori_q", f, 4)
            else:
                dense("ori_final", f, 3)
        else:
            dense("ori_final", f, cfg.ORI_BINS_PER_DIM ** 3)
    return s


def is_trainable(name: str) -> bool:
    return not (name.endswith("moving_mean") or name.endswith("moving_variance"))


def is_regularised(name: str) -> bool:
    """net.py:1008-1011: every trainable weight whose name lacks 'gamma' / 'beta'."""
    return is_trainable(name) and "gamma" not in name and "beta" not in name


def init_weights(cfg, seed: int = 0, pretrained_like: bool = False, dtype=torch.float64) -> Dict[str, torch.Tensor]:
    """Keras defaults: Glorot-uniform kernels, zero biases, BN (1,0,0,1).
    pretrained_like=True also draws non-trivial BN statistics / biases so folding is exercised."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shp in weight_shapes(cfg).items():
        if name.endswith("/kernel"):
            if len(shp) == 4:
                fan_in, fan_out = shp[0] * shp[1] * shp[2], shp[0] * shp[1] * shp[3]
            else:
                fan_in, fan_out = shp
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            w = (torch.rand(shp, generator=g, dtype=torch.float64) * 2 - 1) * lim
        elif name.endswith("/gamma"):
            w = torch.ones(shp, dtype=torch.float64)
            if pretrained_like:
                w = 0.5 + torch.rand(shp, generator=g, dtype=torch.float64)
        elif name.endswith("/moving_variance"):
            w = torch.ones(shp, dtype=torch.float64)
            if pretrained_like:
                w = 0.5 + torch.rand(shp, generator=g, dtype=torch.float64)
        else:  # bias, beta, moving_mean
            w = torch.zeros(shp, dtype=torch.float64)
            if pretrained_like:
                w = 0.1 * torch.randn(shp, generator=g, dtype=torch.float64)
        out[name] = w.to(dtype)
    return out


# --------------------------------------------------------------------------------------
# library-op restatements
# --------------------------------------------------------------------------------------
def same_pad(n: int, k: int, s: int) -> Tuple[int, int]:
    """TF 'SAME': out=ceil(n/s); total=max((out-1)*s+k-n,0); before=total//2, after=rest."""
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def conv2d(x, w, b=None, stride=1, padding="valid"):
    """x NHWC, w HWIO (Keras) -> NHWC.  padding: 'valid' | 'same' | int (explicit symmetric ZeroPadding2D)."""
    kh, kw = w.shape[0], w.shape[1]
    xn = x.permute(0, 3, 1, 2)
    if padding == "same":
        pt, pb = same_pad(x.shape[1], kh, stride)
        pl, pr = same_pad(x.shape[2], kw, stride)
        xn = F.pad(xn, (pl, pr, pt, pb))
    elif isinstance(padding, int) and padding > 0:
        xn = F.pad(xn, (padding,) * 4)
    y = F.conv2d(xn, w.permute(3, 2, 0, 1), b, stride=stride)
    return y.permute(0, 2, 3, 1)


def maxpool3x3s2_same(x):
    """KL.MaxPooling2D((3,3), strides=(2,2), padding='same') (net.py:176,258); padded cells never win."""
    pt, pb = same_pad(x.shape[1], 3, 2)
    pl, pr = same_pad(x.shape[2], 3, 2)
    xn = F.pad(x.permute(0, 3, 1, 2), (pl, pr, pt, pb), value=float("-inf"))
    return F.max_pool2d(xn, 3, 2).permute(0, 2, 3, 1)


def batchnorm_frozen(x, p, name, eps=BN_EPS):
    """BatchNorm(training=False) (net.py:60-76 with config.TRAIN_BN=False): moving-stat affine."""
    scale = p[name + "/gamma"] / torch.sqrt(p[name + "/moving_variance"] + eps)
    return (x - p[name + "/moving_mean"]) * scale + p[name + "/beta"]


# --------------------------------------------------------------------------------------
# the graph (net.py:161-199, 242-282, 288-352, 639-643)
# --------------------------------------------------------------------------------------
class _GradRound(torch.autograd.Function):
    """identity forward; backward rounds the gradient to bf16 (models the engine's bf16 gradient buffers)."""

    @staticmethod
    def forward(ctx, x):
        return x

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


def _q(x):
    """round to bf16 with a straight-through gradient (models bf16 storage of activations / staged weights)."""
    return x + (x.to(torch.bfloat16).to(x.dtype) - x).detach()


def conv_bn(x, p, conv, bn, stride, padding, quant):
    """Conv2D(+bias) -> frozen BatchNorm.  quant=True evaluates the algebraically identical folded form the engine
    uses: conv(x, bf16(W * scale)) + shift with scale = gamma/sqrt(var+eps), shift = (bias-mean)*scale+beta."""
    w, b = p[conv + "/kernel"], p.get(conv + "/bias")
    if not quant:
        y = conv2d(x, w, b, stride, padding)
        return batchnorm_frozen(y, p, bn) if bn else y
    if bn:
        scale = p[bn + "/gamma"] / torch.sqrt(p[bn + "/moving_variance"] + BN_EPS)
        shift = ((b if b is not None else 0.0) - p[bn + "/moving_mean"]) * scale + p[bn + "/beta"]
    else:
        scale, shift = 1.0, (b if b is not None else 0.0)
    return conv2d(x, _q(w * scale), None, stride, padding) + shift


def backbone_forward(p, x, cfg, taps=None, quant=False):
    """quant=True inserts the engine's bf16 rounding points (activations after each fused epilogue, staged weights,
    gradient buffers after each ReLU mask) so that parity can be checked at the kernel's own precision."""
    def tap(name, v):
        if taps is not None:
            taps[name] = v
        return v

    def act(u, relu=True):
        if quant:
            u = _GradRound.apply(u)
        y = F.relu(u) if relu else u
        return _q(y) if quant else y

    if quant:
        x = _q(x)
    if cfg.BACKBONE in ("resnet50", "resnet101"):
        x = conv_bn(x, p, "conv1", "bn_conv1", 2, 3, quant)             # ZeroPadding2D(3) + 7x7/s2 valid
        x = tap("conv1_relu", act(x))
        x = tap("pool1", maxpool3x3s2_same(x))
        if quant:
            x = _GradRound.apply(x)
        for stage, blk, has_sc, stride, _f in deep_blocks(cfg.BACKBONE):
            cb, bb = f"res{stage}{blk}_branch", f"bn{stage}{blk}_branch"
            y = act(conv_bn(x, p, cb + "2a", bb + "2a", stride, "valid", quant))      # stride on FIRST 1x1
            y = act(conv_bn(y, p, cb + "2b", bb + "2b", 1, "same", quant))
            y = conv_bn(y, p, cb + "2c", bb + "2c", 1, "valid", quant)
            sc = conv_bn(x, p, cb + "1", bb + "1", stride, "valid", quant) if has_sc else x
            if quant and has_sc:
                sc = _q(sc)                                                            # shortcut branch stored in bf16
            x = tap(f"res{stage}{blk}_out", act(y + sc))
    else:
        x = conv_bn(x, p, "conv0", "bn_conv0", 2, 3, quant)
        x = tap("conv0_relu", act(x))
        x = tap("pool1", maxpool3x3s2_same(x))
        if quant:
            x = _GradRound.apply(x)
        for stage, block, _filt, stride, cut in shallow_blocks(cfg.BACKBONE):
            base = f"stage{stage + 1}_unit{block + 1}_"
            sc = conv_bn(x, p, base + "sc", None, stride, "valid", quant) if cut == "post" else x   # raw input, no BN
            if quant and cut == "post":
                sc = _q(sc)
            y = act(conv_bn(x, p, base + "conv1", base + "bn2", stride, 1, quant))    # ZeroPadding2D(1) + valid
            y = conv_bn(y, p, base + "conv2", None, 1, 1, quant)
            x = tap(base + "relu2", act(y + sc))
    return x


def forward(p, images, cfg, taps=None, quant=False):
    """images: [B,H,W,3] mean-subtracted ('molded', net.py:1337-1348). Returns (loc, ori).
    Training and inference graphs give the same (loc, ori) (frozen BN), net.py:680/691."""
    c5 = backbone_forward(p, images, cfg, taps, quant)
    c6 = conv_bn(c5, p, "bottleneck_layer", None, 2, "same", quant)
    if quant:
        c6 = _GradRound.apply(c6)
    if taps is not None:
        taps["bottleneck_layer"] = c6
    feat = c6.reshape(c6.shape[0], -1)                                  # NHWC flatten (net.py:298,332)
    outs = {}
    for branch in ("loc", "ori"):
        x = feat
        for i in range(cfg.NR_DENSE_LAYERS):
            n = f"{branch}_dense_{i}"
            x = F.relu(x @ p[n + "/kernel"] + p[n + "/bias"])
        outs[branch] = x
    if cfg.REGRESS_LOC:
        loc = outs["loc"] @ p["loc_final/kernel"] + p["loc_final/bias"]
    else:
        loc = F.relu(outs["loc"] @ p["loc_final/kernel"] + p["loc_final/bias"])
    if cfg.REGRESS_ORI:
        if cfg.ORIENTATION_PARAM == "quaternion":
            q = outs["ori"] @ p["ori_q/kernel"] + p["ori_q/bias"]
            ori = q * torch.rsqrt(torch.clamp((q * q).sum(-1, keepdim=True), min=1e-12))  # K.l2_normalize
        else:
            ori = outs["ori"] @ p["ori_final/kernel"] + p["ori_final/bias"]
    else:
        ori = F.relu(outs["ori"] @ p["ori_final/kernel"] + p["ori_final/bias"])   # ReLU'd logits (net.py:350)
    return loc, ori


# --------------------------------------------------------------------------------------
# losses (net.py:705-762) and compile() (net.py:973-1017)
# --------------------------------------------------------------------------------------
def softmax_loss(y_gt, y_pred):
    """tf.losses.softmax_cross_entropy(onehot_labels=y_gt, logits=y_pred): mean_b(-sum_k y log_softmax)."""
    return -(y_gt * F.log_softmax(y_pred, dim=-1)).sum(-1).mean()


def rel_loss(y_gt, y_pred):
    """tf.norm((y_gt - y_pred) / tf.norm(y_gt)) -- Frobenius norms over the WHOLE batch tensor."""
    return torch.linalg.norm((y_gt - y_pred) / torch.linalg.norm(y_gt))


def one_minus_dot_prod(y_true, y_pred):
    return (1 - (y_true * y_pred).sum(-1, keepdim=True).abs()).mean()


def head_losses(loc, ori, gt_loc, gt_ori, cfg):
    loc_loss = rel_loss(gt_loc, loc) if cfg.REGRESS_LOC else softmax_loss(gt_loc, loc)
    ori_loss = one_minus_dot_prod(gt_ori, ori) if cfg.REGRESS_ORI else softmax_loss(gt_ori, ori)
    return loc_loss, ori_loss


def reg_loss(p, cfg, trainable=None):
    tot = 0.0
    for name, w in p.items():
        if is_regularised(name) and (trainable is None or name in trainable):
            tot = tot + cfg.WEIGHT_DECAY * (w * w).sum() / w.numel()
    return tot


def total_loss(p, batch, cfg, trainable=None, quant=False):
    images, gt_loc, gt_ori = batch
    loc, ori = forward(p, images, cfg, quant=quant)
    loc_loss, ori_loss = head_losses(loc, ori, gt_loc, gt_ori, cfg)
    wl = cfg.LOSS_WEIGHTS.get("loc_loss", 1.0)
    wo = cfg.LOSS_WEIGHTS.get("ori_loss", 1.0)
    tot = wl * loc_loss + wo * ori_loss + reg_loss(p, cfg, trainable)
    return tot, (loc, ori, wl * loc_loss, wo * ori_loss)


def gradients(p, batch, cfg, trainable=None, quant=False):
    """d total_loss / d trainable weights by autograd on the restated forward (what TF autodiff computes)."""
    names = [n for n in p if is_trainable(n) and (trainable is None or n in trainable)]
    leaves = {n: p[n].detach().clone().requires_grad_(True) for n in names}
    q = dict(p); q.update(leaves)
    tot, aux = total_loss(q, batch, cfg, trainable, quant)
    grads = torch.autograd.grad(tot, [leaves[n] for n in names])
    return dict(zip(names, grads)), tot.detach(), tuple(a.detach() for a in aux)


def clip_by_global_norm(grads: Dict[str, torch.Tensor], clipnorm: float):
    """Keras-2 optimizers.get_gradients: norm = sqrt(sum_all ||g||^2); if norm >= c: g *= c/norm."""
    norm = torch.sqrt(sum((g * g).sum() for g in grads.values()))
    if clipnorm > 0 and norm >= clipnorm:
        grads = {n: g * (clipnorm / norm) for n, g in grads.items()}
    return grads, norm


def sgd_step(p, state, grads, lr, momentum, clipnorm):
    """keras.optimizers.SGD(lr, momentum, clipnorm): v = m*v - lr*g ; p = p + v."""
    grads, norm = clip_by_global_norm(grads, clipnorm)
    newp = dict(p)
    for n, g in grads.items():
        v = momentum * state.get(("v", n), torch.zeros_like(g)) - lr * g
        state[("v", n)] = v
        newp[n] = p[n] + v
    return newp, norm


def amsgrad_step(p, state, grads, lr, clipnorm, beta1=0.9, beta2=0.999, eps=1e-7):
    """keras.optimizers.Adam(lr, amsgrad=True, clipnorm) (Keras 2.1.6-2.2.4 get_updates)."""
    grads, norm = clip_by_global_norm(grads, clipnorm)
    t = state.get("t", 0) + 1
    state["t"] = t
    lr_t = lr * math.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
    newp = dict(p)
    for n, g in grads.items():
        m = beta1 * state.get(("m", n), torch.zeros_like(g)) + (1 - beta1) * g
        v = beta2 * state.get(("v", n), torch.zeros_like(g)) + (1 - beta2) * g * g
        vhat = torch.maximum(state.get(("vhat", n), torch.zeros_like(g)), v)
        state[("m", n)], state[("v", n)], state[("vhat", n)] = m, v, vhat
        newp[n] = p[n] - lr_t * m / (torch.sqrt(vhat) + eps)
    return newp, norm


def train_step(p, state, batch, cfg, lr=None, trainable=None, quant=False):
    """One Keras train_on_batch: grads of (weighted head losses + L2 reg) -> clip -> update."""
    lr = cfg.LEARNING_RATE if lr is None else lr
    grads, tot, aux = gradients(p, batch, cfg, trainable, quant)
    if cfg.OPTIMIZER == "SGD":
        newp, norm = sgd_step(p, state, grads, lr, cfg.LEARNING_MOMENTUM, cfg.GRADIENT_CLIP_NORM)
    else:
        eps = 1e-4 if getattr(cfg, "F16", False) else 1e-7        # K.epsilon() (net.py:593)
        newp, norm = amsgrad_step(p, state, grads, lr, cfg.GRADIENT_CLIP_NORM, eps=eps)
    return newp, {"total": tot, "loc_loss": aux[2], "ori_loss": aux[3], "grad_norm": norm, "grads": grads,
                  "loc": aux[0], "ori": aux[1]}


# --------------------------------------------------------------------------------------
# host-side pieces around the path
# --------------------------------------------------------------------------------------
MEAN_PIXEL = (123.7, 116.8, 103.9)  # config.py:81


def mold_image(images_u8, dtype=torch.float64):
    """net.py:1337-1348: float(image) - MEAN_PIXEL."""
    return images_u8.to(dtype) - torch.tensor(MEAN_PIXEL, dtype=dtype)


def clr_triangular(it: int, base_lr: float, max_lr: float, step_size: float) -> float:
    """clr_callback.py:104-111, mode='triangular'."""
    cycle = math.floor(1 + it / (2 * step_size))
    x = abs(it / step_size - 2 * cycle + 1)
    return base_lr + (max_lr - base_lr) * max(0.0, 1 - x)


def angular_error_deg(q_est, q_gt):
    """pose_estimator.py:434: 2*acos|q.q| in degrees."""
    d = min(1.0, abs(float((q_est * q_gt).sum())))
    return 2 * math.degrees(math.acos(d))
