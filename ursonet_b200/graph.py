"""Layer graph of UrsoNet as a flat list of convolution ops + head description.

Restates the Keras graph of the reference (net.py:161-199 deep ResNet, net.py:216-282 shallow ResNet,
net.py:639-643 bottleneck conv, net.py:288-352 heads) as data: every Conv2D becomes a `ConvSpec` carrying its Keras
layer name, the BatchNorm that follows it, the activation buffer it reads, the buffer it writes and the epilogue
(ReLU / residual addend).  Layer and weight names are the reference's, so weights can be exchanged by name.
"""
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple, Union


@dataclass
class ConvSpec:
    name: str                      # Keras Conv2D layer name
    bn: Optional[str]              # Keras BatchNorm layer name (frozen BN folded into the conv) or None
    bias: bool
    k: int
    stride: int
    padding: Union[str, int]       # 'same' (TF), 'valid', or explicit symmetric ZeroPadding2D amount
    cin: int
    cout: int
    src: str                       # activation buffer read
    dst: str                       # activation buffer written
    relu: bool
    addend: Optional[str] = None   # activation buffer added before the ReLU (residual)
    stem: bool = False             # 7x7/s2 on the raw image (staged input)
    out_fp32: bool = False


@dataclass
class DenseSpec:
    name: str
    cin: int
    cout: int
    act: int                       # 0 linear, 1 relu
    src: str
    dst: str


@dataclass
class Graph:
    convs: List[ConvSpec]
    pool_src: str                  # buffer max-pooled into 'pool1'
    dense: List[DenseSpec]
    shapes: Dict[str, Tuple[int, int, int]]   # buffer name -> (H, W, C)
    relu_buffers: set              # buffers that are outputs of a ReLU (their gradient is masked by value > 0)
    nr_features: int
    loc_out: str
    ori_out: str
    ori_mode: str                  # 'classification' | 'quaternion'
    loc_mode: str                  # 'regression' | 'classification'


def _deep_blocks(backbone: str):
    n4 = {"resnet50": 5, "resnet101": 22}[backbone]
    blocks = [(2, "a", True, 1, (64, 64, 256)), (2, "b", False, 1, (64, 64, 256)), (2, "c", False, 1, (64, 64, 256)),
              (3, "a", True, 2, (128, 128, 512))]
    blocks += [(3, c, False, 1, (128, 128, 512)) for c in "bcd"]
    blocks += [(4, "a", True, 2, (256, 256, 1024))]
    blocks += [(4, chr(98 + i), False, 1, (256, 256, 1024)) for i in range(n4)]       # net.py:188-190
    blocks += [(5, "a", True, 2, (512, 512, 2048)), (5, "b", False, 1, (512, 512, 2048)),
               (5, "c", False, 1, (512, 512, 2048))]
    return blocks


def build_graph(cfg) -> Graph:
    if getattr(cfg, "REGRESS_KEYPOINTS", False):
        raise NotImplementedError("REGRESS_KEYPOINTS (experimental branch, net.py:309-313) is not built")
    if cfg.TRAIN_BN is not False:
        raise NotImplementedError("only frozen BatchNorm (config.TRAIN_BN=False, the CLI's only mode) is built")
    H, W = int(cfg.IMAGE_SHAPE[0]), int(cfg.IMAGE_SHAPE[1])
    if H % 64 or W % 64:
        raise Exception("Image size must be dividable by 2 at least 6 times "
                        "to avoid fractions when downscaling and upscaling."
                        "For example, use 256, 320, 384, 448, 512, ... etc. ")      # net.py:596-600
    if cfg.NR_IMAGE_CHANNELS != 3:
        raise NotImplementedError("NR_IMAGE_CHANNELS must be 3")
    convs: List[ConvSpec] = []
    shapes: Dict[str, Tuple[int, int, int]] = {}
    relu_bufs = set()

    def add(spec: ConvSpec, h, w):
        convs.append(spec)
        shapes[spec.dst] = (h, w, spec.cout)
        if spec.relu:
            relu_bufs.add(spec.dst)

    deep = cfg.BACKBONE in ("resnet50", "resnet101")
    if not deep and cfg.BACKBONE not in ("resnet18", "resnet34"):
        raise ValueError(f"unknown backbone {cfg.BACKBONE}")
    stem_name, stem_bn = ("conv1", "bn_conv1") if deep else ("conv0", "bn_conv0")
    h, w = H // 2, W // 2
    add(ConvSpec(stem_name, stem_bn, deep, 7, 2, 3, 3, 64, "image", stem_name + "_relu", True, stem=True), h, w)
    h, w = h // 2, w // 2
    shapes["pool1"] = (h, w, 64)
    x, c = "pool1", 64
    if deep:
        for stage, blk, has_sc, stride, (f1, f2, f3) in _deep_blocks(cfg.BACKBONE):
            cb, bb = f"res{stage}{blk}_branch", f"bn{stage}{blk}_branch"
            h, w = h // stride, w // stride
            out = f"res{stage}{blk}_out"
            add(ConvSpec(cb + "2a", bb + "2a", True, 1, stride, "valid", c, f1, x, cb + "2a", True), h, w)
            add(ConvSpec(cb + "2b", bb + "2b", True, 3, 1, "same", f1, f2, cb + "2a", cb + "2b", True), h, w)
            if has_sc:
                add(ConvSpec(cb + "1", bb + "1", True, 1, stride, "valid", c, f3, x, cb + "1", False), h, w)
                add(ConvSpec(cb + "2c", bb + "2c", True, 1, 1, "valid", f2, f3, cb + "2b", out, True, addend=cb + "1"), h, w)
            else:
                add(ConvSpec(cb + "2c", bb + "2c", True, 1, 1, "valid", f2, f3, cb + "2b", out, True, addend=x), h, w)
            x, c = out, f3
    else:
        reps = [2, 2, 2, 2] if cfg.BACKBONE == "resnet18" else [3, 4, 6, 3]
        for stage, rep in enumerate(reps):
            for block in range(rep):
                filt = 64 * 2 ** stage
                stride = 2 if (block == 0 and stage > 0) else 1
                post = block == 0
                base = f"stage{stage + 1}_unit{block + 1}_"
                h, w = h // stride, w // stride
                out = base + "relu2"
                sc = x
                if post:   # 1x1/s conv on the raw block input: no BN, no bias (net.py:225)
                    add(ConvSpec(base + "sc", None, False, 1, stride, "valid", c, filt, x, base + "sc", False), h, w)
                    sc = base + "sc"
                add(ConvSpec(base + "conv1", base + "bn2", False, 3, stride, 1, c, filt, x, base + "relu1", True), h, w)
                add(ConvSpec(base + "conv2", None, False, 3, 1, 1, filt, filt, base + "relu1", out, True, addend=sc), h, w)
                x, c = out, filt
    bw = int(cfg.BOTTLENECK_WIDTH)
    if bw % 32:
        raise NotImplementedError("BOTTLENECK_WIDTH must be a multiple of 32 in this build")
    h, w = h // 2, w // 2
    assert (h, w) == (H // 64, W // 64)
    add(ConvSpec("bottleneck_layer", None, True, 3, 2, "same", c, bw, x, "bottleneck_layer", False, out_fp32=True), h, w)
    nr_features = int(bw * H * W / 64 ** 2)                                          # net.py:640
    dense: List[DenseSpec] = []
    outs = {}
    for branch in ("loc", "ori"):
        src, f = "bottleneck_layer", nr_features
        for i in range(cfg.NR_DENSE_LAYERS):
            n = f"{branch}_dense_{i}"
            dense.append(DenseSpec(n, f, cfg.BRANCH_SIZE, 1, src, n))
            src, f = n, cfg.BRANCH_SIZE
        outs[branch] = (src, f)
    src, f = outs["loc"]
    if cfg.REGRESS_LOC:
        dense.append(DenseSpec("loc_final", f, 3, 0, src, "loc_final"))
        loc_mode = "regression"
    else:
        dense.append(DenseSpec("loc_final", f, cfg.LOC_BINS_PER_DIM ** 3, 1, src, "loc_final"))
        loc_mode = "classification"
    src, f = outs["ori"]
    if cfg.REGRESS_ORI:
        if cfg.ORIENTATION_PARAM != "quaternion":
            raise NotImplementedError("only the quaternion parameterisation of --regress_ori is built")
        dense.append(DenseSpec("ori_q", f, 4, 0, src, "ori_q"))
        ori_out, ori_mode = "ori_q", "quaternion"
    else:
        dense.append(DenseSpec("ori_final", f, cfg.ORI_BINS_PER_DIM ** 3, 1, src, "ori_final"))    # ReLU'd logits
        ori_out, ori_mode = "ori_final", "classification"
    return Graph(convs, stem_name + "_relu", dense, shapes, relu_bufs, nr_features, "loc_final", ori_out, ori_mode,
                 loc_mode)


def weight_entries(g: Graph):
    """[(keras weight name, shape, trainable, regularised)] in creation order.  Kernels are HWIO, Dense [in,out]."""
    out = []
    for c in g.convs:
        out.append((c.name + "/kernel", (c.k, c.k, c.cin, c.cout), True, True))
        if c.bias:
            out.append((c.name + "/bias", (c.cout,), True, True))
        if c.bn:
            out.append((c.bn + "/gamma", (c.cout,), True, False))      # net.py:1008-1011 skips gamma / beta
            out.append((c.bn + "/beta", (c.cout,), True, False))
            out.append((c.bn + "/moving_mean", (c.cout,), False, False))
            out.append((c.bn + "/moving_variance", (c.cout,), False, False))
    for d in g.dense:
        out.append((d.name + "/kernel", (d.cin, d.cout), True, True))
        out.append((d.name + "/bias", (d.cout,), True, True))
    return out


def backward_groups(g: Graph, sparse_bwd: bool = True):
    """The input-gradient launches of the backward plan as data (what Engine._build_backward builds, without a device):
    one group per activation buffer X, in processing (reverse production) order:
      dict(X, convs=[ConvSpec consumers], add=name of the buffer whose gradient arrives through an identity shortcut,
           mask=bool (X is a ReLU output, or pool1), colsum=bool (the producer of X needs d beta / d bias),
           sparse_in=bool (all consumers' output gradients live on even-even pixels only), stride=effective stride,
           only_phase0=bool).  Also returns the set of buffers whose own gradient is sparse."""
    producers = {c.dst: c for c in g.convs}
    cons_conv, cons_add = {}, {}
    for c in g.convs:
        cons_conv.setdefault(c.src, []).append(c)
        if c.addend:
            cons_add.setdefault(c.addend, []).append(c)
    order = [c.dst for c in g.convs]
    order.insert(1, "pool1")
    sparse, groups = set(), []
    for X in reversed(order):
        if X == "bottleneck_layer" or X == g.pool_src:
            continue
        convs, adds = cons_conv.get(X, []), cons_add.get(X, [])
        if not convs and len(adds) == 1 and X not in g.relu_buffers:
            continue           # linear shortcut branch: aliases its consumer's gradient
        prod = producers.get(X)
        need_cs = prod is not None and bool(prod.bias or prod.bn or prod.addend in
                                            [c.dst for c in g.convs if not c.relu and (c.bias or c.bn)])
        sparse_in = sparse_bwd and all(c.dst in sparse and c.stride == 1 for c in convs)
        if sparse_in and all(c.k == 3 for c in convs) and g.shapes[X][2] <= 64:
            sparse_in = False          # the dense halo / resident-weight launch is faster for the 64-channel 3x3
        stride = convs[0].stride * (2 if sparse_in else 1)
        only_phase0 = stride == 2 and all(c.k == 1 for c in convs)
        h, w, _ = g.shapes[X]
        groups.append(dict(X=X, convs=convs, add=adds[0].dst if adds else None,
                           mask=(X in g.relu_buffers or X == "pool1"), colsum=need_cs, sparse_in=sparse_in,
                           stride=stride, only_phase0=only_phase0))
        if sparse_bwd and not adds and only_phase0 and h % 2 == 0 and w % 2 == 0:
            sparse.add(X)
    return groups, sparse
