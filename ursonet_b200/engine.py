"""Execution engine: static plan of C-ABI launches for the UrsoNet forward / backward / update on one B200.

Replaces the reference's L1/L0 (Keras `fit_generator`/`predict` -> TF Session.run, net.py:1152,1241-1251): the graph
(`graph.build_graph`) is lowered ONCE into flat launch lists over pre-allocated device buffers; every launch goes
through liburso_b200.so (no autograd graph, no torch ops on the data path except memset).  The lists are then
captured into CUDA graphs, so a train step is two graph replays around one NCCL all-reduce of the flat gradient arena.

Memory layout (HBM):
  params / grads / optimizer state : flat fp32 arenas, every tensor 256-element aligned, Keras layouts (HWIO, [in,out])
  activations                      : bf16 NHWC, one buffer per conv output (kept for backward), fp32 for the heads
  gradients of activations (du)    : bf16 NHWC, masked by the consumer's ReLU at production time
  staged GEMM operands             : bf16 K-major weight matrices with the frozen-BN scale folded in (rebuilt per step)
"""
import math
import re
from typing import Dict, List, Optional

import numpy as np
import torch

from . import convplan as P
from . import lib
from .graph import ConvSpec, Graph, build_graph, weight_entries

BN_EPS = 1e-3      # keras BatchNormalization default (net.py:60)
ALIGN = 256        # arena alignment in elements (== optimizer chunk)


def _align(n):
    return (n + ALIGN - 1) // ALIGN * ALIGN


class ParamStore:
    """Flat fp32 arenas for trainable weights (+ grads) and BN moving statistics, addressed by Keras weight names."""

    def __init__(self, graph: Graph, device, weight_decay: float):
        self.entries = weight_entries(graph)
        self.device = device
        off_t = off_s = 0
        self.index: Dict[str, tuple] = {}
        for name, shape, trainable, reg in self.entries:
            n = int(np.prod(shape))
            if trainable:
                self.index[name] = ("t", off_t, shape, reg)
                off_t += _align(n)
            else:
                self.index[name] = ("s", off_s, shape, False)
                off_s += _align(n)
        self.n_train, self.n_stats = off_t, max(off_s, ALIGN)
        self.flat = torch.zeros(self.n_train, dtype=torch.float32, device=device)
        self.stats = torch.zeros(self.n_stats, dtype=torch.float32, device=device)
        nchunks = self.n_train // ALIGN
        coef = torch.zeros(nchunks, dtype=torch.float32)
        self.chunk_layer: List[str] = [""] * nchunks
        for name, (kind, off, shape, reg) in self.index.items():
            if kind != "t":
                continue
            n = int(np.prod(shape))
            c0, c1 = off // ALIGN, (off + _align(n)) // ALIGN
            if reg:   # d/dw [wd * sum(w^2) / size(w)] = 2 wd w / size(w)   (net.py:1008-1012)
                coef[c0:c1] = 2.0 * weight_decay / n
            layer = name.split("/")[0]
            for c in range(c0, c1):
                self.chunk_layer[c] = layer
        self.chunk_coef = coef.to(device)
        self.chunk_lr = torch.ones(nchunks, dtype=torch.float32, device=device)

    def view(self, name, arena=None):
        kind, off, shape, _ = self.index[name]
        base = arena if arena is not None else (self.flat if kind == "t" else self.stats)
        return base[off:off + int(np.prod(shape))].view(*shape)

    def has(self, name):
        return name in self.index

    def names(self):
        return [e[0] for e in self.entries]

    def init_keras_defaults(self, seed=0, pretrained_like=False):
        """Glorot-uniform kernels, zero biases, BN gamma=1 beta=0 mean=0 var=1 (what `--weights none` gives)."""
        g = torch.Generator().manual_seed(seed)
        for name, shape, _t, _r in self.entries:
            if name.endswith("/kernel"):
                if len(shape) == 4:
                    fan_in, fan_out = shape[0] * shape[1] * shape[2], shape[0] * shape[1] * shape[3]
                else:
                    fan_in, fan_out = shape
                lim = math.sqrt(6.0 / (fan_in + fan_out))
                w = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * lim
            elif name.endswith("/gamma") or name.endswith("/moving_variance"):
                w = torch.ones(shape, dtype=torch.float64)
                if pretrained_like:
                    w = 0.5 + torch.rand(shape, generator=g, dtype=torch.float64)
            else:
                w = torch.zeros(shape, dtype=torch.float64)
                if pretrained_like:
                    w = 0.1 * torch.randn(shape, generator=g, dtype=torch.float64)
            self.view(name).copy_(w.to(torch.float32))

    def state_dict(self):
        return {name: self.view(name).detach().cpu().numpy().copy() for name in self.names()}

    def load_state_dict(self, sd, by_name=True, exclude=None):
        """By-name load with an exclude list of LAYER names, like net.py:816-852. Returns the names loaded."""
        loaded = []
        for name in self.names():
            layer = name.split("/")[0]
            if exclude and layer in exclude:
                continue
            if name in sd:
                w = torch.as_tensor(np.asarray(sd[name]), dtype=torch.float32)
                if tuple(w.shape) != tuple(self.index[name][2]):
                    raise ValueError(f"shape mismatch for {name}: {tuple(w.shape)} vs {self.index[name][2]}")
                self.view(name).copy_(w)
                loaded.append(name)
        return loaded

    def set_trainable(self, layer_regex: str):
        """chunk_lr[c] = 1 where the owning layer's name fully matches the regex (net.py:1030-1066)."""
        mask = torch.tensor([1.0 if (l and re.fullmatch(layer_regex, l)) else 0.0 for l in self.chunk_layer])
        self.chunk_lr.copy_(mask)
        return int(mask.sum().item())


class OpRec:
    """One entry of a launch list: the callable plus what it is (for profiling / launch accounting), the LANE (CUDA
    stream) it is issued on and the ops of other lanes it must wait for (-> CUDA events / graph edges)."""
    __slots__ = ("fn", "kind", "name", "flops", "bytes", "launches", "lane", "after", "event", "signal", "segment")

    def __init__(self, fn, kind, name, flops=0.0, nbytes=0.0, launches=1, lane=0, after=()):
        self.fn, self.kind, self.name, self.flops, self.bytes, self.launches = fn, kind, name, flops, nbytes, launches
        self.lane, self.after, self.event, self.signal = lane, [a for a in after if a is not None], None, False
        self.segment = 0     # 1: the part of backward replayed after the first (overlapped) gradient all-reduce
        for a in self.after:
            a.signal = True

    def __call__(self):
        self.fn()


class Engine:
    """One model replica on one GPU. `training=True` also allocates gradient buffers and builds the backward plan."""

    def __init__(self, cfg, batch_size: int, training: bool, device="cuda", world_size: int = 1, seed: int = 0,
                 parity=None, reserve_sms: int = 0, lanes: bool = True, wgrad_lanes: int = 1, sparse_bwd: bool = True,
                 mask_bits: bool = True, pair_l2: bool = False, stage_split: int = 4):
        lib.load()   # fail loudly if the CUDA extension is missing: there is no other path
        if not torch.cuda.is_available():
            raise lib.UrsoError("a CUDA device is required (no CPU fallback)")
        self.cfg, self.B, self.training, self.device, self.world = cfg, int(batch_size), training, device, world_size
        self.parity = bool(getattr(cfg, "PARITY_MODE", False)) if parity is None else bool(parity)
        if self.parity and training:
            raise NotImplementedError("PARITY_MODE (split-bf16 operands) is forward-only")
        # reserve_sms > 0: the convolution launches of the SECOND backward segment (the part that runs while the gradient
        # all-reduce of the arena tail is in flight, see train_step(allreduce_async=...)) are planned with that many SMs
        # left free, so that NCCL's kernel has somewhere to run: the persistent conv CTAs otherwise hold every SM
        self.reserve_sms = int(reserve_sms)
        # pair_l2: the HBM-bound 1x1 "expand" convolutions of stages 2-3 (64 -> 256, 128 -> 512 channels) have a weight
        # gradient and an input gradient that BOTH stream the same large output gradient (629 / 314 MB at the bench shape).
        # The two launches already sit next to each other on two lanes, but each wants every SM, so they serialise and the
        # second one re-reads the tensor from HBM.  Paired, each is planned on half of the SMs and they start together:
        # they run side by side and the follower finds the tensor in L2.
        self.pair_l2 = bool(pair_l2) and training
        # the weight operands of the first stage_split convolutions are staged by their own small launch, so that the stem
        # does not wait for the operands of all layers at the start of a step (0 = one table)
        self.stage_split = int(stage_split)
        self._pair = {}
        self.graph: Graph = build_graph(cfg)
        # Lanes (CUDA streams -> graph branches).  Lane 0 is the dependent chain (forward convs, heads, losses, dgrad
        # chain); the weight-gradient launches run on their own lane(s) (wgrad of a layer only needs the du its dgrad
        # predecessor produced, so it overlaps the dgrad chain below it and fills its tail waves); the small per-layer
        # kernels (BN fold, weight staging, parameter gradients) run on the last, "aux" lane.  lanes=False puts
        # everything on one stream.  (Batch slices -- independent half-batch chains -- were measured slower in round 1:
        # smaller kernels lose more to wave quantisation than the overlap returns; removed.)
        multi = (not self.parity) and bool(lanes)
        # wgrad_lanes = number of weight-gradient lanes (0: wgrad stays in the main chain)
        self.wgrad_lanes = int(wgrad_lanes) if (multi and training) else 0
        self.aux_lane = (1 + self.wgrad_lanes) if multi else 0
        self._wgrad_rr = 0
        self._lane_streams = None
        self.sparse_bwd = training and bool(sparse_bwd)
        self.use_mask_bits = bool(mask_bits)     # False: the bf16 activation itself is the ReLU mask of backward (A/B runs)
        self.sparse = set()      # buffers whose gradient lives on the even-even pixels only
        self.H, self.W = int(cfg.IMAGE_SHAPE[0]), int(cfg.IMAGE_SHAPE[1])
        self.params = ParamStore(self.graph, device, cfg.WEIGHT_DECAY)
        self.params.init_keras_defaults(seed)
        self.n_launches = {"stage": 0, "fwd": 0, "bwd": 0, "update": 0}
        self._keep = []          # plan objects / index tensors kept alive
        self._zero_specs = []    # (name, numel) carved from the zero arena
        self._alloc()
        if self.parity:
            self._build_forward_parity()
        else:
            self._build_forward()
        if training:
            self._build_backward()
            self._build_update()
        self._finalise_zero_arena()
        self._graphs = {}

    # ------------------------------------------------------------------ buffers
    def _new(self, shape, dtype=torch.bfloat16):
        return torch.zeros(*shape, dtype=dtype, device=self.device)

    def _alloc(self):
        g, B = self.graph, self.B
        self.img_u8 = self._new((B, self.H, self.W, 3), torch.uint8)
        self.img_f32 = None      # allocated on demand (reference-style molded fp32 input)
        self.mean3 = torch.tensor(np.asarray(self.cfg.MEAN_PIXEL), dtype=torch.float32, device=self.device)
        self.E = self._new((B, self.H // 2 + 3, self.W // 2 + 3, 16))    # compact space-to-depth staging of the image
        self.act: Dict[str, torch.Tensor] = {}
        for name, (h, w, c) in g.shapes.items():
            f32 = name == "bottleneck_layer" or self.parity       # parity mode: fp32 master + (hi, lo) bf16 pair
            self.act[name] = self._new((B, h, w, c), torch.float32 if f32 else torch.bfloat16)
        if self.parity:
            self.E_lo = torch.zeros_like(self.E)
            self.act_hi = {n: self._new((B, h, w, c)) for n, (h, w, c) in g.shapes.items() if n != "bottleneck_layer"}
            self.act_lo = {n: self._new((B, h, w, c)) for n, (h, w, c) in g.shapes.items() if n != "bottleneck_layer"}
        self.head: Dict[str, torch.Tensor] = {}
        for d in g.dense:
            self.head[d.name] = None   # carved from the zero arena (split-K atomics accumulate into them)
            self._zero_specs.append(("head:" + d.name, B * d.cout))
        self.ori_q = self._new((B, 4), torch.float32) if g.ori_mode == "quaternion" else None
        n_ori = g.dense[-1].cout
        self.gt_loc = self._new((B, 3 if g.loc_mode == "regression" else self.cfg.LOC_BINS_PER_DIM ** 3), torch.float32)
        self.gt_ori = self._new((B, 4 if g.ori_mode == "quaternion" else n_ori), torch.float32)
        self.losses = self._new((2,), torch.float32)        # [loc_loss * w, ori_loss * w]
        self.scale: Dict[str, torch.Tensor] = {}
        self.shift: Dict[str, torch.Tensor] = {}
        for c in g.convs:
            self.scale[c.name] = self._new((c.cout,), torch.float32)
            self.shift[c.name] = self._new((c.cout,), torch.float32)
        if self.training:
            self.grads = torch.zeros_like(self.params.flat)
            self.dact: Dict[str, torch.Tensor] = {}
            self.dhead: Dict[str, torch.Tensor] = {}
            self.hyper = self._new((8,), torch.float32)
            self.sumsq = self._new((2048,), torch.float32)     # [0] = sum g^2, rest: per-block partials (URSO_SUMSQ_SCRATCH)
            self.opt_state = [torch.zeros_like(self.params.flat) for _ in range(1 if self.cfg.OPTIMIZER == "SGD" else 3)]
            self.opt_t = 0

    def _zero_view(self, key):
        off, n = self._zero_index[key]
        return self.zero_arena[off:off + n]

    def _finalise_zero_arena(self):
        """All buffers that must be zero at the start of a step (atomic accumulators) live in ONE arena -> one memset."""
        off = 0
        self._zero_index = {}
        for key, n in self._zero_specs:
            self._zero_index[key] = (off, n)
            off += _align(n)
        self.zero_arena = torch.zeros(max(off, ALIGN), dtype=torch.float32, device=self.device)
        for fn in self._late_binds:
            fn()
        if self.parity:
            return
        # multi-tensor job tables (one launch each for: BN fold, weight-operand staging, parameter gradients per segment)
        self._bn_table = lib.BnFoldTable(self._bn_jobs, self.device)
        # two tables: the operands of the first layers (tiny) so that the stem can start at once, and everything else,
        # which is staged while the stem runs
        early = [op for n, op in self.fwd_ops.items() if n in self._early_layers]
        late = [op for n, op in self.fwd_ops.items() if n not in self._early_layers]
        self._stage_table_a = lib.StageTable(early, [], self.device)
        self._stage_table = lib.StageTable(late, [b["p"] for b in self._dgrad_boxes], self.device)
        self.ops_pgrad = []
        if self.training:
            for seg in (0, 1):
                specs = [(mk, nb) for op, mk, nb in self._pgrad_specs if op.segment == seg]
                if not specs:
                    continue
                table = lib.PgradTable([mk() for mk, _ in specs], self.device)
                op = OpRec(lambda table=table: table.launch(BN_EPS), "param_grads", "segment %d" % seg, 0.0,
                           sum(nb for _, nb in specs), launches=2)
                op.segment = seg
                self.ops_pgrad.append(op)

    def _run_stage_early(self):
        self._bn_table.launch(BN_EPS)
        self._stage_table_a.launch()

    def _run_stage_tables(self):
        self._stage_table.launch()

    def _run_pgrad(self, segments=(0, 1)):
        """Parameter gradients of all convolutions from the raw wgrads (main stream, after the lanes have joined)."""
        for op in self.ops_pgrad:
            if op.segment in segments:
                op()

    # ------------------------------------------------------------------ plan helpers
    def _idx(self, values):
        t = torch.tensor(list(values), dtype=torch.int32, device=self.device)
        self._keep.append(t)
        return t

    def _conv_weight_ptrs(self, c: ConvSpec):
        p = self.params
        w = p.view(c.name + "/kernel")
        bias = p.view(c.name + "/bias") if c.bias else None
        if c.bn:
            bn = [p.view(f"{c.bn}/{k}") for k in ("gamma", "beta", "moving_mean", "moving_variance")]
        else:
            bn = [None] * 4
        return w, bias, bn

    def _add(self, lst, op):
        """Append an op to a launch list, remembering the last op of its lane (join points depend on it)."""
        lst.append(op)
        self._last[op.lane] = op
        return op

    def _conv_shape(self, c: ConvSpec):
        """urso_conv2d_shape of a graph conv (the stem is the ksize-7 operator over the staged tensor E)."""
        if c.stem:
            return lib.conv_shape(self.B, self.H, self.W, 3, c.cout, 7, 2, 3)
        h, w, _ = self.graph.shapes[c.src]
        return lib.conv_shape(self.B, h, w, c.cin, c.cout, c.k, c.stride, c.padding)

    def _build_forward(self):
        """Forward launch list.  Every convolution is ONE C-ABI operator (urso_conv2d_fwd_*): the library plans the
        K-segments / parity views / tiles and owns the weight-operand layout; Python only wires tensors."""
        g, B = self.graph, self.B
        S = lib.stream_ptr
        self.ops_stage, self.ops_fwd, self.ops_loss = [], [], []
        self._late_binds = []
        self._last = {}
        self.fwd_ops: Dict[str, lib.Conv2dFwd] = {}
        # BN fold + weight-operand staging of ALL layers: two multi-tensor launches (job tables built in
        # _finalise_zero_arena, once the gradient operators exist too) instead of ~170 tiny per-layer launches
        self._bn_jobs, self._dgrad_boxes, self._pgrad_specs = [], [], []
        self._early_layers = set(c.name for c in g.convs[:self.stage_split])     # stem + the first block's convs
        self._stage_op_a = self._add(self.ops_stage, OpRec(self._run_stage_early, "stage", "bn fold + first layers",
                                                           launches=2, lane=self.aux_lane))
        self._stage_op = self._add(self.ops_stage, OpRec(self._run_stage_tables, "stage", "all other layers", launches=1,
                                                         lane=self.aux_lane))
        self.relu_bits: Dict[str, torch.Tensor] = {}
        aux = self.aux_lane
        for c in g.convs:
            w, bias, bn = self._conv_weight_ptrs(c)
            sc, sh = self.scale[c.name], self.shift[c.name]
            self._bn_jobs.append((bn[0], bn[1], bn[2], bn[3], bias, sc, sh, c.cout))
            shape = self._conv_shape(c)
            oh, ow = lib.out_hw(shape)
            assert (oh, ow) == tuple(g.shapes[c.dst][:2]), (c.name, oh, ow, g.shapes[c.dst])
            x = self.E if c.stem else self.act[c.src]
            out = self.act[c.dst]
            addend = self.act[c.addend] if c.addend else None
            # training: the ReLU mask backward needs is written bit-packed by the epilogue (1 bit per element; the
            # gradient launches then read 1/16 of the bytes of the bf16 activation and need no smem ring for it)
            bits = None
            if self.training and self.use_mask_bits and c.relu and c.dst in g.relu_buffers and c.dst != g.pool_src:
                bits = self.relu_bits[c.dst] = torch.zeros((B, oh, ow, c.cout // 32), dtype=torch.int32, device=self.device)
            op = lib.Conv2dFwd(shape, x, w, sc, sh, out, addend=addend, relu=c.relu, relu_bits=bits)
            self.fwd_ops[c.name] = op
            staged = self._stage_op_a if c.name in self._early_layers else self._stage_op
            if c.stem:
                ph, pw, _ = g.shapes[c.dst]
                self.argmax = self._new((B, ph // 2, pw // 2, 64), torch.uint8) if self.training else None
            fl = 2.0 * B * oh * ow * c.cout * c.k * c.k * c.cin
            # algorithmic bytes: a strided 1x1 conv only touches the pixels it samples
            in_elems = self.E[0].numel() if c.stem else \
                g.shapes[c.src][0] * g.shapes[c.src][1] * c.cin // (c.stride * c.stride if c.k == 1 else 1)
            nb = 2.0 * B * in_elems + (4.0 if c.out_fp32 else 2.0) * B * oh * ow * c.cout * (2 if c.addend else 1) \
                + 2.0 * c.cout * c.k * c.k * c.cin
            self._add(self.ops_fwd, OpRec(op.launch, "conv_fwd", c.name, fl, nb, after=[staged]))
            if c.stem:
                src, dst, amax = self.act[c.dst], self.act["pool1"], self.argmax
                self._add(self.ops_fwd, OpRec(lambda src=src, dst=dst, amax=amax, ph=ph, pw=pw: lib.call(
                    "urso_maxpool_fwd", src.data_ptr(), dst.data_ptr(), lib.ptr(amax), B, ph, pw, 64, S()),
                    "pool_fwd", "pool1", 0.0, 2.0 * B * ph * pw * 64 * 1.25))
        self._build_heads_forward()

    def _build_heads_forward(self):
        g, B = self.graph, self.B
        S = lib.stream_ptr
        # ---- heads: fp32 Dense layers on the flattened NHWC bottleneck output (net.py:298,332).  The two branches are
        # independent: the location branch runs on lane 1 (idle between forward and backward), the orientation branch on
        # lane 0 -- the heads sit on the critical path between the two conv stacks with every SM otherwise idle
        self.head_lane = 1 if self.aux_lane else 0
        feat_op = self._last.get(0)
        for d in g.dense:
            w, b = self.params.view(d.name + "/kernel"), self.params.view(d.name + "/bias")
            lane = self.head_lane if d.name.startswith("loc") else 0

            def bind(d=d):
                self.head[d.name] = self._zero_view("head:" + d.name).view(self.B, d.cout)
            self._late_binds.append(bind)

            def run(d=d, w=w, b=b):
                x = self.act["bottleneck_layer"] if d.src == "bottleneck_layer" else self.head[d.src]
                y = self.head[d.name]
                lib.call("urso_dense_fwd", x.data_ptr(), w.data_ptr(), y.data_ptr(), self.B, d.cin, d.cout, S())
                lib.call("urso_dense_bias_act", y.data_ptr(), b.data_ptr(), self.B, d.cout, d.act, S())
            self._add(self.ops_fwd, OpRec(run, "dense_fwd", d.name, 2.0 * B * d.cin * d.cout, 4.0 * d.cin * d.cout, 2,
                                          lane=lane, after=[feat_op] if (lane and d.src == "bottleneck_layer") else ()))
        if g.ori_mode == "quaternion":   # inference output is the normalised quaternion (net.py:345-346)
            self._add(self.ops_fwd, OpRec(lambda: lib.call("urso_quat_head", self.head["ori_q"].data_ptr(), None,
                                                           self.ori_q.data_ptr(), None, None, self.B, 1.0, S()),
                                          "misc", "quat_head"))

    def _build_forward_parity(self):
        """Forward plan in split-bf16 precision (cfg.PARITY_MODE): every activation x is carried as fp32 plus the pair
        hi = bf16(x), lo = bf16(x - hi); a conv is three K-segment groups of Engine F, A_hi.W_hi + A_hi.W_lo + A_lo.W_hi,
        accumulated in fp32 in TMEM; the epilogue adds the folded BN shift and writes fp32; urso_split_f32 then applies
        the residual add + ReLU in fp32 and regenerates the (hi, lo) pair.  Same launch machinery, ~3x the MMA work."""
        g, B = self.graph, self.B
        S = lib.stream_ptr
        self.ops_stage, self.ops_fwd, self.ops_loss = [], [], []
        self._late_binds = []
        self._last = {}
        self.Bf = {}
        for c in g.convs:
            w, bias, bn = self._conv_weight_ptrs(c)
            sc, sh = self.scale[c.name], self.shift[c.name]
            self.ops_stage.append(lambda bn=bn, bias=bias, sc=sc, sh=sh, c=c: lib.call(
                "urso_bn_fold", lib.ptr(bn[0]), lib.ptr(bn[1]), lib.ptr(bn[2]), lib.ptr(bn[3]), lib.ptr(bias), BN_EPS,
                sc.data_ptr(), sh.data_ptr(), c.cout, S()))
            if c.stem:
                segs, idx = P.stem_segments(), P.stem_weight_index(3)
                views_hi, views_lo = [lib.stem_view(self.E)], [lib.stem_view(self.E_lo)]
            else:
                h, w_, _ = g.shapes[c.src]
                geom = P.make_geom(c.k, c.stride, c.padding, c.cin, c.cout, h, w_)
                segs, idx = P.fwd_segments(geom)
                views_hi = P.input_views(self.act_hi[c.src], c.stride)
                views_lo = P.input_views(self.act_lo[c.src], c.stride)
            K, nv = len(idx), len(views_hi)
            bmat = self._new((c.cout, 3 * K))
            self.Bf[c.name] = bmat
            idx_d = self._idx(idx)
            for col, part in ((0, 0), (K, 1), (2 * K, 0)):        # [W_hi | W_lo | W_hi]
                dst = bmat[:, col:]
                self.ops_stage.append(lambda w=w, sc=sc, dst=dst, idx_d=idx_d, K=K, c=c, part=part: lib.call(
                    "urso_stage_weight_rows", w.data_ptr(), sc.data_ptr(), dst.data_ptr(), idx_d.data_ptr(), K, c.cout,
                    c.cout, 3 * K, part, S()))
            segs3 = list(segs) + list(segs) + [(m + nv, dh, dw, ch) for (m, dh, dw, ch) in segs]   # A_hi, A_hi, A_lo
            out32 = self.act[c.dst]
            oh, ow = g.shapes[c.dst][0], g.shapes[c.dst][1]
            tw, th = P.pick_patch(oh, ow, 128)
            plan = lib.ConvGemm(views_hi + views_lo, bmat, segs3, out32, ow, oh, B, tw, th, shift=sh)
            self._keep.append(plan)
            self.ops_fwd.append(OpRec(plan.launch, "conv_fwd", c.name, 6.0 * B * oh * ow * c.cout * c.k * c.k * c.cin, 0.0))
            if c.dst != "bottleneck_layer":
                addend = self.act[c.addend] if c.addend else None
                hi, lo = self.act_hi[c.dst], self.act_lo[c.dst]
                self.ops_fwd.append(lambda out32=out32, addend=addend, hi=hi, lo=lo, c=c: lib.call(
                    "urso_split_f32", out32.data_ptr(), lib.ptr(addend), out32.data_ptr(), hi.data_ptr(), lo.data_ptr(),
                    out32.numel(), int(c.relu), S()))
            if c.stem:
                ph, pw, _ = g.shapes[c.dst]
                self.argmax = None
                src, dst = self.act[c.dst], self.act["pool1"]
                hi, lo = self.act_hi["pool1"], self.act_lo["pool1"]
                self.ops_fwd.append(lambda src=src, dst=dst, hi=hi, lo=lo, ph=ph, pw=pw: (
                    lib.call("urso_maxpool_fwd_f32", src.data_ptr(), dst.data_ptr(), B, ph, pw, 64, S()),
                    lib.call("urso_split_f32", dst.data_ptr(), None, None, hi.data_ptr(), lo.data_ptr(), dst.numel(), 0, S())))
        self._build_heads_forward()

    # ------------------------------------------------------------------ backward plan
    def _build_backward(self):
        g, B, cfg = self.graph, self.B, self.cfg
        S = lib.stream_ptr
        self.ops_bwd = []
        wl = float(cfg.LOSS_WEIGHTS.get("loc_loss", 1.0))
        wo = float(cfg.LOSS_WEIGHTS.get("ori_loss", 1.0))
        for d in g.dense:
            self.dhead[d.name] = self._new((B, d.cout), torch.float32)
        self.dfeat = [self._new((B, g.nr_features), torch.float32) for _ in range(2)]
        loc_l, ori_l = self.losses[0:1], self.losses[1:2]

        # ---- losses (net.py:656-669, 705-762) -> gradients w.r.t. the head outputs
        def loss_loc():
            loc, dloc = self.head["loc_final"], self.dhead["loc_final"]
            if g.loc_mode == "regression":
                lib.call("urso_rel_loss", loc.data_ptr(), self.gt_loc.data_ptr(), dloc.data_ptr(), loc_l.data_ptr(), B, 3,
                         wl, S())
            else:
                lib.call("urso_softmax_xent", loc.data_ptr(), self.gt_loc.data_ptr(), dloc.data_ptr(), loc_l.data_ptr(),
                         B, loc.shape[1], wl, S())

        def loss_ori():
            if g.ori_mode == "quaternion":
                lib.call("urso_quat_head", self.head["ori_q"].data_ptr(), self.gt_ori.data_ptr(), self.ori_q.data_ptr(),
                         self.dhead["ori_q"].data_ptr(), ori_l.data_ptr(), B, wo, S())
            else:
                z = self.head["ori_final"]
                lib.call("urso_softmax_xent", z.data_ptr(), self.gt_ori.data_ptr(), self.dhead["ori_final"].data_ptr(),
                         ori_l.data_ptr(), B, z.shape[1], wo, S())
        hl = self.head_lane
        self._add(self.ops_loss, OpRec(loss_loc, "loss", "loc_loss", lane=hl))       # behind loc_final on its lane
        self._add(self.ops_loss, OpRec(loss_ori, "loss", "ori_loss"))

        # ---- heads backward (reverse order); dx of the first layer of each branch goes to dfeat[branch]
        for d in reversed(g.dense):
            branch = 0 if d.name.startswith("loc") else 1
            gw = self.params.view(d.name + "/kernel", self.grads)
            gb = self.params.view(d.name + "/bias", self.grads)
            w = self.params.view(d.name + "/kernel")

            def run(d=d, w=w, gb=gb, branch=branch):
                # critical path (lane 0): ReLU mask + bias gradient + input gradient
                x = self.act["bottleneck_layer"] if d.src == "bottleneck_layer" else self.head[d.src]
                dx = self.dfeat[branch] if d.src == "bottleneck_layer" else self.dhead[d.src]
                lib.call("urso_dense_bwd", x.data_ptr(), w.data_ptr(), self.head[d.name].data_ptr(),
                         self.dhead[d.name].data_ptr(), dx.data_ptr(), None, gb.data_ptr(), B, d.cin, d.cout,
                         d.act, S())

            def run_w(d=d, w=w, gw=gw):
                # weight gradient: nothing downstream waits for it -> aux lane, behind the (already masked) dy
                x = self.act["bottleneck_layer"] if d.src == "bottleneck_layer" else self.head[d.src]
                lib.call("urso_dense_bwd", x.data_ptr(), w.data_ptr(), None, self.dhead[d.name].data_ptr(), None,
                         gw.data_ptr(), None, B, d.cin, d.cout, 0, S())
            op_d = self._add(self.ops_bwd, OpRec(run, "dense_bwd", d.name, 2.0 * B * d.cin * d.cout, 4.0 * d.cin * d.cout, 2,
                                                 lane=hl if branch == 0 else 0))
            if branch == 0:
                last_loc_bwd = op_d
            self._add(self.ops_bwd, OpRec(run_w, "dense_wgrad", d.name, 2.0 * B * d.cin * d.cout, 8.0 * d.cin * d.cout, 2,
                                          lane=self.aux_lane, after=[op_d]))
        assert cfg.NR_DENSE_LAYERS in range(3)      # net.py:293,327 (the CLI fixes it to 1, pose_estimator.py:820)

        # ---- gradient buffers of activations
        bw = g.shapes["bottleneck_layer"][2]
        h6, w6 = g.shapes["bottleneck_layer"][:2]
        self.dact["bottleneck_layer"] = self._new((B, h6, w6, P.ceil64(bw)))   # zero padded bf16 operand
        self._add(self.ops_bwd, OpRec(lambda: lib.call(
            "urso_pad_cast_rows", self.dfeat[0].data_ptr(), self.dfeat[1].data_ptr(),
            self.dact["bottleneck_layer"].data_ptr(), B * h6 * w6, bw, P.ceil64(bw), S()), "misc", "pad_cast",
            after=[last_loc_bwd] if hl else ()))
        producers = {c.dst: c for c in g.convs}
        cons_conv: Dict[str, List[ConvSpec]] = {}
        cons_add: Dict[str, List[ConvSpec]] = {}
        for c in g.convs:
            cons_conv.setdefault(c.src, []).append(c)
            if c.addend:
                cons_add.setdefault(c.addend, []).append(c)
        self.colsum: Dict[str, Optional[torch.Tensor]] = {}

        def colsum_for(buf):
            key = "colsum:" + buf
            if key not in [k for k, _ in self._zero_specs]:
                self._zero_specs.append((key, g.shapes[buf][2] if buf != "bottleneck_layer" else P.ceil64(bw)))
            return key

        # process buffers in reverse production order
        order = [c.dst for c in g.convs]
        order.insert(1, "pool1")
        self.dgrad_ops = {}
        self.dgrad_meta = []     # == graph.backward_groups(self.graph)[0] (checked by the tests)
        self._bwd_root = None
        self._bwd_split = None
        for X in reversed(order):
            if X == "bottleneck_layer":
                key = colsum_for(X)
                # whole batch on lane 0; every slice lane starts its backward chain after this op (and after the
                # dgrad weight staging of the aux lane)
                self._bwd_root = self._add(self.ops_bwd, OpRec(lambda key=key: lib.call(
                    "urso_colsum_bf16", self.dact["bottleneck_layer"].data_ptr(), self._zero_view(key).data_ptr(),
                    B * h6 * w6, P.ceil64(bw), S()), "misc", "colsum_bottleneck"))
                self.colsum[X] = key
            elif X == g.pool_src:     # stem output: gradient arrives through the max-pool
                h, w, c = g.shapes[X]
                self.dact[X] = self._new((B, h, w, c))
                key = colsum_for(X)
                am, dp, dx = self.argmax, self.dact["pool1"], self.dact[X]
                self._add(self.ops_bwd, OpRec(lambda am=am, dp=dp, dx=dx, h=h, w=w, c=c, key=key: lib.call(
                    "urso_maxpool_bwd", None, am.data_ptr(), dp.data_ptr(), dx.data_ptr(),
                    self._zero_view(key).data_ptr(), B, h, w, c, S()),
                    "pool_bwd", X, 0.0, 2.0 * B * h * w * c * 1.4, 1))
                self.colsum[X] = key
            else:
                convs = cons_conv.get(X, [])
                adds = cons_add.get(X, [])
                if not convs and len(adds) == 1 and X not in g.relu_buffers:
                    # linear shortcut branch: d(out)/d(sc) = 1 -> alias the consumer's gradient
                    self.dact[X] = self.dact[adds[0].dst]
                    self.colsum[X] = self.colsum[adds[0].dst]
                else:
                    prod = producers.get(X)
                    need_cs = prod is not None and bool(prod.bias or prod.bn or prod.addend in
                                                        [c.dst for c in g.convs if not c.relu and (c.bias or c.bn)])
                    self._build_dgrad_group(X, convs, adds, colsum_for, need_cs)
            if X in producers:
                self._build_wgrad(producers[X])
                # split point for the overlapped gradient all-reduce: once >= 90 % of the parameters (the arena tail:
                # heads, bottleneck, late stages) have their gradients, the rest of backward hides their all-reduce
                if self._bwd_split is None:
                    off = self.params.index[producers[X].name + "/kernel"][1]
                    if self.params.n_train - off >= 0.9 * self.params.n_train and off > 0:
                        self._bwd_split = (len(self.ops_bwd), off)
        if self._bwd_split is not None:
            for op in self.ops_bwd[self._bwd_split[0]:]:
                op.segment = 1

    def _build_dgrad_group(self, X, convs, adds, colsum_for, need_cs):
        """du_X = mask_X( sum_convs dgrad(du_conv.dst, W_conv) + sum_adds du_add.dst ): ONE C-ABI operator
        (urso_conv2d_dgrad_*; one Engine-F launch per output phase with the consumers' K ranges concatenated)."""
        g, B = self.graph, self.B
        h, w, cin = g.shapes[X]
        assert len(adds) <= 1 and convs, (X, len(adds), len(convs))
        # Structural sparsity: a buffer consumed only by 1x1/stride-2 convolutions (the Keras-v1 block puts the stride on
        # the first 1x1, net.py:138,152) has a gradient that is non-zero on the even-even pixels only, and so has
        # everything it feeds through 1x1 convolutions.  For such a consumer the gradient operators run the conv as a
        # stride-2 conv on the decimated gradient grid (dy_sparse): 4x less work for a 1x1 (and X is sparse again),
        # 9 -> 2.25 taps on average for a 3x3.  The never-written odd phases stay zero from allocation (torch.zeros), so
        # any launch that reads such a buffer densely (e.g. as the gradient fan-in addend) is still exact.
        sparse_in = self.sparse_bwd and all(c.dst in self.sparse and c.stride == 1 for c in convs)
        if sparse_in and all(c.k == 3 for c in convs) and cin <= 64:
            # measured (profiles/r02_progress.md): for the 64-channel 3x3 the dense launch (halo + resident weights + two
            # pipelines, 0.088 ms) beats the four decimated phase launches (0.104 ms); it just multiplies the zeros
            sparse_in = False
        stride = convs[0].stride * (2 if sparse_in else 1)
        assert all(c.stride == convs[0].stride for c in convs)
        assert not (adds and convs[0].stride != 1)
        dX = self.dact[X] = self._new((B, h, w, cin))
        key = colsum_for(X) if need_cs else None
        self.colsum[X] = key
        # pool1 = max of post-ReLU values: masking its gradient by (pool1 > 0) IS the stem's ReLU mask (a window's max
        # is 0 only when all its inputs are 0), so the max-pool backward does not have to read the stem output
        mask = self.act[X] if (X in g.relu_buffers or X == "pool1") else None
        mask_bits = self.relu_bits.get(X)
        if mask_bits is not None:
            mask = None
        addend = self.dact[adds[0].dst] if adds else None
        shapes = [self._conv_shape(c) for c in convs]
        dys = [self.dact[c.dst] for c in convs]
        ws = [self.params.view(c.name + "/kernel") for c in convs]
        scs = [self.scale[c.name] for c in convs]
        box = {}
        # a 1x1 consumer at (effective) stride 2 only reaches the even-even phase of dX
        only_phase0 = stride == 2 and all(c.k == 1 for c in convs)
        fl = touched = a_elems = 0.0
        for c, sh_ in zip(convs, shapes):
            oh, ow = lib.out_hw(sh_)
            if sparse_in:
                oh, ow = (oh + 1) // 2, (ow + 1) // 2
            fl += 2.0 * B * oh * ow * c.cout * c.k * c.k * c.cin      # executed work (decimated grid when sparse)
            a_elems += B * oh * ow * P.ceil64(c.cout)
        touched = dX.numel() / (4 if only_phase0 else 1)
        nbytes = 2.0 * (a_elems + touched * (1 + (mask is not None) + (addend is not None))
                        + sum(c.cout * c.k * c.k * c.cin for c in convs)) + (touched / 8 if mask_bits is not None else 0)
        stage_op = self._stage_op
        self._dgrad_boxes.append(box)
        pair = self._pair.get(convs[0].name) if (len(convs) == 1 and not adds and not sparse_in) else None
        if stride == 2 and not self.sparse_bwd:      # phases no filter tap reaches must read as zero
            self._add(self.ops_bwd, OpRec(lambda: box["p"].untouched and dX.zero_(), "fill", X, 0.0, 2.0 * dX.numel(),
                                          after=self._bwd_deps()))
        op = self._add(self.ops_bwd, OpRec(lambda: box["p"].launch(), "conv_dgrad", X, fl, nbytes,
                                           after=self._bwd_deps() + [stage_op] + ([pair] if pair is not None else [])))

        def bind():
            cs = self._zero_view(key) if key else None
            with self._cta_limit(op, half=pair is not None):
                box["p"] = lib.Conv2dDgrad(shapes, dys, ws, scs, dX, mask=mask, addend=addend, colsum=cs,
                                           dy_sparse=sparse_in, mask_bits=mask_bits)
            op.launches = box["p"].n_launches
            assert (box["p"].untouched == 0b1110) == only_phase0, (X, box["p"].untouched)
        self._late_binds.append(bind)
        self.dgrad_ops[X] = box
        self.dgrad_meta.append(dict(X=X, convs=convs, add=adds[0].dst if adds else None,
                                    mask=mask is not None or mask_bits is not None,
                                    colsum=key is not None, sparse_in=sparse_in, stride=stride, only_phase0=only_phase0))
        if self.sparse_bwd and not adds and only_phase0 and h % 2 == 0 and w % 2 == 0:
            self.sparse.add(X)

    def _cta_limit(self, op, half=False):
        """Context manager: plan the operator of a second-segment backward op with `reserve_sms` SMs left free, or (half)
        on half of the SMs (an L2-sharing pair)."""
        import contextlib

        @contextlib.contextmanager
        def cm():
            n_sm = lib.load().urso_num_sms()
            limit = (self.reserve_sms > 0 and op.segment == 1) or half
            if limit:
                n = n_sm - (self.reserve_sms if op.segment == 1 else 0)
                lib.load().urso_set_max_ctas(max(1, n // 2 if half else n))
            try:
                yield
            finally:
                if limit:
                    lib.load().urso_set_max_ctas(0)
        return cm()

    def _bwd_deps(self):
        """Cross-lane dependencies of a main-lane backward op: none (lane 0 is ordered by its stream); kept as a hook."""
        return []

    def _build_wgrad(self, c: ConvSpec):
        """Raw weight gradient (urso_conv2d_wgrad_*) on the wgrad lane + urso_conv_param_grads on the aux lane."""
        g, B = self.graph, self.B
        S = lib.stream_ptr
        shape = self._conv_shape(c)
        oh, ow = lib.out_hw(shape)
        sparse_du = (not c.stem) and self.sparse_bwd and c.dst in self.sparse and c.stride == 1
        if sparse_du:
            oh, ow = (oh + 1) // 2, (ow + 1) // 2
        n_rows = 4 * 64 if c.stem else c.k * c.k * c.cin
        row_map = self._idx(lib.stem_grad_row_map()) if c.stem else None
        gkey = "G:" + c.name
        self._zero_specs.append((gkey, n_rows * c.cout))
        skey = "S:" + c.name
        if c.bn:
            self._zero_specs.append((skey, c.cout))
        x = self.E if c.stem else self.act[c.src]
        du = self.dact[c.dst]
        box = {}

        def bind():
            with self._cta_limit(wg_op, half=paired):
                box["p"] = lib.Conv2dWgrad(shape, x, du, self._zero_view(gkey), dy_sparse=sparse_du)
        self._late_binds.append(bind)
        fl = 2.0 * B * oh * ow * c.cout * c.k * c.k * c.cin
        sub = 4 if ((c.k == 1 and c.stride == 2) or (sparse_du and c.k == 1)) else 1     # pixels a strided 1x1 samples
        nb = 2.0 * (x.numel() / sub + du.numel() / (4 if sparse_du else 1)) + 4.0 * n_rows * c.cout
        if self.wgrad_lanes:   # own lane: ordered behind the dgrad that produced du (the last op of the main lane)
            lane = 1 + self._wgrad_rr % self.wgrad_lanes
            deps = [self._last.get(0), self._bwd_root]
        else:
            lane, deps = 0, []
        paired = (self.pair_l2 and self.wgrad_lanes == 1 and not c.stem and c.k == 1 and c.stride == 1 and not sparse_du
                  and c.cin <= 128 and c.cout >= 4 * c.cin and 2.0 * du.numel() >= 2.5e8)
        if paired:   # marker on the wgrad lane: the input-gradient launch that shares du starts when this lane gets here
            self._pair[c.name] = self._add(self.ops_bwd, OpRec(lambda: None, "misc", "pair:" + c.name, launches=0, lane=lane))
        wg_op = self._add(self.ops_bwd, OpRec(lambda: box["p"].launch(), "conv_wgrad", c.name, fl, nb, lane=lane,
                                              after=deps))
        self._wgrad_rr += 1
        # parameter gradients from the raw wgrad (BN scale folded back, d gamma / d beta / d bias in closed form):
        # aux lane, after the wgrad (and with it the d-beta column sums)
        w, bias, bn = self._conv_weight_ptrs(c)
        pv = self.params.view
        dW = pv(c.name + "/kernel", self.grads)
        dbias = pv(c.name + "/bias", self.grads) if c.bias else None
        dgamma = pv(c.bn + "/gamma", self.grads) if c.bn else None
        dbeta = pv(c.bn + "/beta", self.grads) if c.bn else None
        cs_key = self.colsum.get(c.dst) if (c.bias or c.bn) else None
        sc = self.scale[c.name]
        R = c.k * c.k * c.cin

        def make_job():
            return dict(G=self._zero_view(gkey), row_map=row_map, w=w, colsum=self._zero_view(cs_key) if cs_key else None,
                        scale=sc, gamma=bn[0], mean=bn[2], var=bn[3], bias=bias, dW=dW, dbias=dbias, dgamma=dgamma,
                        dbeta=dbeta, S=self._zero_view(skey) if c.bn else None, R=R, CO=c.cout)
        self._pgrad_specs.append((wg_op, make_job, 12.0 * R * c.cout))

    def _build_update(self):
        S = lib.stream_ptr
        p = self.params
        n = p.n_train
        self.ops_update = []
        self.ops_update.append(OpRec(lambda: lib.call(
            "urso_add_reg_sumsq", self.grads.data_ptr(), p.flat.data_ptr(), p.chunk_coef.data_ptr(),
            p.chunk_lr.data_ptr(), 1.0 / self.world, self.sumsq.data_ptr(), n, S()), "update", "reg_sumsq"))
        if self.cfg.OPTIMIZER == "SGD":
            self.ops_update.append(OpRec(lambda: lib.call(
                "urso_sgd_step", p.flat.data_ptr(), self.opt_state[0].data_ptr(), self.grads.data_ptr(),
                p.chunk_lr.data_ptr(), self.sumsq.data_ptr(), self.hyper.data_ptr(), n, S()), "update", "sgd"))
        else:
            self.ops_update.append(OpRec(lambda: lib.call(
                "urso_amsgrad_step", p.flat.data_ptr(), self.opt_state[0].data_ptr(), self.opt_state[1].data_ptr(),
                self.opt_state[2].data_ptr(), self.grads.data_ptr(), p.chunk_lr.data_ptr(), self.sumsq.data_ptr(),
                self.hyper.data_ptr(), n, S()), "update", "amsgrad"))

    # ------------------------------------------------------------------ running
    def _fork(self):
        """Start of a multi-lane phase: every lane stream waits for what the current (main) stream has been given."""
        self._main = torch.cuda.current_stream()
        if self.aux_lane == 0:
            return
        if self._lane_streams is None:
            self._lane_streams = [None] + [torch.cuda.Stream(device=self.device) for _ in range(self.aux_lane)]
        for st in self._lane_streams[1:]:
            st.wait_stream(self._main)

    def _join(self):
        if self.aux_lane == 0:
            return
        for st in self._lane_streams[1:]:
            self._main.wait_stream(st)

    _serial = False      # profile_ops: issue everything on the current stream, in list order
    _split_run = False   # True while the two-graph (overlapped all-reduce) schedule is being issued / captured

    def _run(self, ops):
        if self.aux_lane == 0 or self._serial:
            for op in ops:
                op()
            return
        main, streams = self._main, self._lane_streams
        for op in ops:
            st = main if op.lane == 0 else streams[op.lane]
            for a in op.after:
                # Only in the split schedule does an earlier segment end with a full join (and live in another captured
                # graph, whose events cannot be waited on); the single-graph schedule keeps every cross-lane edge.
                if a.lane != op.lane and not (self._split_run and a.segment != op.segment):
                    st.wait_event(a.event)
            if op.lane == 0:
                op()
            else:
                with torch.cuda.stream(st):
                    op()
            if op.signal:
                if op.event is None:
                    op.event = torch.cuda.Event()
                op.event.record(st)

    def _stage_input(self):
        S = lib.stream_ptr
        if self._input_kind == "u8":
            args = (self.img_u8.data_ptr(), 1, 1, self.mean3.data_ptr())
        else:
            args = (self.img_f32.data_ptr(), 0, 0, None)
        lib.call("urso_stem_stage", *args, self.E.data_ptr(), self.B, self.H, self.W, 0, S())
        if self.parity:
            lib.call("urso_stem_stage", *args, self.E_lo.data_ptr(), self.B, self.H, self.W, 1, S())

    _input_kind = "u8"

    def set_input_kind(self, kind):
        """'u8': raw uint8 RGB, mean subtracted on device.  'molded': fp32 already mean-subtracted (net.py:1337-1348)."""
        assert kind in ("u8", "molded")
        if kind == "molded" and self.img_f32 is None:
            self.img_f32 = self._new((self.B, self.H, self.W, 3), torch.float32)
        if kind != self._input_kind:
            self._graphs = {}
        self._input_kind = kind

    def _fwd_body(self):
        self.zero_arena.zero_()
        self._stage_input()
        self._fork()
        self._run(self.ops_stage)
        self._run(self.ops_fwd)

    def _phase_fwd(self):
        self._fwd_body()
        self._join()

    def _phase_update(self):
        for op in self.ops_update:
            op()

    def _phase_train(self):
        self._fwd_body()
        self._run(self.ops_loss)
        self._run(self.ops_bwd)
        self._join()
        self._run_pgrad()

    def _phase_train_a(self):      # forward + losses + the first part of backward (gradients of the arena tail)
        self._split_run = True
        try:
            self._fwd_body()
            self._run(self.ops_loss)
            self._run(self.ops_bwd[:self._bwd_split[0]])
            self._join()
            self._run_pgrad((0,))
        finally:
            self._split_run = False

    def _phase_train_b(self):      # the rest of backward; runs while the tail's all-reduce is in flight
        self._split_run = True
        try:
            self._fork()
            self._run(self.ops_bwd[self._bwd_split[0]:])
            self._join()
            self._run_pgrad((1,))
        finally:
            self._split_run = False

    def _replay(self, key, fn, use_graph=True, warm=True):
        """Run `fn`'s launches through a captured CUDA graph.  On the first call with warm=True the launches run eagerly
        (sets function attributes, loads modules) and THAT eager run is the execution: the graph is captured afterwards
        and not replayed, so a non-idempotent fn (the optimizer update) runs exactly once per call."""
        if not use_graph:
            fn()
            return
        gr = self._graphs.get(key)
        if gr is None:
            if warm:
                fn()
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                fn()
            self._graphs[key] = gr
            if warm:
                return
        gr.replay()

    def forward(self, use_graph=True):
        """Runs the forward plan on the current input buffers; returns (loc [B,3|bins], ori [B,n|4]) device tensors."""
        self._replay("fwd", self._phase_fwd, use_graph)
        ori = self.ori_q if self.graph.ori_mode == "quaternion" else self.head[self.graph.ori_out]
        return self.head["loc_final"], ori

    def eval_losses(self):
        """Forward + the two weighted head losses on the current input/label buffers, no backward (the validation
        steps of fit_generator, net.py:1158-1159).  Returns [loc_loss, ori_loss] as Python floats."""
        g, B, S = self.graph, self.B, lib.stream_ptr
        wl = float(self.cfg.LOSS_WEIGHTS.get("loc_loss", 1.0))
        wo = float(self.cfg.LOSS_WEIGHTS.get("ori_loss", 1.0))
        self._replay("fwd", self._phase_fwd, True)
        loc = self.head["loc_final"]
        if g.loc_mode == "regression":
            lib.call("urso_rel_loss", loc.data_ptr(), self.gt_loc.data_ptr(), None, self.losses[0:1].data_ptr(), B, 3, wl, S())
        else:
            lib.call("urso_softmax_xent", loc.data_ptr(), self.gt_loc.data_ptr(), None, self.losses[0:1].data_ptr(), B,
                     loc.shape[1], wl, S())
        if g.ori_mode == "quaternion":
            lib.call("urso_quat_head", self.head["ori_q"].data_ptr(), self.gt_ori.data_ptr(), self.ori_q.data_ptr(), None,
                     self.losses[1:2].data_ptr(), B, wo, S())
        else:
            z = self.head["ori_final"]
            lib.call("urso_softmax_xent", z.data_ptr(), self.gt_ori.data_ptr(), None, self.losses[1:2].data_ptr(), B,
                     z.shape[1], wo, S())
        return self.losses.tolist()

    def set_hyper(self, lr, momentum=None, clipnorm=None):
        cfg = self.cfg
        momentum = cfg.LEARNING_MOMENTUM if momentum is None else momentum
        clipnorm = cfg.GRADIENT_CLIP_NORM if clipnorm is None else clipnorm
        if cfg.OPTIMIZER == "SGD":
            h = [lr, momentum, 0, 0, clipnorm, 0, 0, 0]
        else:   # Keras Adam(amsgrad): lr_t = lr * sqrt(1-b2^t) / (1-b1^t), t counted from 1
            t = self.opt_t + 1
            eps = 1e-4 if getattr(cfg, "F16", False) else 1e-7
            h = [lr * math.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t), 0.9, 0.999, eps, clipnorm, 0, 0, 0]
        if h != getattr(self, "_hyper_host", None):      # unchanged hyper-parameters: no copy (SGD with a constant lr)
            self.hyper.copy_(torch.tensor(h, dtype=torch.float32), non_blocking=True)
            self._hyper_host = h

    # ------------------------------------------------------------------ host-fed steps: pipelined upload
    _copy_stream = None

    def upload_async(self, h_img, h_loc, h_ori):
        """Start the host -> device copy of the NEXT batch (pinned host tensors: uint8 [B,H,W,3], fp32 labels) on a
        dedicated copy stream into device staging buffers.  It overlaps with the step the compute stream is running;
        `swap_in()` makes the batch current.  (The reference feeds every step through feed_dict, net.py:1152.)"""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._st = [torch.empty_like(self.img_u8), torch.empty_like(self.gt_loc), torch.empty_like(self.gt_ori)]
            self._ev_uploaded, self._ev_free = torch.cuda.Event(), torch.cuda.Event()
            self._ev_free.record(torch.cuda.current_stream())
        cs = self._copy_stream
        cs.wait_event(self._ev_free)            # the previous swap_in has drained the staging buffers
        with torch.cuda.stream(cs):
            for dst, src in zip(self._st, (h_img, h_loc, h_ori)):
                dst.copy_(src, non_blocking=True)
            self._ev_uploaded.record(cs)

    def wait_upload(self):
        """Host-side wait until the last upload_async has finished reading its (pinned) source buffers."""
        if self._copy_stream is not None:
            self._ev_uploaded.synchronize()

    def swap_in(self, aug_params=None):
        """Device -> device copy of the uploaded batch into the buffers the captured graphs read (compute stream).
        With `aug_params` (ursonet_b200.augment.draw_params records, one per image) the image copy IS the sim2real
        augmentation kernel (net.py:390-406 on the device: staging buffer -> network input, out of place)."""
        main = torch.cuda.current_stream()
        main.wait_event(self._ev_uploaded)
        if aug_params is not None:
            from . import augment
            self._aug_keep = augment.sim2real_device(self._st[0], self.img_u8, aug_params)
        else:
            self.img_u8.copy_(self._st[0], non_blocking=True)
        self.gt_loc.copy_(self._st[1], non_blocking=True)
        self.gt_ori.copy_(self._st[2], non_blocking=True)
        self._ev_free.record(main)

    grad_acc = None

    def accumulate(self, micro, n_micro, use_graph=True):
        """Forward + losses + backward of micro-batch `micro` (0-based) of `n_micro` on the current input / label buffers.
        Gradients of the micro-batches are averaged: after the last one `self.grads` holds the mean (what one step on
        the concatenated batch gives for the per-sample-mean losses; rel_loss is normalised per micro-batch, like per
        shard under data parallelism).  Global batch 256 on one GPU = 8 micro-batches of 32 (BASELINE configs[4])."""
        assert self.training and 0 <= micro < n_micro
        self._replay("train", self._phase_train, use_graph)
        if n_micro > 1:
            if self.grad_acc is None:
                self.grad_acc = torch.zeros_like(self.grads)
            last = micro == n_micro - 1
            lib.call("urso_grad_accumulate", self.grad_acc.data_ptr(), self.grads.data_ptr(),
                     self.grads.data_ptr() if last else None, 0.0 if micro == 0 else 1.0, 1.0 / n_micro,
                     self.grads.numel(), lib.stream_ptr())

    def apply_update(self, lr, allreduce=None, use_graph=True, allreduce_async=None):
        """All-reduce of the flat gradient arena (if any) + regulariser + global-norm clip + optimizer step."""
        self.set_hyper(lr)
        if allreduce_async is not None:
            allreduce_async(self.grads).wait()
        elif allreduce is not None:
            allreduce(self.grads)
        self._replay("update", self._phase_update, use_graph)
        self.opt_t += 1

    def train_step(self, lr, allreduce=None, use_graph=True, allreduce_async=None):
        """One optimisation step on the current input/label buffers: fwd + loss + bwd (graph A), all-reduce of the flat
        gradient arena, regulariser + clip + update (graph B).  `allreduce(t)` reduces in place on the current stream;
        `allreduce_async(t)` returns a handle with .wait() (torch.distributed async_op=True) and enables the overlapped
        schedule: backward is replayed in two graphs and the all-reduce of the arena tail (>= 90 % of the parameters,
        whose gradients are complete first) runs on NCCL's stream while the rest of backward executes."""
        assert self.training
        if allreduce_async is not None and self._bwd_split is not None:
            self.set_hyper(lr)
            off = self._bwd_split[1]
            if use_graph and "train_a" not in self._graphs:
                self._phase_train()    # eager warm-up of EVERY kernel; part B alone is not idempotent (it accumulates
                #                        into the arena part A zeroed), so the two graphs are captured without one
            self._replay("train_a", self._phase_train_a, use_graph, warm=False)
            w1 = allreduce_async(self.grads[off:])
            self._replay("train_b", self._phase_train_b, use_graph, warm=False)
            w2 = allreduce_async(self.grads[:off])
            w1.wait()
            w2.wait()
            self._replay("update", self._phase_update, use_graph)
            self.opt_t += 1
        else:
            self.accumulate(0, 1, use_graph)
            self.apply_update(lr, allreduce, use_graph, allreduce_async)

    def count_launches(self, train=True):
        """Kernels of liburso_b200.so launched per train step (or per forward): bench.py's gpu_launches."""
        def n(ops):
            return sum(getattr(o, "launches", 1) for o in ops)
        total = n(self.ops_stage) + 1 + n(self.ops_fwd)
        if train and self.training:
            total += n(self.ops_loss) + n(self.ops_bwd) + n(self.ops_update) + n(getattr(self, "ops_pgrad", []))
        return total

    def profile_ops(self, train=True, reps=3):
        """Eager per-launch timing with CUDA events on the launching stream (cold-ish caches between different ops,
        which is what a real step sees).  Returns [dict(kind, name, ms, flops, bytes)] for OpRec-tagged launches."""
        ops = list(self.ops_fwd)
        if train and self.training:
            ops += list(self.ops_loss) + list(self.ops_bwd) + list(self.ops_pgrad)
        self._phase_train() if (train and self.training) else self._phase_fwd()    # valid buffers everywhere
        torch.cuda.synchronize()
        recs = {}
        self._serial = True
        for _ in range(reps):
            self.zero_arena.zero_()
            self._run(self.ops_stage)
            self._stage_input()
            for i, op in enumerate(ops):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                op()
                e1.record()
                recs.setdefault(i, []).append((e0, e1))
        torch.cuda.synchronize()
        self._serial = False
        out = []
        for i, op in enumerate(ops):
            if not isinstance(op, OpRec):
                continue
            ms = min(a.elapsed_time(b) for a, b in recs[i])
            out.append(dict(kind=op.kind, name=op.name, ms=ms, flops=op.flops, bytes=op.bytes, launches=op.launches))
        return out
