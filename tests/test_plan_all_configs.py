"""CPU-side sweep of the host planners (csrc/conv_ops.cu + the Engine-F / Engine-W shared-memory and pipeline planning)
over EVERY convolution launch of every backbone at the BASELINE.json shapes, in the library's dry-run mode
(urso_set_dry_run: plans for a 148-SM B200, encodes no tensor maps, touches no device).  Catches "does not fit in shared
memory" / segment-limit failures without a GPU, and pins which launches get the two-pipeline / halo / resident-weight
modes."""
import ctypes as C

import pytest

from ursonet_b200 import lib
from ursonet_b200.config import Config
from ursonet_b200.graph import backward_groups, build_graph

FAKE = 0x7f0000000000      # fake, aligned "device" addresses: the dry run never dereferences them


def make_cfg(backbone, h, w, classify=True, ori_bins=16):
    cfg = Config()
    cfg.BACKBONE, cfg.BOTTLENECK_WIDTH, cfg.BRANCH_SIZE, cfg.NR_DENSE_LAYERS = backbone, 32, 1024, 1
    cfg.ORI_BINS_PER_DIM, cfg.REGRESS_ORI, cfg.REGRESS_LOC = ori_bins, not classify, True
    cfg.IMAGE_RESIZE_MODE, cfg.IMAGE_MIN_DIM, cfg.IMAGE_MAX_DIM = "pad64", h, w
    cfg.update()
    return cfg


@pytest.fixture
def dry():
    l = lib.load()
    l.urso_set_dry_run(1)
    yield l
    l.urso_set_dry_run(0)


def shape_of(g, c, B, H, W):
    if c.stem:
        return lib.conv_shape(B, H, W, 3, c.cout, 7, 2, 3)
    h, w, _ = g.shapes[c.src]
    return lib.conv_shape(B, h, w, c.cin, c.cout, c.k, c.stride, c.padding)


def plan_network(l, cfg, B):
    g = build_graph(cfg)
    H, W = int(cfg.IMAGE_SHAPE[0]), int(cfg.IMAGE_SHAPE[1])
    infos = {}
    for c in g.convs:
        d = lib.Conv2dFwdDesc()
        d.shape = shape_of(g, c, B, H, W)
        d.x, d.w, d.scale, d.shift, d.y, d.workspace = FAKE, FAKE, FAKE, FAKE, FAKE, FAKE
        d.addend = FAKE if c.addend else None
        d.relu, d.out_fp32 = int(c.relu), int(c.out_fp32)
        d.relu_bits = FAKE if (c.relu and c.dst in g.relu_buffers and c.dst != g.pool_src) else None    # as the engine does
        h = C.c_void_p()
        rc = l.urso_conv2d_fwd_create(C.byref(d), C.byref(h))
        assert rc == 0, (c.name, l.urso_last_error())
        v = (C.c_int32 * 9)()
        l.urso_conv2d_fwd_plan_info(h, v)
        infos[c.name] = dict(zip(("block_n", "npipe", "stages", "kpack", "halo", "bres", "a_stages", "smem", "grid"), v))
        assert infos[c.name]["smem"] <= 227 * 1024
        infos[c.name]["tail_split"] = l.urso_conv2d_fwd_tail_split(h)
        l.urso_conv2d_fwd_destroy(h)
        # weight gradient
        wd = lib.Conv2dWgradDesc()
        wd.shape, wd.x, wd.dy, wd.G = d.shape, FAKE, FAKE, FAKE
        hw = C.c_void_p()
        assert l.urso_conv2d_wgrad_create(C.byref(wd), C.byref(hw)) == 0, (c.name, l.urso_last_error())
        w10 = (C.c_int32 * 10)()
        l.urso_conv2d_wgrad_plan_info(hw, w10)
        infos["w:" + c.name] = dict(zip(("block_q", "halo", "halo_w", "stages", "stage_bytes", "units", "groups", "split_k",
                                         "grid", "pair_mode"), w10))
        assert w10[3] >= 2 and w10[3] * w10[4] <= 200 * 1024, (c.name, list(w10))
        l.urso_conv2d_wgrad_destroy(hw)
    groups, sparse = backward_groups(g)
    n_dgrad = 0
    for grp in groups:
        d = lib.Conv2dDgradDesc()
        d.n_convs = len(grp["convs"])
        for i, c in enumerate(grp["convs"]):
            d.shape[i], d.dy[i], d.w[i], d.scale[i] = shape_of(g, c, B, H, W), FAKE, FAKE, FAKE
        d.dy_sparse = int(grp["sparse_in"])
        bits = grp["mask"] and grp["X"] != "pool1"          # the engine: bit-packed masks except for pool1
        d.mask = FAKE if (grp["mask"] and not bits) else None
        d.mask_bits = FAKE if bits else None
        d.addend = FAKE if grp["add"] else None
        d.dx, d.colsum, d.workspace = FAKE, (FAKE if grp["colsum"] else None), FAKE
        assert l.urso_conv2d_dgrad_workspace_bytes(C.byref(d)) > 0, (grp["X"], l.urso_last_error())
        h = C.c_void_p()
        rc = l.urso_conv2d_dgrad_create(C.byref(d), C.byref(h))
        assert rc == 0, (grp["X"], l.urso_last_error())
        assert (l.urso_conv2d_dgrad_untouched_phases(h) == 0b1110) == grp["only_phase0"], grp["X"]
        n_dgrad += l.urso_conv2d_dgrad_num_launches(h)
        v = (C.c_int32 * 9)()
        l.urso_conv2d_dgrad_plan_info(h, 0, v)
        infos["d:" + grp["X"]] = dict(zip(("block_n", "npipe", "stages", "kpack", "halo", "bres", "a_stages", "smem", "grid"), v))
        infos["d:" + grp["X"]]["tail_split"] = l.urso_conv2d_dgrad_tail_split(h, 0)
        l.urso_conv2d_dgrad_destroy(h)
    return g, infos, n_dgrad, sparse


CONFIGS = [   # BASELINE.json configs[0..4] + the other backbones at the bench shape
    ("resnet18", 256, 320, 1, True, 16),
    ("resnet50", 640, 960, 32, True, 16),
    ("resnet50", 1216, 1920, 16, False, 16),
    ("resnet101", 640, 960, 8, True, 24),
    ("resnet34", 640, 960, 32, True, 16),
    ("resnet18", 640, 960, 32, False, 16),
    ("resnet50", 128, 192, 2, True, 8),        # the toy shape of the GPU model tests
]


@pytest.mark.parametrize("backbone,h,w,B,classify,bins", CONFIGS)
def test_every_launch_of_the_network_plans(dry, backbone, h, w, B, classify, bins):
    cfg = make_cfg(backbone, h, w, classify, bins)
    g, infos, n_dgrad, sparse = plan_network(dry, cfg, B)
    assert sum(1 for k in infos if k[:2] not in ("d:", "w:")) == len(g.convs) and n_dgrad >= len(g.convs) // 2


def test_bench_workload_gets_the_intended_modes(dry):
    """RN-50, 640x960, B = 32: the stem and the stage-2 3x3 convs run in halo mode with the weight operand resident and
    two pipelines; the stage-3 3x3 convs in halo mode with streamed weights; every BLOCK_N <= 128 launch with enough tiles
    gets two pipelines; BLOCK_N = 256 launches keep one."""
    cfg = make_cfg("resnet50", 640, 960)
    _, infos, _, sparse = plan_network(dry, cfg, 32)
    for name in ("conv1", "res2a_branch2b", "res2b_branch2b", "res2c_branch2b"):
        assert infos[name]["halo"] == 1 and infos[name]["bres"] == 1 and infos[name]["npipe"] == 2, (name, infos[name])
    # the 3x3 input gradients plan like their forward twins now that the ReLU mask is bit-packed (no smem ring for it)
    # (the last block of a stage has a sparse output gradient: from 128 channels on its 3x3 dgrad runs as 4 decimated phase
    # launches instead; the 64-channel one stays dense, which is faster)
    for name in ("d:res2a_branch2a", "d:res2b_branch2a", "d:res2c_branch2a"):
        assert infos[name]["halo"] == 1 and infos[name]["bres"] == 1 and infos[name]["npipe"] == 2, (name, infos[name])
    for name in ("d:res3a_branch2a", "d:res3c_branch2a"):
        assert infos[name]["halo"] == 1 and infos[name]["npipe"] == 2, (name, infos[name])
    for name in ("res3a_branch2b", "res3d_branch2b"):
        assert infos[name]["halo"] == 1 and infos[name]["bres"] == 0 and infos[name]["npipe"] == 2, (name, infos[name])
    # Engine W: the stem and the 64 / 128-channel 3x3 weight gradients read all taps of a CTA from one box + halo per K step
    for name in ("w:conv1", "w:res2a_branch2b", "w:res2b_branch2b", "w:res3a_branch2b", "w:res3d_branch2b"):
        assert infos[name]["halo"] == 1 and infos[name]["stages"] >= 5, (name, infos[name])
    assert infos["w:res2a_branch2b"]["units"] == 5 and infos["w:res2a_branch2b"]["groups"] == 1      # all 9 taps in one CTA
    assert infos["w:res4b_branch2b"]["halo"] == 0 and infos["w:res5b_branch2b"]["halo"] == 0
    # N-split tail: the K-heavy BLOCK_N = 256 launches of stages 4 and 5 cut the tiles of their partial last wave along N
    # (600 tiles on 148 CTAs: 8 tail tiles x 4 sub-tiles; 640 tiles: 48 tail tiles x 2)
    assert infos["res4b_branch2a"]["tail_split"] == 4 and infos["res4b_branch2b"]["tail_split"] == 2
    assert infos["res5b_branch2b"]["tail_split"] == 4 and infos["d:res4b_branch2b"]["tail_split"] == 4
    assert infos["res2a_branch2c"]["tail_split"] == 1 and infos["res3a_branch2b"]["tail_split"] == 1
    for name, i in infos.items():
        if name.startswith("w:"):
            continue
        if i["block_n"] == 256:
            assert i["npipe"] == 1, (name, i)
        assert i["npipe"] == 1 or i["block_n"] <= 128
        assert i["a_stages"] >= 2 if i["halo"] else i["stages"] >= 2, (name, i)
    # buffers consumed only by the 1x1/stride-2 convs of the next stage, and what feeds them through 1x1 convs
    assert {"res2c_out", "res3d_out", "res4f_out", "res2c_branch2b"} <= sparse


def test_dry_run_refuses_to_launch(dry):
    d = lib.Conv2dFwdDesc()
    d.shape = lib.conv_shape(1, 16, 16, 64, 64, 1, 1, "valid")
    d.x, d.w, d.y, d.workspace = FAKE, FAKE, FAKE, FAKE
    h = C.c_void_p()
    assert dry.urso_conv2d_fwd_create(C.byref(d), C.byref(h)) == 0
    assert dry.urso_conv2d_fwd_launch(h, None) != 0 and b"dry run" in dry.urso_last_error()
    dry.urso_conv2d_fwd_destroy(h)
