"""The product's layer graph (ursonet_b200/graph.py) and the oracle's independent restatement agree on every Keras
weight name and shape, for all four backbones and both orientation heads (CPU only)."""
import pytest

from oracle import ursonet_oracle as O
from ursonet_b200.config import Config
from ursonet_b200.graph import build_graph, weight_entries


def cfg_for(backbone, regress_ori, h=256, w=320, bw=32, bins=16):
    c = Config()
    c.BACKBONE, c.REGRESS_ORI, c.BOTTLENECK_WIDTH, c.ORI_BINS_PER_DIM = backbone, regress_ori, bw, bins
    c.NR_DENSE_LAYERS = 1
    c.IMAGE_MIN_DIM, c.IMAGE_MAX_DIM = h, w
    c.update()
    return c


@pytest.mark.parametrize("backbone", ["resnet18", "resnet34", "resnet50", "resnet101"])
@pytest.mark.parametrize("regress_ori", [False, True])
def test_weight_names_and_shapes_match_oracle(backbone, regress_ori):
    cfg = cfg_for(backbone, regress_ori)
    mine = {n: tuple(s) for n, s, _t, _r in weight_entries(build_graph(cfg))}
    ref = {n: tuple(s) for n, s in O.weight_shapes(cfg).items()}
    assert mine == ref
    for n, _s, trainable, reg in weight_entries(build_graph(cfg)):
        assert trainable == O.is_trainable(n)
        assert reg == O.is_regularised(n)


def test_param_counts_match_survey():
    # trainable parameter counts (incl. gamma/beta, excl. moving stats)
    def count(cfg):
        return sum(int(__import__("numpy").prod(s)) for n, s, t, _ in weight_entries(build_graph(cfg)) if t)
    assert abs(count(cfg_for("resnet50", False, 640, 960, 32, 16)) - 38.16e6) < 0.02e6
    assert abs(count(cfg_for("resnet18", False, 256, 320, 32, 16)) - 16.84e6) < 0.02e6
    assert abs(count(cfg_for("resnet101", False, 640, 960, 32, 24)) - 67.15e6) < 0.02e6
    assert abs(count(cfg_for("resnet50", True, 1216, 1920, 32, 16)) - 61.49e6) < 0.02e6


def test_image_size_check_and_unbuilt_branches():
    c = cfg_for("resnet50", False, 250, 320)
    with pytest.raises(Exception, match="dividable by 2"):
        build_graph(c)
    c = cfg_for("resnet50", False)
    c.REGRESS_KEYPOINTS = True
    with pytest.raises(NotImplementedError):
        build_graph(c)


def test_graph_geometry_resnet50():
    g = build_graph(cfg_for("resnet50", False, 640, 960))
    assert g.shapes["pool1"] == (160, 240, 64)
    assert g.shapes["res5c_out"] == (20, 30, 2048)
    assert g.shapes["bottleneck_layer"] == (10, 15, 32)
    assert g.nr_features == 4800
    assert len(g.convs) == 54          # 53 backbone convs + bottleneck (54 convs)
    g101 = build_graph(cfg_for("resnet101", False, 640, 960))
    assert "res4w_branch2a" in [c.name for c in g101.convs]     # chr(98+21) == 'w' (net.py:190)
