"""Host-side data path (CPU): URSO-format reader, resize/pad rules, batch generator contract, CLI flag -> Config map."""
import numpy as np
import pytest

from ursonet_b200 import data as D
from ursonet_b200 import pose_estimator as PE


def test_resize_pad64_rules():
    img = np.full((1200, 1920, 3), 7, np.uint8)
    out, window, scale, padding, crop = D.resize_image(img, min_dim=640, max_dim=960, min_scale=0, mode="pad64")
    assert out.shape == (640, 960, 3) and out.dtype == np.uint8
    assert scale == 0.5 and window == (20, 0, 620, 960) and padding[0] == (20, 20)   # utils.py:480-500: centred rows
    assert (out[:20] == 0).all() and (out[620:] == 0).all() and (out[20:620] == 7).all()
    out, window, *_ = D.resize_image(np.zeros((960, 1280, 3), np.uint8), min_dim=256, max_dim=320, min_scale=0, mode="pad64")
    assert out.shape == (256, 320, 3) and window == (8, 0, 248, 320)
    sq, window, *_ = D.resize_image(np.zeros((960, 1280, 3), np.uint8), min_dim=256, max_dim=320, mode="square")
    assert sq.shape == (320, 320, 3) and window == (40, 0, 280, 320)


def test_cli_config_mapping_matches_reference_rules():
    args = PE.build_parser().parse_args(["train", "--dataset", "speed", "--weights", "none", "--image_scale", "0.5",
                                         "--batch_size", "32", "--ori_resolution", "16"])
    cfg = PE.make_config(args)
    assert tuple(cfg.IMAGE_SHAPE) == (640, 960, 3) and cfg.BATCH_SIZE == 32          # 600 -> 640 (pose_estimator.py:856-860)
    assert cfg.REGRESS_ORI is False and cfg.REGRESS_LOC is True and cfg.OPTIMIZER == "SGD"
    assert cfg.BOTTLENECK_WIDTH == 32 and cfg.BRANCH_SIZE == 1024 and cfg.NR_DENSE_LAYERS == 1
    args = PE.build_parser().parse_args(["train", "--dataset", "speed", "--weights", "none", "--image_scale", "0.25"])
    with pytest.raises(Exception, match="Scale problem"):                              # 480 % 64 != 0
        PE.make_config(args)
    args = PE.build_parser().parse_args(["evaluate", "--dataset", "soyuz_easy", "--weights", "last", "--image_scale", "0.25",
                                         "--backbone", "resnet18", "--regress_ori"])
    cfg = PE.make_config(args)
    assert tuple(cfg.IMAGE_SHAPE) == (256, 320, 3) and cfg.BATCH_SIZE == 1 and cfg.REGRESS_ORI is True


def test_synthetic_urso_reader_and_generator(tmp_path):
    d = tmp_path / "datasets" / "synth"
    D.write_synthetic_urso(str(d), n_train=3, n_val=2, n_test=1, width=320, height=240)
    args = PE.build_parser().parse_args(["train", "--dataset", "synth", "--weights", "none", "--image_scale", "0.25",
                                         "--backbone", "resnet18", "--batch_size", "2", "--ori_resolution", "8"])
    cfg = PE.make_config(args)
    ds = D.Urso()
    ds.load_dataset(str(d), cfg, "train")
    assert len(ds.image_ids) == 3
    assert all(ds.load_quaternion(i)[3] >= 0 for i in ds.image_ids)                     # hemisphere rule (urso.py:57-61)
    assert ds.load_orientation_encoded(0).shape == (512,) and abs(ds.load_orientation_encoded(0).sum() - 1) < 1e-5
    img = ds.load_image(0)
    assert img.shape == (240, 320, 3) and img.dtype == np.uint8
    gen = D.data_generator(ds, cfg, shuffle=False, batch_size=2, raw_uint8=False)
    (images, metas, locs, oris), outs = next(gen)
    assert outs == [] and images.shape == (2, 256, 320, 3) and images.dtype == np.float32
    assert metas.shape == (2, 12) and locs.shape == (2, 3) and oris.shape == (2, 512)
    raw, _, _, _ = next(D.data_generator(ds, cfg, shuffle=False, batch_size=2, raw_uint8=True))[0]
    assert raw.dtype == np.uint8
    # molded == raw - MEAN_PIXEL, including the pad rows which enter the net as -MEAN_PIXEL
    assert np.allclose(images, raw.astype(np.float32) - cfg.MEAN_PIXEL)
    assert np.allclose(images[0, 0, 0], -cfg.MEAN_PIXEL)


def test_rotation_augmentation_flags_run(tmp_path):
    """--rot_aug / --rot_image_aug (the reference's own URSO training recipe, README.md:103) go through the generator."""
    d = tmp_path / "ds"
    D.write_synthetic_urso(str(d), 2, 1, 1, width=320, height=240)
    args = PE.build_parser().parse_args(["train", "--dataset", "x", "--weights", "none", "--image_scale", "0.25", "--rot_aug",
                                         "--rot_image_aug"])
    cfg = PE.make_config(args)
    ds = D.Urso(); ds.load_dataset(str(d), cfg, "train")
    np.random.seed(1)
    (images, metas, locs, oris), _ = next(D.data_generator(ds, cfg, batch_size=2, raw_uint8=True))
    assert images.dtype == np.uint8 and images.shape[0] == 2 and np.isfinite(locs).all()
    assert np.allclose(oris.sum(1), 1.0, atol=1e-4)


def test_parallel_loader_yields_batches_from_worker_threads(tmp_path):
    """The multi-threaded host loader (fit_generator(workers=...) counterpart) produces well-formed batches, covers the
    dataset across its workers and shuts down cleanly."""
    from ursonet_b200 import data as D
    from ursonet_b200.config import Config
    root = tmp_path / "ds"
    D.write_synthetic_urso(str(root), n_train=6, n_val=2, n_test=2, width=320, height=256)
    cfg = Config()
    cfg.BACKBONE, cfg.ORI_BINS_PER_DIM, cfg.REGRESS_ORI, cfg.ROT_AUG = "resnet18", 4, False, False
    cfg.IMAGE_MIN_DIM, cfg.IMAGE_MAX_DIM, cfg.IMAGE_RESIZE_MODE = 256, 320, "pad64"
    cfg.update()
    ds = D.Urso()
    ds.load_dataset(str(root), cfg, "train")
    loader = D.ParallelLoader(ds, cfg, batch_size=2, workers=3, shuffle=True, raw_uint8=True)
    assert loader.workers == 3
    seen = set()
    for _ in range(9):
        (images, metas, locs, oris), _ = next(loader)
        assert images.shape == (2, 256, 320, 3) and images.dtype.name == "uint8"
        assert locs.shape == (2, 3) and oris.shape == (2, 64)
        seen.update(int(m[0]) for m in metas)
    loader.close()
    assert seen == set(range(6))
