"""Camera-rotation augmentations (net.py:409-438) against golden vectors produced by the reference's own
utils.rotate_cam / utils.rotate_image (tests/golden/make_golden_rotaug.py)."""
import os

import numpy as np

from ursonet_b200 import data as D

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "rotaug_golden.npz"))


def test_rotate_cam_and_rotate_image_match_reference():
    img, K = G["image"], G["K"]
    for i in range(4):
        t, q = G[f"t{i}"], G[f"q{i}"]
        im, tn, qn = D.rotate_cam(img.copy(), t, q, K, 20, rng=np.random.RandomState(100 + i))
        assert np.array_equal(im, G[f"cam_img{i}"])
        assert np.allclose(tn, G[f"cam_t{i}"], rtol=0, atol=1e-12) and np.allclose(qn, G[f"cam_q{i}"], rtol=0, atol=1e-12)
        im, tn, qn = D.rotate_image(img.copy(), t, q, K, rng=np.random.RandomState(200 + i))
        assert np.array_equal(im, G[f"img_img{i}"])
        assert np.allclose(tn, G[f"img_t{i}"], rtol=0, atol=1e-12) and np.allclose(qn, G[f"img_q{i}"], rtol=0, atol=1e-12)
        assert abs(np.linalg.norm(qn) - 1) < 1e-12


def test_quaternion_of_rotation_all_branches():
    # the four branches of the matrix -> JPL quaternion conversion agree with composing the rotation back
    for ang in [(10, 20, 30), (170, 5, 5), (5, 170, 5), (5, 5, 170), (-120, 80, 45)]:
        R = D._rot_xyz_deg(*ang)
        q = D._rot_to_quat_jpl(R)
        x, y, z, w = q
        # JPL convention: R = (2w^2 - 1) I - 2w [q]x + 2 q q^T
        qx = np.array([[0, -z, y], [z, 0, -x], [-y, x, 0]])
        R2 = (2 * w * w - 1) * np.eye(3) - 2 * w * qx + 2 * np.outer(q[:3], q[:3])
        assert np.allclose(R, R2, atol=1e-12), ang


def test_generator_applies_rotation_and_reencodes(tmp_path):
    from ursonet_b200.config import Config
    D.write_synthetic_urso(str(tmp_path / "synth"), n_train=3, n_val=1, n_test=1, height=96, width=128)
    cfg = Config()
    cfg.NAME, cfg.BACKBONE, cfg.ORI_BINS_PER_DIM, cfg.REGRESS_ORI = "t", "resnet18", 8, False
    cfg.IMAGE_MAX_DIM, cfg.IMAGE_MIN_DIM, cfg.IMAGE_RESIZE_MODE = 128, 128, "pad64"
    cfg.ROT_AUG, cfg.ROT_IMAGE_AUG = True, True
    cfg.update()
    ds = D.Urso(); ds.load_dataset(str(tmp_path / "synth"), cfg, "train")
    np.random.seed(0)
    plain = ds.load_orientation_encoded(0)
    image, meta, loc, ori = D.load_image_gt(ds, cfg, 0)
    assert image.dtype == np.uint8 and image.shape == (128, 128, 3)
    assert ori.shape == plain.shape and abs(ori.sum() - 1) < 1e-4 and not np.allclose(ori, plain)   # re-encoded label
    assert not np.allclose(loc, ds.load_location(0))
    assert abs(np.linalg.norm(loc) - np.linalg.norm(ds.load_location(0))) < 1e-9                  # a pure rotation
