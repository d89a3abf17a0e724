"""sim2real augmentation: host parameter drawing, the numpy oracle against the reference's own luma lines,
and -- on the GPU -- the device kernel against the oracle, bit for bit."""
import numpy as np
import pytest

from oracle import sim2real_oracle as S
from ursonet_b200 import augment


def test_luma_matches_reference_lines():
    # net.py:391-394 verbatim (float64 weighted sum stored into the uint8 image = truncation)
    rng = np.random.RandomState(1)
    img = rng.randint(0, 256, (37, 53, 3), dtype=np.uint8)
    ref = img.copy()
    image_gray = 0.2126 * ref[:, :, 0] + 0.7152 * ref[:, :, 1] + 0.0722 * ref[:, :, 2]
    ref[:, :, 0] = image_gray; ref[:, :, 1] = image_gray; ref[:, :, 2] = image_gray
    assert np.array_equal(S.luma_reference(img), ref)
    prm = augment.draw_params(rng, [[0, 0, 37, 53]])
    prm["apply"] = 0
    assert np.array_equal(S.augment_batch(img[None], prm)[0], ref)


def test_param_ranges_and_layout():
    rng = np.random.RandomState(0)
    prm = augment.draw_params(rng, np.tile([[20, 0, 620, 960]], (256, 1)))
    assert augment.AUG_DTYPE.itemsize == 96
    assert 0.35 < prm["apply"].mean() < 0.65                                   # p = 0.5 (net.py:395)
    assert all(sorted(o) == [0, 1, 2, 3, 4] for o in prm["order"])             # random_order=True
    assert prm["add"].min() >= -20 and prm["add"].max() <= 20 and prm["add"].min() < -10 < 10 < prm["add"].max()
    assert prm["mul"].min() >= 0.5 and prm["mul"].max() <= 2.0
    assert prm["blur_sigma"].min() >= 0 and prm["blur_sigma"].max() <= 1.5
    assert np.allclose(prm["blur_w"].sum(1), 1.0, atol=1e-6)
    assert set(np.unique(prm["drop_thresh"])) <= {0, int(0.03 * 2 ** 32)}
    assert prm["drop_h"].min() >= int(600 * 0.02) and prm["drop_h"].max() <= int(600 * 0.1)
    # noise: Irwin-Hall(4) scaled to sigma = 2.55 grey levels
    z = S.hash_u32(123, np.arange(200000, dtype=np.uint32))
    zz = ((z & 255).astype(np.int64) + ((z >> 8) & 255) + ((z >> 16) & 255) + (z >> 24)) - 510
    n = np.sign(zz * 1131) * (np.abs(zz * int(prm["noise_q"][0]) + np.where(zz >= 0, 32768, -32768)) // 65536)
    assert abs(n.mean()) < 0.05 and abs(n.std() - 2.55) < 0.15


def test_oracle_operations_behave():
    rng = np.random.RandomState(3)
    img = rng.randint(0, 256, (2, 64, 128, 3), dtype=np.uint8)
    prm = augment.draw_params(rng, [[4, 0, 60, 128], [0, 8, 64, 120]])
    prm["apply"] = 1
    prm["drop_thresh"] = int(0.03 * 2 ** 32)
    out = S.augment_batch(img, prm)
    assert out.dtype == np.uint8 and out.shape == img.shape
    assert np.array_equal(out[..., 0], out[..., 1]) and np.array_equal(out[..., 1], out[..., 2])
    # outside the window only the luma step happened
    assert np.array_equal(out[0, :4], S.luma_reference(img[0])[:4])
    assert np.array_equal(out[1, :, :8], S.luma_reference(img[1])[:, :8])
    assert not np.array_equal(out[0, 4:60], S.luma_reference(img[0])[4:60])


@pytest.mark.gpu
@pytest.mark.parametrize("H,W,win", [(64, 128, (4, 0, 60, 128)), (128, 192, (0, 0, 128, 192)), (80, 70, (3, 5, 77, 66))])
def test_device_matches_oracle_bit_exact(H, W, win):
    import torch
    rng = np.random.RandomState(H + W)
    B = 6
    img = rng.randint(0, 256, (B, H, W, 3), dtype=np.uint8)
    prm = augment.draw_params(rng, np.tile([list(win)], (B, 1)))
    prm["apply"][:5] = 1
    prm["apply"][5] = 0
    prm["drop_thresh"][:3] = int(0.03 * 2 ** 32)
    for b in range(5):                       # every augmenter position for the blur gets exercised
        prm["order"][b] = np.roll(np.arange(5), b)
    prm["blur_sigma"][4] = 0.0               # imgaug skips the blur below 1e-3
    src = torch.from_numpy(img).cuda()
    dst = torch.empty_like(src)
    keep = augment.sim2real_device(src, dst, prm)
    torch.cuda.synchronize()
    del keep
    ref = S.augment_batch(img, prm)
    got = dst.cpu().numpy()
    assert np.array_equal(got, ref), (np.argwhere(got != ref)[:5], int((got != ref).sum()))
