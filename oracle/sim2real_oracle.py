"""TEST INFRASTRUCTURE (oracle): numpy restatement of the reference's sim2real augmentation (net.py:390-406), in the exact
arithmetic of the device kernel ursonet_b200/csrc/augment.cu so that the two can be compared bit for bit.

Pinned to the reference: the luma step (`luma_reference` is the reference's own expression, net.py:391-394).
NOT pinned ("parity unpinned"): the five imgaug augmenters -- imgaug is not installed here and the reference holds no
fixtures for them.  They are restated from imgaug's documented behaviour:
  AdditiveGaussianNoise(scale=2.55, per_channel=False): one N(0, 2.55) sample per pixel, added, clipped to uint8
     (here: Irwin-Hall(4) integer approximation of the normal, so that CPU and GPU agree exactly);
  GaussianBlur(sigma): imgaug's cv2 path -- kernel size 5 for sigma < 3 (`_compute_gaussian_blur_ksize`: max(3.3*sigma, 5)),
     BORDER_REFLECT_101, skipped below sigma = 1e-3; float32 separable filter, rounded half to even;
  Add(v), Multiply(m): per-image scalar, result rounded and clipped to [0, 255];
  CoarseDropout(p, size_percent): Bernoulli(p) on a low-resolution grid of int(H*s) x int(W*s) cells (>= 1), nearest-
     neighbour upsampling, dropped pixels set to 0;
  iaa.Sequential(random_order=True): the five run in a per-image random order; every augmenter returns uint8.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import numpy as np


def luma_reference(image):
    """The reference's own lines (net.py:391-394) on a uint8 HxWx3 array (in place on a copy)."""
    image = image.copy()
    image_gray = 0.2126 * image[:, :, 0] + 0.7152 * image[:, :, 1] + 0.0722 * image[:, :, 2]
    image[:, :, 0] = image_gray
    image[:, :, 1] = image_gray
    image[:, :, 2] = image_gray
    return image


def hash_u32(seed, idx):
    x = (idx.astype(np.uint64) * np.uint64(0x9E3779B1) + np.uint64(seed)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16); x = (x * np.uint64(0x7FEB352D)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(15); x = (x * np.uint64(0x846CA68B)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16)
    return x.astype(np.uint32)


def _round_clip(v):
    return np.clip(np.rint(v), 0, 255).astype(np.float32)


def _blur(v, w):
    """5-tap separable float32 filter, un-fused multiply-adds in tap order, BORDER_REFLECT_101."""
    w = np.asarray(w, np.float32)

    def pass1d(a, axis):
        pad = [(0, 0), (0, 0)]
        pad[axis] = (2, 2)
        ap = np.pad(a, pad, mode="reflect")          # numpy 'reflect' == cv2 BORDER_REFLECT_101
        n = a.shape[axis]
        sl = lambda k: np.take(ap, np.arange(k, k + n), axis=axis)
        acc = (sl(0) * w[0]).astype(np.float32)
        for k in range(1, 5):
            acc = (acc + (sl(k) * w[k]).astype(np.float32)).astype(np.float32)
        return acc
    return pass1d(pass1d(v.astype(np.float32), 1), 0)


def augment_image(image, prm, W_full, y_off=0, x_off=0):
    """image: uint8 HxWx3 (the un-padded window); prm: one record of ursonet_b200.augment.AUG_DTYPE; W_full / offsets:
    geometry of the padded frame the kernel indexes its per-pixel hash with.  Returns uint8 HxWx3."""
    g = luma_reference(image)[:, :, 0].astype(np.float32)
    h, w = g.shape
    if prm["apply"]:
        yy, xx = np.meshgrid(np.arange(h) + y_off, np.arange(w) + x_off, indexing="ij")
        for op in prm["order"]:
            if op == 0:
                hsh = hash_u32(int(prm["noise_seed"]), (yy * W_full + xx).astype(np.uint32))
                z = ((hsh & 255).astype(np.int64) + ((hsh >> 8) & 255) + ((hsh >> 16) & 255) + (hsh >> 24)) - 510
                t = z * int(prm["noise_q"]) + np.where(z >= 0, 32768, -32768)
                n = np.sign(t) * (np.abs(t) // 65536)
                g = np.clip(g + n.astype(np.float32), 0, 255).astype(np.float32)
            elif op == 1:
                if prm["blur_sigma"] >= 1e-3 and (h > 1 and w > 1):
                    g = _round_clip(_blur(g, prm["blur_w"]))
            elif op == 2:
                g = np.clip(g + np.float32(prm["add"]), 0, 255).astype(np.float32)
            elif op == 3:
                g = _round_clip((g * np.float32(prm["mul"])).astype(np.float32))
            elif op == 4:
                cy = ((yy - y_off).astype(np.int64) * int(prm["drop_h"])) // h
                cx = ((xx - x_off).astype(np.int64) * int(prm["drop_w"])) // w
                hsh = hash_u32(int(prm["drop_seed"]), (cy * int(prm["drop_w"]) + cx).astype(np.uint32))
                g = np.where(hsh < np.uint32(prm["drop_thresh"]), np.float32(0), g)
    out = g.astype(np.uint8)
    return np.stack([out] * 3, -1)


def augment_batch(images, params):
    """images uint8 [B,H,W,3] (pad64 frames), params [B] records with the window of each image: the padding stays as is
    (but goes through the luma step like every pixel the kernel touches)."""
    B, H, W, _ = images.shape
    out = np.empty_like(images)
    for b in range(B):
        y1, x1, y2, x2 = (int(v) for v in params[b]["win"])
        frame = luma_reference(images[b])
        frame[y1:y2, x1:x2] = augment_image(images[b, y1:y2, x1:x2], params[b], W, y1, x1)
        out[b] = frame
    return out
