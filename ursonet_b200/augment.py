"""Host side of the device sim2real augmentation (reference: net.py:390-406, run per image inside load_image_gt).

`draw_params` draws, per image, everything imgaug draws per call of the reference's pipeline -- whether to augment at all
(p = 0.5), the random order of the five augmenters, and each augmenter's parameters -- into records of the C struct
`urso_aug_params` (include/urso_b200.h); the per-pixel work (luma, noise, blur, add, multiply, coarse dropout) runs in ONE
kernel on the uploaded uint8 batch (csrc/augment.cu).  Deviation from the reference, by construction of a device-side
pipeline: the reference augments the full-size image BEFORE resize + pad64; here the network-resolution frame is augmented
(inside the image window only, so the padding stays zero as in the reference)."""

import numpy as np

from . import lib

AUG_DTYPE = np.dtype([("apply", "<i4"), ("order", "<i4", (5,)), ("noise_q", "<i4"), ("noise_seed", "<u4"),
                      ("blur_sigma", "<f4"), ("blur_w", "<f4", (5,)), ("add", "<i4"), ("mul", "<f4"),
                      ("drop_thresh", "<u4"), ("drop_h", "<i4"), ("drop_w", "<i4"), ("drop_seed", "<u4"),
                      ("win", "<i4", (4,))], align=True)


def gaussian_kernel5(sigma):
    """cv2.getGaussianKernel(5, sigma) for sigma > 0: exp(-(i-2)^2 / (2 sigma^2)), normalised (float32)."""
    x = np.arange(5, dtype=np.float64) - 2
    w = np.exp(-(x * x) / (2.0 * float(sigma) ** 2))
    return (w / w.sum()).astype(np.float32)


def draw_params(rng, windows, p_apply=0.5):
    """windows: [B,4] int (y1, x1, y2, x2) of each image inside its pad64 frame.  rng: numpy RandomState."""
    windows = np.asarray(windows, dtype=np.int32).reshape(-1, 4)
    B = windows.shape[0]
    prm = np.zeros(B, dtype=AUG_DTYPE)
    for b in range(B):
        r = prm[b]
        h, w = int(windows[b, 2] - windows[b, 0]), int(windows[b, 3] - windows[b, 1])
        r["win"] = windows[b]
        r["apply"] = int(rng.rand() > 1.0 - p_apply)                  # net.py:395: np.random.rand(1) > 0.5
        r["order"] = rng.permutation(5)                               # random_order=True
        r["noise_q"] = int(round(65536.0 * (0.01 * 255) / 147.8))     # AdditiveGaussianNoise(scale=0.01*255)
        r["noise_seed"] = rng.randint(0, 2 ** 31 - 1)
        sigma = rng.uniform(0.0, 1.5)                                 # GaussianBlur(sigma=(0.0, 1.5))
        r["blur_sigma"] = sigma
        r["blur_w"] = gaussian_kernel5(max(sigma, 1e-3))
        r["add"] = rng.randint(-20, 21)                               # Add((-20, 20))
        r["mul"] = rng.uniform(0.5, 2.0)                              # Multiply((0.5, 2.0))
        p = (0.0, 0.03)[rng.randint(0, 2)]                            # CoarseDropout([0.0, 0.03], ...)
        s = rng.uniform(0.02, 0.1)                                    # size_percent=(0.02, 0.1)
        r["drop_thresh"] = min(int(p * 2.0 ** 32), 2 ** 32 - 1)
        r["drop_h"], r["drop_w"] = max(1, int(h * s)), max(1, int(w * s))
        r["drop_seed"] = rng.randint(0, 2 ** 31 - 1)
    return prm


def check_layout():
    if lib.load().urso_sizeof_aug_params() != AUG_DTYPE.itemsize:
        raise lib.UrsoError("AUG_DTYPE does not match struct urso_aug_params (rebuild the library)")


def params_to_device(params, device):
    """AUG_DTYPE records -> uint8 CUDA tensor (pinned staging, asynchronous copy on the current stream)."""
    import torch
    host = torch.from_numpy(np.frombuffer(params.tobytes(), dtype=np.uint8).copy()).pin_memory()
    return host.to(device, non_blocking=True)


def sim2real_device(src_u8, dst_u8, params):
    """src_u8, dst_u8: uint8 CUDA tensors [B,H,W,3] (distinct); params: AUG_DTYPE records [B] (host) or their device copy
    from params_to_device.  Asynchronous on the current stream; returns the device parameter table (keep it alive until
    the kernel has run)."""
    import torch
    check_layout()
    B, H, W, _ = src_u8.shape
    assert src_u8.dtype == torch.uint8 and dst_u8.shape == src_u8.shape
    p_dev = params if isinstance(params, torch.Tensor) else params_to_device(params, src_u8.device)
    assert p_dev.numel() == B * AUG_DTYPE.itemsize
    lib.call("urso_sim2real_aug", src_u8.data_ptr(), dst_u8.data_ptr(), p_dev.data_ptr(), B, H, W, lib.stream_ptr())
    return p_dev
