"""Generates tests/golden/rotaug_golden.npz by RUNNING THE REFERENCE's own utils.rotate_cam / utils.rotate_image
(utils.py:30-86, with se3lib) from /root/reference in this container; the outputs are committed as a small fixture.
    python tests/golden/make_golden_rotaug.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import import_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rotaug_golden.npz")


def main():
    se3lib, utils = import_reference()
    h, w = 60, 80
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([(3 * xx) % 256, (4 * yy) % 256, ((xx - 40) ** 2 + (yy - 30) ** 2 < 200) * 255], -1).astype(np.uint8)
    fx = w / (2 * np.tan(np.deg2rad(90.0) / 2))
    fy = -h / (2 * np.tan(np.deg2rad(73.7) / 2))           # the URSO camera's sign convention (urso.py)
    K = np.matrix([[fx, 0, w / 2], [0, fy, h / 2], [0, 0, 1]])
    rng = np.random.RandomState(7)
    out = {"image": img, "K": np.asarray(K)}
    for i in range(4):
        t = np.array([rng.uniform(-2, 2), rng.uniform(-2, 2), rng.uniform(5, 40)])
        q = rng.randn(4); q /= np.linalg.norm(q)
        out[f"t{i}"], out[f"q{i}"] = t, q
        np.random.seed(100 + i)
        im, tn, qn = utils.rotate_cam(img.copy(), t, q, K, 20)
        out[f"cam_img{i}"], out[f"cam_t{i}"], out[f"cam_q{i}"] = im, np.asarray(tn, dtype=np.float64), np.asarray(qn, dtype=np.float64)
        np.random.seed(200 + i)
        im, tn, qn = utils.rotate_image(img.copy(), t, q, K)
        out[f"img_img{i}"], out[f"img_t{i}"], out[f"img_q{i}"] = im, np.asarray(tn, dtype=np.float64), np.asarray(qn, dtype=np.float64)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
