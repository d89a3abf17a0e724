"""Generates tests/golden/labels_golden.npz by RUNNING THE REFERENCE's own functions (se3lib.py, utils.py) from
/root/reference in this container.  The reference cannot travel to the GPU box, so its outputs are committed as a
small fixture; re-run this script to regenerate:   python tests/golden/make_golden.py

utils.py imports tensorflow / skimage / matplotlib at module level but never uses them in the functions captured here
(utils.py:11 is a dead import), so they are stubbed in sys.modules.
"""
import contextlib
import io
import os
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "labels_golden.npz")


def import_reference():
    for name in ["tensorflow", "skimage", "skimage.color", "skimage.io", "skimage.transform", "matplotlib",
                 "matplotlib.pyplot"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, REF)
    import se3lib  # noqa
    import utils   # noqa
    return se3lib, utils


def main():
    se3lib, utils = import_reference()
    rng = np.random.RandomState(20260101)
    out = {}
    # --- se3lib.euler2quat on a set of angles incl. the poles and wrap-around values
    ang = np.concatenate([rng.uniform([-180, -90, -180], [180, 90, 180], size=(40, 3)),
                          np.array([[0, 0, 0], [180, 90, 180], [-180, -90, -180], [37.5, 90, -12], [10, -90, 170]])])
    out["euler_in"] = ang
    out["euler2quat"] = np.stack([np.asarray(se3lib.euler2quat(*a)).reshape(4) for a in ang])
    # --- unit quaternions with q4 >= 0 (urso.py:57-61)
    q = rng.randn(12, 4)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    q *= np.where(q[:, 3:4] < 0, -1.0, 1.0)
    out["quats"] = q
    lo, hi = np.array([-180, -90, -180]), np.array([180, 90, 180])
    for n, beta in [(8, 6.0), (16, 6.0), (12, 3.0)]:
        with contextlib.redirect_stdout(io.StringIO()):
            enc, H_quat, red = utils.encode_ori(q, n, beta, lo, hi)
        out[f"enc_{n}_{beta}"] = enc
        out[f"Hquat_{n}"] = H_quat
        out[f"red_{n}"] = red
        fast = np.stack([utils.encode_ori_fast(q[i], beta, H_quat, red) for i in range(3)])
        out[f"encfast_{n}_{beta}"] = fast
        # decode path: stable_softmax of pseudo-logits, then quat_weighted_avg (pose_estimator.py:406-409)
        logits = np.maximum(0.0, np.log(enc[:4].astype(np.float64) + 1e-9) + 12.0)     # ReLU'd logits peaked like enc
        pm = np.stack([utils.stable_softmax(l) for l in logits])
        out[f"logits_{n}_{beta}"] = logits
        out[f"pmf_{n}_{beta}"] = pm
        out[f"qavg_{n}_{beta}"] = np.stack([np.asarray(se3lib.quat_weighted_avg(H_quat, p)[0]).reshape(4) for p in pm])
        out[f"qavg_enc_{n}_{beta}"] = np.stack(
            [np.asarray(se3lib.quat_weighted_avg(H_quat, e)[0]).reshape(4) for e in enc[:4]])
    # --- angular error formula (pose_estimator.py:434)
    out["angle_between"] = np.array([float(np.asarray(se3lib.angle_between_quats(q[i], q[i + 1])).reshape(-1)[0])
                                     for i in range(6)])
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
