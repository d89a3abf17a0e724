"""Host-side data path feeding the engine: image formatting, dataset readers and the batch generator.

Mirrors the parts of the reference that sit either side of the hot path:
  utils.resize_image (utils.py:398-511), net.mold_image / compose_image_meta (net.py:1314-1355),
  dataset.Dataset (dataset.py), urso.Urso (urso.py:27-153), speed.Speed (speed.py), net.load_image_gt /
  data_generator (net.py:358-559).
Differences, all deliberate and documented in INTEGRATION.md:
  * skimage is not available: images are read with cv2 and resized with cv2.INTER_LINEAR (the reference uses
    skimage.transform.resize(order=1), whose anti-aliasing default is version dependent);
  * the generator can yield RAW uint8 frames (`raw_uint8=True`): mean subtraction then happens on the GPU inside the
    stem staging kernel, which cuts the per-step host->device copy 4x (59 MB instead of 236 MB at 640x960x32);
  * pd.read_csv(header=-1) (urso.py:42) is invalid on pandas >= 1.0: header=None is used;
  * sim2real augmentation (net.py:390-406) runs on the device (augment.py / csrc/augment.cu) for uint8 feeds; the
    camera-rotation augmentations (net.py:409-438, utils.py:30-86) run on the host with cv2 like the reference's.
"""
import json
import logging
import os

import cv2
import numpy as np

from . import labels


# ------------------------------------------------------------------------------------------------ image formatting
def resize_image(image, min_dim=None, max_dim=None, min_scale=None, mode="square"):
    """Same contract as utils.resize_image: returns (image, window, scale, padding, crop)."""
    image_dtype = image.dtype
    h, w = image.shape[:2]
    window, scale, padding, crop = (0, 0, h, w), 1, [(0, 0), (0, 0), (0, 0)], None
    if mode == "none":
        return image, window, scale, padding, crop
    if min_dim:
        scale = min_dim / min(h, w)
    if min_scale and scale < min_scale:
        scale = min_scale
    if max_dim and mode != "crop":
        if round(max(h, w) * scale) > max_dim:
            scale = max_dim / max(h, w)
    if scale != 1:
        image = cv2.resize(image, (round(w * scale), round(h * scale)), interpolation=cv2.INTER_LINEAR)
    h, w = image.shape[:2]
    if mode == "square":
        top, left = (max_dim - h) // 2, (max_dim - w) // 2
        pads = [(top, max_dim - h - top), (left, max_dim - w - left)]
    elif mode == "pad64":
        assert min_dim % 64 == 0, "Minimum dimension must be a multiple of 64"
        pads = []
        for n in (h, w):
            if n % 64 > 0:
                tot = n - (n % 64) + 64 - n
                pads.append((tot // 2, tot - tot // 2))
            else:
                pads.append((0, 0))
    elif mode == "crop":
        y = np.random.randint(0, h - min_dim + 1)
        x = np.random.randint(0, w - min_dim + 1)
        return image[y:y + min_dim, x:x + min_dim].astype(image_dtype), (0, 0, min_dim, min_dim), scale, padding, \
            (y, x, min_dim, min_dim)
    else:
        raise Exception("Mode {} not supported".format(mode))
    padding = pads + [(0, 0)] if image.ndim > 2 else pads
    image = np.pad(image, padding, mode="constant", constant_values=0)
    window = (pads[0][0], pads[1][0], h + pads[0][0], w + pads[1][0])
    return image.astype(image_dtype), window, scale, padding, crop


def mold_image(images, config):
    """RGB float image minus MEAN_PIXEL (net.py:1337-1346)."""
    return images.astype(np.float32) - config.MEAN_PIXEL


def unmold_image(normalized_images, config):
    return (normalized_images + config.MEAN_PIXEL).astype(np.uint8)


def compose_image_meta(image_id, original_image_shape, image_shape, window, scale):
    """12 floats: id, original shape (3), molded shape (3), window (4), scale (net.py:1314-1334)."""
    return np.array([image_id] + list(original_image_shape) + list(image_shape) + list(window) + [scale])


# ------------------------------------------------------------------------------------------------ datasets
class Dataset(object):
    """dataset.Dataset: list of per-image info dicts with accessor methods."""

    def __init__(self):
        self._image_ids = []
        self.image_info = []

    def add_image(self, source, image_id, path, **kwargs):
        info = {"id": image_id, "source": source, "path": path}
        info.update(kwargs)
        self.image_info.append(info)

    @property
    def image_ids(self):
        return self._image_ids

    def source_image_link(self, image_id):
        return self.image_info[image_id]["path"]

    def load_location(self, image_id):
        return self.image_info[image_id]["location"]

    def load_quaternion(self, image_id):
        return self.image_info[image_id]["quaternion"]

    def load_orientation_encoded(self, image_id):
        return self.image_info[image_id]["ori_map"]

    def load_image(self, image_id):
        """[H,W,3] uint8 RGB; grayscale replicated, alpha dropped (urso.py:138-153)."""
        img = cv2.imread(self.image_info[image_id]["path"], cv2.IMREAD_UNCHANGED)
        if img is None:
            raise IOError("cannot read " + self.image_info[image_id]["path"])
        if img.ndim == 2:
            img = np.stack([img] * 3, -1)
        elif img.shape[-1] == 4:
            img = img[..., [2, 1, 0]]
        else:
            img = img[..., ::-1]
        return np.ascontiguousarray(img)


class UrsoCamera:
    fov_x, fov_y = 90.0 * np.pi / 180, 73.7 * np.pi / 180
    width, height = 1280, 960
    fx = width / (2 * np.tan(fov_x / 2))
    fy = -height / (2 * np.tan(fov_y / 2))
    K = np.array([[fx, 0, width / 2], [0, fy, height / 2], [0, 0, 1]])


class SpeedCamera:
    fx, fy = 0.0176 / 5.86e-6, 0.0176 / 5.86e-6      # speed.py:15-25: f = 17.6 mm, 5.86 um pixels
    width, height = 1920, 1200
    K = np.array([[fx, 0, width / 2], [0, fy, height / 2], [0, 0, 1]])


def _hemisphere(q):
    q = np.asarray(q, dtype=np.float32)
    return -q if q[3] < 0 else q          # injectivity: keep q4 >= 0 (urso.py:57-61)


class Urso(Dataset):
    """URSO layout: <dir>/<subset>_images.csv (one file name per line) + <subset>_poses_gt.csv (x,y,z,q1..q4)."""
    camera = UrsoCamera()

    def load_dataset(self, dataset_dir, config, subset):
        import pandas as pd
        self.name = "Urso"
        if not os.path.exists(dataset_dir):
            print("Image directory '" + dataset_dir + "' not found.")
            return None
        files = list(pd.read_csv(os.path.join(dataset_dir, subset + "_images.csv"), names=["filename"], header=None)["filename"])
        poses = pd.read_csv(os.path.join(dataset_dir, subset + "_poses_gt.csv"))
        q = np.stack([_hemisphere([poses[k][i] for k in ("q1", "q2", "q3", "q4")]) for i in range(len(files))])
        self._finish(dataset_dir, files, q, np.stack([poses["x"], poses["y"], poses["z"]], 1)[:len(files)], config, "URSO")

    def _finish(self, dataset_dir, files, q, t, config, source):
        if not config.REGRESS_LOC:
            raise NotImplementedError("location classification (--classify_loc, experimental) is not built")
        if not config.REGRESS_ORI:
            self.encoder = labels.OrientationEncoder(config.ORI_BINS_PER_DIM, config.BETA)
            enc = self.encoder.encode(q)
            self.ori_histogram_map, self.ori_output_mask = self.encoder.H_quat, self.encoder.redundant
        for i, f in enumerate(files):
            self.add_image(source, image_id=i, path=os.path.join(dataset_dir, f), location=[float(v) for v in t[i]],
                           quaternion=q[i], ori_map=enc[i] if not config.REGRESS_ORI else [])
        self.num_images = len(self.image_info)
        self._image_ids = np.arange(self.num_images)


class Speed(Urso):
    """SPEED layout (speed.py:29-110): <dir>/<subset>.json = [{filename, q_vbs2tango [w,x,y,z], r_Vo2To_vbs_true}],
    images under <dir>/images/<subset>/."""
    camera = SpeedCamera()

    def load_dataset(self, dataset_dir, config, subset):
        assert subset in ["train", "train_no_val", "val", "test", "real", "real_test", "train_total"]   # speed.py:35
        self.name = "Speed"
        with open(os.path.join(dataset_dir, subset + ".json")) as f:
            items = json.load(f)
        folder = "train" if subset in ("train_no_val", "val") else subset      # speed.py:92-95
        files = [os.path.join("images", folder, it["filename"]) for it in items]
        if items and "q_vbs2tango" in items[0]:
            q = np.stack([_hemisphere([it["q_vbs2tango"][1], it["q_vbs2tango"][2], it["q_vbs2tango"][3],
                                       it["q_vbs2tango"][0]]) for it in items])
            t = np.asarray([it["r_Vo2To_vbs_true"] for it in items], dtype=np.float32)
        else:
            q = np.tile(np.array([0, 0, 0, 1], np.float32), (len(items), 1))
            t = np.zeros((len(items), 3), np.float32)
        self._finish(dataset_dir, files, q, t, config, "SPEED")


def write_synthetic_urso(dataset_dir, n_train=4, n_val=2, n_test=2, width=1280, height=960, seed=0):
    """A tiny URSO-format dataset of random frames (tests / BASELINE config 1: no real data is reachable here)."""
    rng = np.random.RandomState(seed)
    os.makedirs(dataset_dir, exist_ok=True)
    for subset, n in (("train", n_train), ("val", n_val), ("test", n_test)):
        names, rows = [], []
        for i in range(n):
            name = f"{subset}_{i}_rgb.png"
            img = rng.randint(0, 256, (height // 8, width // 8, 3), dtype=np.uint8)
            cv2.imwrite(os.path.join(dataset_dir, name), cv2.resize(img, (width, height), interpolation=cv2.INTER_NEAREST))
            q = rng.randn(4)
            q /= np.linalg.norm(q)
            rows.append([rng.uniform(-2, 2), rng.uniform(-2, 2), rng.uniform(5, 40)] + list(q))
            names.append(name)
        with open(os.path.join(dataset_dir, subset + "_images.csv"), "w") as f:
            f.write("\n".join(names) + "\n")
        with open(os.path.join(dataset_dir, subset + "_poses_gt.csv"), "w") as f:
            f.write("x,y,z,q1,q2,q3,q4\n" + "\n".join(",".join(repr(float(v)) for v in r) for r in rows) + "\n")


# ------------------------------------------------------------------------------------------------ rotation augmentation
def _rot_xyz_deg(pitch, yaw, roll):
    """R = Rz(roll) Ry(yaw) Rx(pitch), angles in degrees (the matrix of se3lib.euler2SO3_left, se3lib.py:38-51)."""
    a, b, c = np.deg2rad([pitch, yaw, roll])
    rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
    rz = np.array([[np.cos(c), -np.sin(c), 0], [np.sin(c), np.cos(c), 0], [0, 0, 1]])
    return rz @ ry @ rx


def _rot_to_quat_jpl(R):
    """Rotation matrix -> JPL (left-handed) quaternion [x, y, z, w], the four-branch formula of Trawny & Roumeliotis that
    se3lib.SO32quat (se3lib.py:77-115) uses; written around the largest diagonal term."""
    tr = R[0, 0] + R[1, 1] + R[2, 2]
    if tr > 0:
        z = np.sqrt(tr + 1.0) * 2
        return np.array([(R[1, 2] - R[2, 1]) / z, (R[2, 0] - R[0, 2]) / z, (R[0, 1] - R[1, 0]) / z, 0.25 * z])
    i = 0 if (R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]) else (1 if R[1, 1] > R[2, 2] else 2)
    j, k = (i + 1) % 3, (i + 2) % 3
    z = np.sqrt(1.0 + 2 * R[i, i] - tr) * 2
    q = np.zeros(4)
    q[i] = 0.25 * z
    q[j] = (R[i, j] + R[j, i]) / z
    q[k] = (R[i, k] + R[k, i]) / z
    q[3] = (R[j, k] - R[k, j]) / z
    return q


def _quat_mult_jpl(a, b):
    """a (x) b in the JPL convention of se3lib.quat_mult (se3lib.py:164-179), normalised."""
    ax, ay, az, aw = a
    L = np.array([[aw, az, -ay, ax], [-az, aw, ax, ay], [ay, -ax, aw, az], [-ax, -ay, -az, aw]])
    r = L @ np.asarray(b, dtype=np.float64)
    return r / np.linalg.norm(r)


def warp_by_camera_rotation(image, t, q, K, pitch_yaw_roll_deg):
    """The image / pose update shared by utils.rotate_cam and utils.rotate_image (utils.py:30-86): the camera is rotated by
    R, the image is re-projected with the homography K R K^-1 and the pose becomes t R^T, q_R (x) q.
    Parity trap: the reference passes cv2.WARP_INVERSE_MAP as the FOURTH positional argument of cv2.warpPerspective
    (utils.py:47,76), which is `dst`, not `flags` -- so the flag is never set and M is applied as a forward map with
    bilinear interpolation.  That behaviour is what the golden vectors (tests/golden/rotaug_golden.npz) pin."""
    R = _rot_xyz_deg(*pitch_yaw_roll_deg)
    K = np.asarray(K, dtype=np.float64)
    M = K @ R @ np.linalg.inv(K)
    h, w = image.shape[:2]
    warped = cv2.warpPerspective(image, M, (w, h))
    t_new = np.asarray(t, dtype=np.float64) @ R.T
    q_new = _quat_mult_jpl(_rot_to_quat_jpl(R), q)
    return warped, t_new, q_new


def rotate_cam(image, t, q, K, magnitude, rng=np.random):
    """Random camera-orientation perturbation, +-magnitude/2 degrees per Euler angle (utils.py:30-57)."""
    return warp_by_camera_rotation(image, t, q, K, (rng.rand(3) - 0.5) * magnitude)


def rotate_image(image, t, q, K, rng=np.random):
    """Random in-plane (roll) rotation of +-85 degrees (utils.py:59-86)."""
    return warp_by_camera_rotation(image, t, q, K, (0.0, 0.0, (rng.rand(1) - 0.5)[0] * 170))


# ------------------------------------------------------------------------------------------------ batch generator
def load_image_gt(dataset, config, image_id, device_aug=False):
    """(image uint8 [H,W,3] resized+padded, image_meta, loc, ori) -- net.py:358-456 without the augmentations."""
    image = dataset.load_image(image_id)
    loc = dataset.load_location(image_id)
    ori = dataset.load_quaternion(image_id) if config.REGRESS_ORI else dataset.load_orientation_encoded(image_id)
    rot_aug, rot_img = getattr(config, "ROT_AUG", False), getattr(config, "ROT_IMAGE_AUG", False)
    if getattr(config, "SIM2REAL_AUG", False) and not device_aug:
        # luma written back into the uint8 image (net.py:391-394).  The stochastic imgaug pipeline (net.py:395-406) only
        # exists on the device (csrc/augment.cu, applied when the uploaded batch is swapped in): with device_aug the
        # frame is left untouched here and the kernel does the luma step too
        gray = (0.2126 * image[:, :, 0] + 0.7152 * image[:, :, 1] + 0.0722 * image[:, :, 2]).astype(np.uint8)
        image = np.stack([gray] * 3, -1)
    if rot_aug or rot_img:
        # net.py:409-438: one of the two warps, mutually exclusive on a coin flip; the pose follows the camera and the
        # orientation soft label is re-encoded from the rotated quaternion
        assert config.REGRESS_LOC and getattr(config, "ORIENTATION_PARAM", "quaternion") == "quaternion"
        dice = np.random.rand(1)
        warp = None
        if rot_aug and dice > 0.5:
            warp = lambda im, t, q: rotate_cam(im, t, q, dataset.camera.K, 20)
        elif rot_img and dice <= 0.5:
            warp = lambda im, t, q: rotate_image(im, t, q, dataset.camera.K)
        if warp is not None:
            image, loc, q_new = warp(image, loc, dataset.load_quaternion(image_id))
            ori = q_new if config.REGRESS_ORI else dataset.encoder.encode(np.asarray(q_new)[None])[0]
    original_shape = image.shape
    image, window, scale, _padding, _crop = resize_image(image, min_dim=config.IMAGE_MIN_DIM, min_scale=config.IMAGE_MIN_SCALE,
                                                         max_dim=config.IMAGE_MAX_DIM, mode=config.IMAGE_RESIZE_MODE)
    meta = compose_image_meta(image_id, original_shape, image.shape, window, scale)
    return image, meta, loc, ori


def data_generator(dataset, config, shuffle=True, batch_size=1, raw_uint8=False, device_aug=False, rank=0, world=1,
                   seed=0):
    """Infinite generator of ([images, image_meta, gt_locs, gt_oris], []) like net.data_generator (net.py:458-559).
    raw_uint8=True yields un-molded uint8 images (the engine subtracts MEAN_PIXEL on the GPU).
    world > 1 (one process per GPU): every pass over the dataset uses ONE permutation shared by all ranks
    (dp.shard_indices, seeded by `seed` and the pass number) and each rank takes its strided share, so an epoch visits
    every image once across the job instead of `world` overlapping private shuffles."""
    if getattr(config, "SIM2REAL_AUG", False) and not (raw_uint8 and device_aug):
        # the stochastic imgaug pipeline (net.py:395-406) exists only in the device kernel (csrc/augment.cu), which
        # needs the raw uint8 feed: refuse to train silently without it
        raise ValueError("SIM2REAL_AUG needs the raw uint8 feed with the on-device augmentation "
                         "(data_generator(raw_uint8=True, device_aug=True))")
    b, image_index, error_count, n_pass = 0, -1, 0, 0
    all_ids = np.copy(dataset.image_ids)
    if world > 1:
        from .dp import shard_indices
        image_ids = all_ids[shard_indices(len(all_ids), rank, world, seed, 0)] if shuffle else all_ids[rank::world]
        assert len(image_ids) > 0, "fewer images than ranks"
    else:
        image_ids = all_ids
    n_ori = 4 if config.REGRESS_ORI else config.ORI_BINS_PER_DIM ** 3
    while True:
        try:
            image_index = (image_index + 1) % len(image_ids)
            if shuffle and image_index == 0:
                if world > 1:
                    image_ids = all_ids[shard_indices(len(all_ids), rank, world, seed, n_pass)]
                    n_pass += 1
                else:
                    np.random.shuffle(image_ids)
            image_id = image_ids[image_index]
            image, meta, loc, ori = load_image_gt(dataset, config, image_id, device_aug=device_aug and raw_uint8)
            if b == 0:
                metas = np.zeros((batch_size,) + meta.shape, dtype=meta.dtype)
                images = np.zeros((batch_size,) + image.shape, dtype=np.uint8 if raw_uint8 else np.float32)
                locs = np.zeros((batch_size, 3), dtype=np.float32)
                oris = np.zeros((batch_size, n_ori), dtype=np.float32)
            metas[b] = meta
            images[b] = image if raw_uint8 else mold_image(image, config)
            locs[b], oris[b] = loc, ori
            b += 1
            if b >= batch_size:
                yield [images, metas, locs, oris], []
                b = 0
        except (GeneratorExit, KeyboardInterrupt):
            raise
        except NotImplementedError:
            raise
        except Exception:
            logging.exception("Error processing image {}".format(dataset.image_info[image_id]))
            error_count += 1
            if error_count > 5:
                raise


class ParallelLoader:
    """Background batch producers for the train loop -- the counterpart of `fit_generator(workers=cpu_count,
    use_multiprocessing=True, max_queue_size=100)` (net.py:1147-1163).  `workers` threads each run their own
    `data_generator` over a disjoint share of the dataset (worker w of rank r takes share r*workers + w of
    world*workers) and put finished batches into one bounded queue; image decoding, resizing and the warps are OpenCV /
    numpy calls that release the GIL, so the threads really overlap.  Iterating yields the same ([images, meta, locs,
    oris], []) tuples as the generator.  A single-threaded generator feeds ~60 images/s of 1280x960 PNGs -- two orders
    below what one B200 consumes at the bench shape."""

    def __init__(self, dataset, config, batch_size, workers=None, queue_size=16, rank=0, world=1, seed=0, **gen_kwargs):
        import queue
        import threading
        n = len(dataset.image_ids)
        if workers is None:
            workers = max(1, (os.cpu_count() or 2) // max(1, world) - 1)
        workers = max(1, min(int(workers), max(1, n // max(1, world))))
        self.workers = workers
        self._q = queue.Queue(maxsize=max(2, min(int(queue_size), 100)))
        self._stop = threading.Event()
        self._err = None
        self._threads = []
        for w in range(workers):
            gen = data_generator(dataset, config, batch_size=batch_size, rank=rank * workers + w, world=world * workers,
                                 seed=seed, **gen_kwargs)
            t = threading.Thread(target=self._work, args=(gen,), daemon=True)
            t.start()
            self._threads.append(t)

    def _work(self, gen):
        try:
            for batch in gen:
                while not self._stop.is_set():
                    try:
                        self._q.put(batch, timeout=0.1)
                        break
                    except Exception:       # queue.Full
                        continue
                if self._stop.is_set():
                    return
        except BaseException as e:           # surfaced to the consumer: the train loop must not hang on a dead worker
            self._err = e
            self._stop.set()

    def __iter__(self):
        return self

    def __next__(self):
        import queue
        while True:
            if self._err is not None:
                raise self._err
            try:
                return self._q.get(timeout=0.5)
            except queue.Empty:
                if self._stop.is_set() and self._err is None:
                    raise StopIteration

    def close(self):
        self._stop.set()
        for t in self._threads:
            t.join(timeout=2.0)
