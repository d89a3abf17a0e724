"""`UrsoNet` facade: the reference's `net.UrsoNet` method surface (net.py:566-1259) over the B200 engine.

Drop-in for the Keras path of `pose_estimator.py`: same constructor, `build / compile / set_trainable / train /
detect / mold_inputs / load_weights / find_last / get_last_checkpoint / set_log_dir`, same attributes read by the CLI
(`config`, `epoch`, `log_dir`, `checkpoint_path`).  `keras_model` is a small handle exposing `predict()` and by-name
weight access -- there is no Keras.  Everything numeric runs in `engine.Engine` through liburso_b200.so.

Deviations: checkpoints are `.npz` keyed by Keras weight names by default; Keras `.h5` weight files are
read and (config.CHECKPOINT_FORMAT = 'h5') written by the pure-Python ursonet_b200/hdf5.py; pretrained-weight downloads
raise (no network); multi-GPU is real (one process per GPU under torchrun, NCCL all-reduce of the flat gradient arena)
instead of the reference's commented-out stub.
"""
import datetime
import os
import re
import time

import numpy as np
import torch

from . import data as D
from .engine import Engine


def log(text, array=None):
    """net.py:46-57."""
    if array is not None:
        text = text.ljust(25)
        text += ("shape: {:20}  min: {:10.5f}  max: {:10.5f}  {}".format(
            str(array.shape), array.min() if array.size else "", array.max() if array.size else "", array.dtype))
    print(text)


class _ModelHandle:
    """Stand-in for `keras_model`: predict() + weight lookup by Keras name."""

    def __init__(self, owner):
        self._o = owner

    def predict(self, molded_images, verbose=0):
        return self._o._predict(np.asarray(molded_images))

    def get_weights_by_name(self):
        return self._o.engine.params.state_dict()

    @property
    def layers(self):
        return sorted({n.split("/")[0] for n in self._o.engine.params.names()})


class BatchLogger:
    """Per-batch loss history returned by train() (net.py:1106-1115)."""

    def __init__(self):
        self.ori_loss_acc, self.loc_loss_acc = [], []


class UrsoNet:
    def __init__(self, mode, config, model_dir):
        assert mode in ["training", "inference"]
        self.mode, self.config, self.model_dir = mode, config, model_dir
        self.set_log_dir()
        self.keras_model = self.build(mode=mode, config=config)

    # ------------------------------------------------------------------ build / compile
    def build(self, mode, config):
        assert mode in ["training", "inference"]
        if getattr(config, "F16", False):
            # Reference: K.set_floatx('float16') + epsilon 1e-4 (net.py:590-593): pure fp16 variables and maths, no loss
            # scaling, no master copy.  This build serves the flag with its one 16-bit engine: bf16 storage / MMA operands
            # (same bytes and tensor-core rate as fp16, fp32 exponent range so no loss scaling is needed), fp32
            # accumulation and fp32 master weights; the fp16 Adam epsilon (1e-4) is honoured (Engine.set_hyper).
            # DESIGN.md section 8 records the decision.
            log("--f16: 16-bit storage is bf16 in this build (fp32 accumulate + fp32 master weights), Adam eps = 1e-4")
        world = torch.distributed.get_world_size() if torch.distributed.is_initialized() else 1
        self.world = world
        per_gpu = int(config.BATCH_SIZE) // max(1, int(getattr(config, "GPU_COUNT", 1)))
        self.engine = Engine(config, per_gpu if mode == "training" else int(config.BATCH_SIZE), training=(mode == "training"),
                             world_size=world)
        self._lr, self._momentum = config.LEARNING_RATE, config.LEARNING_MOMENTUM
        return _ModelHandle(self)

    def compile(self, learning_rate, momentum):
        """Loss weights, L2 regulariser and optimizer are part of the engine's plan (net.py:973-1028); this records
        the hyper-parameters the update kernel reads from device memory."""
        if self.config.OPTIMIZER not in ("SGD", "ADAM", "Adam", "adam"):
            raise ValueError("unknown optimizer " + str(self.config.OPTIMIZER))
        self._lr, self._momentum = learning_rate, momentum

    def set_trainable(self, layer_regex, keras_model=None, indent=0, verbose=1):
        n = self.engine.params.set_trainable(layer_regex)
        if verbose > 0:
            log("Selecting layers to train: {} parameter chunks match '{}'".format(n, layer_regex))

    # ------------------------------------------------------------------ checkpoints
    def set_log_dir(self, model_path=None):
        """Log directory + epoch counter; epoch parsed from '..._<epoch:04d>.<ext>' (net.py:944-967)."""
        self.epoch = 0
        now = datetime.datetime.now()
        self.log_dir = os.path.join(self.model_dir, "{}{:%Y%m%dT%H%M}".format(self.config.NAME.lower(), now))
        if model_path:
            m = re.search(r"_(\d{4})\.(h5|npz)$", model_path)
            if m:
                self.log_dir = os.path.dirname(model_path)
                self.epoch = int(m.group(1))
        ext = "h5" if getattr(self.config, "CHECKPOINT_FORMAT", "npz") == "h5" else "npz"
        self.checkpoint_path = os.path.join(self.log_dir, "weights_{}_*epoch*.{}".format(self.config.NAME.lower(), ext))
        self.checkpoint_path = self.checkpoint_path.replace("*epoch*", "{epoch:04d}")

    def get_last_checkpoint(self, model_name):
        dir_names = next(os.walk(self.model_dir))[1]
        assert model_name in dir_names
        model_path = os.path.join(self.model_dir, model_name)
        cks = sorted(f for f in next(os.walk(model_path))[2] if f.startswith("weights"))
        return (model_path, os.path.join(model_path, cks[-1])) if cks else (model_path, None)

    def find_last(self):
        if not os.path.isdir(self.model_dir):
            return None, None
        key = self.config.NAME.lower()
        dir_names = sorted(d for d in next(os.walk(self.model_dir))[1] if d.startswith(key))
        for d in reversed(dir_names):     # the newest directory may be this run's own, still empty
            dir_name = os.path.join(self.model_dir, d)
            cks = sorted(f for f in next(os.walk(dir_name))[2] if f.startswith("weights"))
            if cks:
                return dir_name, os.path.join(dir_name, cks[-1])
        return (os.path.join(self.model_dir, dir_names[-1]), None) if dir_names else (None, None)

    @staticmethod
    def _read_weight_file(path):
        """'.npz' (this build's own format) or a Keras HDF5 weight file (`model.save_weights`, or a full `model.save`
        file with a 'model_weights' group, net.py:831-832), read by ursonet_b200/hdf5.py (no h5py needed)."""
        if path.endswith(".npz"):
            with np.load(path) as z:
                return {k: z[k] for k in z.files}
        from . import hdf5
        return hdf5.keras_to_state_dict(hdf5.read_keras_weights(path))

    def load_weights(self, weights_in_path, weights_out_path, by_name=False, exclude=None):
        """By-name load with an exclude list (net.py:816-852); then set_log_dir(weights_out_path)."""
        if weights_in_path is None:
            raise ValueError("no weights file given")
        sd = self._read_weight_file(weights_in_path)
        loaded = self.engine.params.load_state_dict(sd, by_name=True, exclude=exclude)
        print("Loaded {} weight tensors from {}".format(len(loaded), weights_in_path))
        self.set_log_dir(weights_out_path)

    def save_weights(self, path):
        """Weights only, like ModelCheckpoint(save_weights_only=True) (net.py:1120): '.npz', or a Keras-layout HDF5 file
        when the path ends in '.h5' (config.CHECKPOINT_FORMAT = 'h5' makes train() write those)."""
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        if path.endswith(".h5"):
            from . import hdf5
            hdf5.write_keras_weights(path, self.engine.params.state_dict())
        else:
            np.savez(path, **self.engine.params.state_dict())

    def get_imagenet_weights(self, backbone):
        raise RuntimeError("pretrained ImageNet weights must be downloaded (net.py:854-893): no network in this build; "
                           "pass --weights <dir with .npz/.h5> or --weights none")

    def get_urso_weights(self, name):
        raise RuntimeError("released URSO weights must be downloaded (net.py:895-940): no network in this build")

    # ------------------------------------------------------------------ training
    def _put_batch(self, inputs):
        images, _meta, gt_loc, gt_ori = inputs
        e = self.engine
        if images.dtype == np.uint8:
            e.set_input_kind("u8")
            e.img_u8.copy_(torch.from_numpy(images), non_blocking=True)
        else:
            e.set_input_kind("molded")
            e.img_f32.copy_(torch.from_numpy(np.ascontiguousarray(images, dtype=np.float32)), non_blocking=True)
        e.gt_loc.copy_(torch.from_numpy(np.ascontiguousarray(gt_loc, dtype=np.float32)), non_blocking=True)
        e.gt_ori.copy_(torch.from_numpy(np.ascontiguousarray(gt_ori, dtype=np.float32)), non_blocking=True)

    _pin = None
    _aug = None
    _aug_rng = None

    def _feed(self, inputs):
        """Pipelined feed of a uint8 batch: numpy -> pinned host buffers -> asynchronous H2D on the engine's copy stream
        (overlaps the step that is running); `engine.swap_in()` makes it current.  Returns False (and feeds
        synchronously) for molded fp32 inputs."""
        images, _meta, gt_loc, gt_ori = inputs
        e = self.engine
        if images.dtype != np.uint8:
            self._put_batch(inputs)
            return False
        e.set_input_kind("u8")
        if self._pin is None:
            self._pin = [torch.empty(tuple(t.shape), dtype=t.dtype).pin_memory() for t in (e.img_u8, e.gt_loc, e.gt_ori)]
        e.wait_upload()                      # the previous H2D has finished reading the pinned buffers
        self._pin[0].copy_(torch.from_numpy(images))
        self._pin[1].copy_(torch.from_numpy(np.ascontiguousarray(gt_loc, dtype=np.float32)))
        self._pin[2].copy_(torch.from_numpy(np.ascontiguousarray(gt_ori, dtype=np.float32)))
        e.upload_async(*self._pin)
        # sim2real augmentation runs on the device when the batch is swapped in: draw this batch's parameters now
        self._aug = None
        if getattr(self.config, "SIM2REAL_AUG", False):
            from . import augment
            if self._aug_rng is None:
                self._aug_rng = np.random.RandomState(getattr(self.config, "AUG_SEED", None))
            wins = np.asarray(_meta)[:, 7:11].astype(np.int32)      # image_meta: window (y1, x1, y2, x2), net.py:1278-1301
            self._aug = augment.draw_params(self._aug_rng, wins)
        return True

    def _lr_at(self, it, base_lr):
        cfg = self.config
        if not getattr(cfg, "CLR", False):
            return base_lr
        # triangular cyclical LR (clr_callback.py:104-111)
        cycle = np.floor(1 + it / (2 * cfg.CLR_STEP_SIZE))
        x = np.abs(it / cfg.CLR_STEP_SIZE - 2 * cycle + 1)
        return float(cfg.BASE_LEARNING_RATE + (cfg.MAX_LEARNING_RATE - cfg.BASE_LEARNING_RATE) * max(0.0, 1 - x))

    def train(self, train_dataset, val_dataset, learning_rate, epochs, layers, use_graph=True, raw_uint8=True):
        """fit_generator semantics (net.py:1068-1167): STEPS_PER_EPOCH train steps, then VALIDATION_STEPS forward-only
        steps, checkpoint (weights only) per epoch.  Returns the per-batch loss history."""
        assert self.mode == "training", "Create model in training mode."
        layer_regex = {
            "heads": r"(ori\_.*)|(loc\_.*)|(fpn\_.*)|(bottleneck_layer)",
            "3+": r"(res3.*)|(bn3.*)|(res4.*)|(bn4.*)|(res5.*)|(bn5.*)|(loc\_.*)|(ori\_.*)|(fpn\_.*)|(bottleneck_layer)",
            "4+": r"(res4.*)|(bn4.*)|(res5.*)|(bn5.*)|(loc\_.*)|(ori\_.*)|(fpn\_.*)|(bottleneck_layer)",
            "5+": r"(res5.*)|(bn5.*)|(loc\_.*)|(ori\_.*)|(fpn\_.*)|(bottleneck_layer)",
            "all": ".*",
        }
        layers = layer_regex.get(layers, layers)
        cfg, e = self.config, self.engine
        rank = torch.distributed.get_rank() if torch.distributed.is_initialized() else 0
        dev_aug = bool(raw_uint8 and getattr(cfg, "SIM2REAL_AUG", False))
        shard = dict(rank=rank, world=self.world, seed=int(getattr(cfg, "DATA_SEED", 0)))
        # background producers (the reference: fit_generator(workers=cpu_count, max_queue_size=100), net.py:1147-1163)
        nw = getattr(cfg, "DATA_WORKERS", None)
        train_gen = D.ParallelLoader(train_dataset, cfg, e.B, workers=nw, shuffle=True, raw_uint8=raw_uint8,
                                     device_aug=dev_aug, **shard)
        val_gen = D.ParallelLoader(val_dataset, cfg, e.B, workers=1, shuffle=True, raw_uint8=raw_uint8,
                                   device_aug=dev_aug, **shard) if len(val_dataset.image_ids) else None
        history = BatchLogger()
        log("\nStarting at epoch {}. LR={}\n".format(self.epoch, learning_rate))
        log("Checkpoint Path: {}".format(self.checkpoint_path))
        self.set_trainable(layers)
        self.compile(learning_rate, cfg.LEARNING_MOMENTUM)
        allreduce = (lambda g: torch.distributed.all_reduce(g)) if self.world > 1 else None
        ar_async = None      # Engine.train_step(allreduce_async=...) overlaps the all-reduce; measured no gain at 2 GPUs
        it = 0       # the reference creates a fresh CyclicLR per train() call: its iteration counter restarts (net.py:1126-1130)
        for epoch in range(self.epoch, epochs):
            t0 = time.time()
            inputs, _ = next(train_gen)
            piped = self._feed(inputs)
            for step in range(cfg.STEPS_PER_EPOCH):
                if piped:
                    e.swap_in(self._aug)
                e.train_step(self._lr_at(it, learning_rate), allreduce, use_graph, ar_async)
                if step + 1 < cfg.STEPS_PER_EPOCH:      # next batch: generator work and H2D overlap the running step
                    inputs, _ = next(train_gen)
                    piped = self._feed(inputs)
                loc_l, ori_l = e.losses.tolist()
                history.loc_loss_acc.append(loc_l)
                history.ori_loss_acc.append(ori_l)
                it += 1
            val = []
            # fit_generator(validation_steps=VALIDATION_STEPS) on an infinite generator (net.py:1158-1159)
            for _ in range(cfg.VALIDATION_STEPS if len(val_dataset.image_ids) else 0):
                inputs, _ = next(val_gen)
                if self._feed(inputs):       # same feed as training: the reference's val generator augments too
                    e.swap_in(self._aug)
                val.append(e.eval_losses())
            if rank == 0:
                n = cfg.STEPS_PER_EPOCH
                msg = "Epoch {}/{} - {:.1f}s - loc_loss: {:.4f} - ori_loss: {:.4f}".format(
                    epoch + 1, epochs, time.time() - t0, float(np.mean(history.loc_loss_acc[-n:])),
                    float(np.mean(history.ori_loss_acc[-n:])))
                if val:
                    msg += " - val_loc_loss: {:.4f} - val_ori_loss: {:.4f}".format(*np.mean(np.asarray(val), 0))
                print(msg)
                self.save_weights(self.checkpoint_path.format(epoch=epoch + 1))
        train_gen.close()
        if val_gen is not None:
            val_gen.close()
        self.epoch = max(self.epoch, epochs)
        return history

    # ------------------------------------------------------------------ inference
    def mold_inputs(self, images):
        """net.py:1169-1205: resize+pad, mean-subtract, image meta, windows."""
        cfg = self.config
        molded, metas, windows = [], [], []
        for image in images:
            m, window, scale, _p, _c = D.resize_image(image, min_dim=cfg.IMAGE_MIN_DIM, min_scale=cfg.IMAGE_MIN_SCALE,
                                                      max_dim=cfg.IMAGE_MAX_DIM, mode=cfg.IMAGE_RESIZE_MODE)
            m = D.mold_image(m, cfg)
            metas.append(D.compose_image_meta(0, image.shape, m.shape, window, scale))
            molded.append(m)
            windows.append(window)
        return np.stack(molded), np.stack(metas), np.stack(windows)

    def _predict(self, molded_images):
        e = self.engine
        assert molded_images.shape[0] == e.B, "batch must equal BATCH_SIZE"
        assert tuple(molded_images.shape[1:3]) == (e.H, e.W), \
            "molded image size {} does not match the configured IMAGE_SHAPE {}".format(molded_images.shape[1:3], (e.H, e.W))
        e.set_input_kind("molded")
        e.img_f32.copy_(torch.from_numpy(np.ascontiguousarray(molded_images, dtype=np.float32)))
        loc, ori = e.forward()
        return loc.float().cpu().numpy(), ori.float().cpu().numpy()

    def detect(self, images, verbose=0):
        """List of images -> [{'loc': (3,), 'ori': (n^3,) | (4,)}] (net.py:1207-1259)."""
        assert self.mode == "inference", "Create model in inference mode."
        assert len(images) == self.config.BATCH_SIZE, "len(images) must be equal to BATCH_SIZE"
        if verbose:
            log("Processing {} images".format(len(images)))
            for image in images:
                log("image", image)
        molded_images, image_metas, _windows = self.mold_inputs(images)
        image_shape = molded_images[0].shape
        for g in molded_images[1:]:
            assert g.shape == image_shape, \
                "After resizing, all images must have the same size. Check IMAGE_RESIZE_MODE and image sizes."
        if verbose:
            log("molded_images", molded_images)
            log("image_metas", image_metas)
        loc_pred, ori_pred = self.keras_model.predict(molded_images, verbose=0)
        return [{"loc": loc_pred[i], "ori": ori_pred[i]} for i in range(len(images))]
