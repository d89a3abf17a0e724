"""Command line of the B200 build: same sub-commands and flags as the reference's pose_estimator.py
(pose_estimator.py:764-973).

  python -m ursonet_b200.pose_estimator train    --dataset D --weights {none|last|<run dir>} [flags]
  python -m ursonet_b200.pose_estimator evaluate --dataset D --weights {last|<run dir>}      [flags]
  python -m ursonet_b200.pose_estimator test     --dataset D --weights ...                   (per-image errors, no plots)

Under `torchrun --nproc-per-node N` the train command is data parallel: every rank reads its own shard order of the
dataset and gradients are all-reduced over NCCL (the reference's GPU_COUNT is never changed by its CLI,
pose_estimator.py:870; its multi-GPU wrapper is a commented-out stub, net.py:694-697).
Plots, video and the ESA submission writer (pose_estimator.py:42-320, 463-745) are outside the hot path and not built.
"""
import argparse
import os
import sys

import numpy as np

ROOT_DIR = os.path.abspath("./")
DEFAULT_LOGS_DIR = os.path.join(ROOT_DIR, "logs")
DATA_DIR = os.path.join(ROOT_DIR, "datasets")
OrientationParamOptions = ["quaternion", "euler_angles", "angle_axis"]


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument("command", metavar="<command>", help="'train' or 'evaluate'")
    p.add_argument("--backbone", required=False, default="resnet50", help="Backbone architecture")
    p.add_argument("--dataset", required=True, help="Dataset name")
    p.add_argument("--epochs", required=False, default=100, type=int, help="Number of epochs")
    p.add_argument("--image_scale", required=False, default=1.0, type=float, help="Resize scale")
    p.add_argument("--ori_weight", required=False, default=1.0, type=float, help="Loss weight")
    p.add_argument("--loc_weight", required=False, default=1.0, type=float, help="Loss weight")
    p.add_argument("--bottleneck", required=False, default=32, type=int, help="Bottleneck width")
    p.add_argument("--branch_size", required=False, default=1024, type=int, help="Branch input size")
    p.add_argument("--learn_rate", required=False, default=0.001, type=float, help="Learning rate")
    p.add_argument("--batch_size", required=False, default=4, type=int, help="Number of images per GPU")
    for flag in ("rot_aug", "rot_image_aug", "regress_keypoints", "sim2real", "clr", "f16", "square_image"):
        p.add_argument("--" + flag, dest=flag, action="store_true")
        p.set_defaults(**{flag: False})
    p.add_argument("--classify_ori", dest="regress_ori", action="store_false")
    p.add_argument("--regress_ori", dest="regress_ori", action="store_true")
    p.set_defaults(regress_ori=False)
    p.add_argument("--classify_loc", dest="regress_loc", action="store_false")
    p.add_argument("--regress_loc", dest="regress_loc", action="store_true")
    p.set_defaults(regress_loc=True)
    p.add_argument("--ori_param", required=False, default="quaternion", help="'quaternion' 'euler_angles' 'angle_axis'")
    p.add_argument("--ori_resolution", required=False, default=16, type=int, help="Number of bins assigned to each angle")
    p.add_argument("--weights", required=True, help="Path to weights file / run directory, 'last' or 'none'")
    p.add_argument("--logs", required=False, default=DEFAULT_LOGS_DIR, help="Logs and checkpoints directory")
    p.add_argument("--image", required=False, help="Image to evaluate")
    p.add_argument("--video", required=False, help="Video to evaluate")
    # additions of this build
    p.add_argument("--data_dir", required=False, default=DATA_DIR, help="Root of the datasets (default ./datasets)")
    p.add_argument("--steps_per_epoch", required=False, default=None, type=int, help="Override min(1000, N/batch)")
    return p


def make_config(args):
    """Flags -> Config exactly as pose_estimator.py:815-872."""
    from .config import Config
    from .data import SpeedCamera, UrsoCamera
    assert args.ori_param in OrientationParamOptions
    config = Config()
    config.ORIENTATION_PARAM = args.ori_param
    config.ORI_BINS_PER_DIM = args.ori_resolution
    config.NAME = args.dataset
    config.EPOCHS = args.epochs
    config.NR_DENSE_LAYERS = 1
    config.LEARNING_RATE = args.learn_rate
    config.BOTTLENECK_WIDTH = args.bottleneck
    config.BRANCH_SIZE = args.branch_size
    config.BACKBONE = args.backbone
    config.ROT_AUG = args.rot_aug
    config.F16 = args.f16
    config.SIM2REAL_AUG = args.sim2real
    config.CLR = args.clr
    config.ROT_IMAGE_AUG = args.rot_image_aug
    config.OPTIMIZER = "SGD"
    config.REGRESS_ORI = args.regress_ori
    config.REGRESS_LOC = args.regress_loc
    config.REGRESS_KEYPOINTS = args.regress_keypoints
    config.LOSS_WEIGHTS["loc_loss"] = args.loc_weight
    config.LOSS_WEIGHTS["ori_loss"] = args.ori_weight
    config.IMAGE_RESIZE_MODE = "square" if args.square_image else "pad64"
    cam = SpeedCamera if args.dataset == "speed" else UrsoCamera
    config.IMAGE_MAX_DIM = round(cam.width * args.image_scale)
    if config.IMAGE_MAX_DIM % 64 > 0:
        raise Exception("Scale problem. Image maximum dimension must be dividable by 2 at least 6 times.")
    height_scaled = round(cam.height * args.image_scale)
    config.IMAGE_MIN_DIM = height_scaled - height_scaled % 64 + 64 if height_scaled % 64 > 0 else height_scaled
    config.IMAGES_PER_GPU = args.batch_size if args.command == "train" else 1
    config.update()
    return config


def load_sets(args, config, subsets):
    from . import data as D
    dataset_dir = os.path.join(args.data_dir, args.dataset)
    out = []
    for subset in subsets:
        ds = D.Speed() if args.dataset == "speed" else D.Urso()
        ds.load_dataset(dataset_dir, config, subset)
        out.append(ds)
    return out


def decode_orientation(model, dataset, ori):
    """Network output -> quaternion (pose_estimator.py:376-409)."""
    from . import labels
    if model.config.REGRESS_ORI:
        return np.asarray(ori, dtype=np.float64)
    return labels.quat_weighted_avg(dataset.ori_histogram_map, labels.stable_softmax(np.asarray(ori, dtype=np.float64)))


def evaluate(model, dataset, out_dir="."):
    """Mean location error, angular error (deg) and ESA score over a dataset (pose_estimator.py:321-459)."""
    import pandas as pd
    loc_err_acc, ori_err_acc, esa_acc, dist_acc = [], [], [], []
    for image_id in dataset.image_ids:
        loc_gt = np.asarray(dataset.load_location(image_id), dtype=np.float64)
        q_gt = np.asarray(dataset.load_quaternion(image_id), dtype=np.float64)
        res = model.detect([dataset.load_image(image_id)], verbose=0)[0]
        q_est = decode_orientation(model, dataset, res["ori"])
        ang = 2 * np.arccos(min(1.0, abs(float(np.dot(q_est, q_gt)))))
        loc_err = float(np.linalg.norm(res["loc"] - loc_gt))
        ori_err_acc.append(ang * 180 / np.pi)
        loc_err_acc.append(loc_err)
        esa_acc.append(loc_err / np.linalg.norm(loc_gt) + ang)
        dist_acc.append(loc_gt[2])
        print("Image ID:", image_id, " Loc Error: ", loc_err, " Ori Error: ", ori_err_acc[-1])
    print("Mean est. location error: ", np.mean(loc_err_acc))
    print("Mean est. orientation error: ", np.mean(ori_err_acc))
    print("ESA score: ", np.mean(esa_acc))
    pd.DataFrame(np.asarray(ori_err_acc)).to_csv(os.path.join(out_dir, "ori_err.csv"))
    pd.DataFrame(np.asarray(loc_err_acc)).to_csv(os.path.join(out_dir, "loc_err.csv"))
    pd.DataFrame(np.asarray(dist_acc)).to_csv(os.path.join(out_dir, "dists_err.csv"))
    return float(np.mean(loc_err_acc)), float(np.mean(ori_err_acc)), float(np.mean(esa_acc))


def train(model, dataset_train, dataset_val, config, steps_per_epoch=None):
    """pose_estimator.py:747-758."""
    config.STEPS_PER_EPOCH = steps_per_epoch or min(1000, int(len(dataset_train.image_ids) / config.BATCH_SIZE))
    config.write_to_file(os.path.join(model.log_dir, "config_" + str(model.epoch) + ".json"))
    print("Training")
    return model.train(dataset_train, dataset_val, learning_rate=config.LEARNING_RATE, epochs=config.EPOCHS, layers="all")


def main(argv=None):
    args = build_parser().parse_args(argv)
    print("Command: ", args.command)
    print("Dataset: ", args.dataset)
    print("Logs: ", args.logs)
    import torch
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and not torch.distributed.is_initialized():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        torch.distributed.init_process_group("nccl")
    config = make_config(args)
    config.display()
    from . import net
    os.makedirs(args.logs, exist_ok=True)
    model = net.UrsoNet(mode="training" if args.command == "train" else "inference", config=config, model_dir=args.logs)
    w = args.weights.lower()
    if w in ("coco", "imagenet"):
        model.get_imagenet_weights(config.BACKBONE)
    elif w in ("soyuz_hard", "dragon_hard", "speed"):
        model.get_urso_weights(args.weights)
    elif w == "last":
        _, weights_path = model.find_last()
        model.load_weights(weights_path, weights_path, by_name=True)
    elif w != "none":
        path = args.weights
        if os.path.isdir(os.path.join(args.logs, args.weights)):
            _, path = model.get_last_checkpoint(args.weights)
        model.load_weights(path, path, by_name=True)
    if args.command == "train":
        tr, va = load_sets(args, config, ["train_no_val", "val"] if args.dataset == "speed" else ["train", "val"])
        train(model, tr, va, config, args.steps_per_epoch)
    elif args.command in ("evaluate", "test"):
        (ds,) = load_sets(args, config, ["val"] if args.dataset == "speed" else ["test"])
        evaluate(model, ds, out_dir=args.logs)
    elif args.command == "submit":
        raise NotImplementedError("the ESA submission writer (pose_estimator.py:217-320) is outside the hot path")
    else:
        print("wrong command")
    return model


if __name__ == "__main__":
    main()
