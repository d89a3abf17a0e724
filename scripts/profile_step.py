"""One eager train step of the bench workload inside a cudaProfilerStart/Stop range (for ncu --profile-from-start off)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--backbone", default="resnet50")
ap.add_argument("--width", type=int, default=960)
ap.add_argument("--height", type=int, default=600)
ap.add_argument("--ori_resolution", type=int, default=16)
ap.add_argument("--steps", type=int, default=1)
a = ap.parse_args()
from ursonet_b200.engine import Engine
cfg = bench.make_cfg(a)
eng = Engine(cfg, a.batch, training=True)
img, loc, ori = bench.synth_batch(cfg, a.batch, 0)
eng.img_u8.copy_(img); eng.gt_loc.copy_(loc); eng.gt_ori.copy_(ori)
eng.train_step(1e-3, use_graph=False)
eng.train_step(1e-3, use_graph=False)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(a.steps):
    eng.train_step(1e-3, use_graph=False)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled", a.steps, "step(s);", eng.count_launches(True), "launches of liburso_b200 per step")
