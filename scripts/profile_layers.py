"""Run selected launches of the bench workload inside a cudaProfilerStart/Stop range (for ncu --profile-from-start off):
  ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/x \
      python scripts/profile_layers.py conv_dgrad:res2a_out conv_fwd:res2a_branch2c ..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

class A: pass
a = A(); a.batch = int(os.environ.get("BATCH", "32")); a.backbone = os.environ.get("BACKBONE", "resnet50")
a.width, a.height, a.ori_resolution, a.regress_ori = 960, 600, 16, False
from ursonet_b200.engine import Engine
from ursonet_b200 import lib as _lib
if os.environ.get("RESIDUAL_MMA", "1") == "0":
    _lib.load().urso_set_residual_mma(0)
cfg = bench.make_cfg(a)
eng = Engine(cfg, a.batch, training=True)
img, loc, ori = bench.synth_batch(cfg, a.batch, 0)
eng.img_u8.copy_(img); eng.gt_loc.copy_(loc); eng.gt_ori.copy_(ori)
eng.train_step(1e-3, use_graph=False)
eng._phase_train()
torch.cuda.synchronize()
want = [tuple(s.split(":")) for s in sys.argv[1:]]
ops = {(o.kind, o.name): o for o in list(eng.ops_fwd) + list(eng.ops_bwd) if hasattr(o, "kind")}
eng._serial = True
for key in want:
    ops[key]()          # warm (cold-start effects out of the capture)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for key in want:
    ops[key]()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled", want)
