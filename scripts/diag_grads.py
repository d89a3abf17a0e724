"""Diagnostic: per-tensor gradient error of the engine vs the fp64 oracle (prints in graph order)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import ursonet_oracle as O
from tests.test_gpu_model import make_cfg, make_batch, load_oracle_weights
from ursonet_b200.engine import Engine

backbone = sys.argv[1] if len(sys.argv) > 1 else "resnet18"
h, w = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (128, 192)
pre = (sys.argv[4] != "plain") if len(sys.argv) > 4 else True
cfg = make_cfg(backbone, True, h, w)
B = 2
p64 = O.init_weights(cfg, seed=2, pretrained_like=pre)
eng = Engine(cfg, B, training=True)
load_oracle_weights(eng, p64)
img, gt_loc, gt_ori = make_batch(cfg, B, seed=3)
eng.img_u8.copy_(img); eng.gt_loc.copy_(gt_loc); eng.gt_ori.copy_(gt_ori)
eng.train_step(1e-3, use_graph=False)
torch.cuda.synchronize()
quant = os.environ.get("QUANT", "1") == "1"
newp, info = O.train_step(p64, {}, (O.mold_image(img), gt_loc.double(), gt_ori.double()), cfg, lr=1e-3, quant=quant)
print("losses", eng.losses.tolist(), info["loc_loss"].item(), info["ori_loss"].item())
for name, gref in info["grads"].items():
    got = eng.params.view(name, eng.grads).double().cpu()
    n = gref.norm().item()
    e = (got - gref).norm().item() / max(n, 1e-30)
    cos = (got * gref).sum().item() / max(got.norm().item() * n, 1e-30)
    print(f"{name:32s} relerr {e:8.4f} cos {cos:8.5f} |ref| {n:10.4g} |got| {got.norm().item():10.4g}")

print("---- forward taps: rms rel err vs quant oracle | vs fp64 oracle")
tq, t64 = {}, {}
with torch.no_grad():
    O.forward(p64, O.mold_image(img), cfg, tq, quant=True)
    O.forward(p64, O.mold_image(img), cfg, t64, quant=False)
for name in tq:
    if name not in eng.act:
        continue
    got = eng.act[name].double().cpu()
    eq = (got - tq[name]).norm().item() / max(tq[name].norm().item(), 1e-30)
    e64 = (got - t64[name]).norm().item() / max(t64[name].norm().item(), 1e-30)
    mq = (got - tq[name]).abs().max().item() / max(tq[name].abs().max().item(), 1e-30)
    print(f"{name:28s} rms_q {eq:9.5f} max_q {mq:9.5f} rms_64 {e64:9.5f}")
