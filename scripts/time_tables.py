"""Isolated CUDA-event timing of the multi-tensor job-table launches of the bench workload (RN-50, 640x960, B=32)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

class A: pass
a = A(); a.batch = 32; a.backbone = "resnet50"; a.width, a.height, a.ori_resolution, a.regress_ori = 960, 600, 16, False
from ursonet_b200.engine import Engine, BN_EPS
cfg = bench.make_cfg(a)
eng = Engine(cfg, a.batch, training=True)
img, loc, ori = bench.synth_batch(cfg, a.batch, 0)
eng.img_u8.copy_(img); eng.gt_loc.copy_(loc); eng.gt_ori.copy_(ori)
eng.train_step(1e-3, use_graph=False)
torch.cuda.synchronize()

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

print("bn_fold table   %.4f ms (%d jobs)" % (timeit(lambda: eng._bn_table.launch(BN_EPS)), eng._bn_table.n))
print("stage table A   %.4f ms (%d jobs, %d blocks)" % (timeit(eng._stage_table_a.launch), eng._stage_table_a.n, eng._stage_table_a.total))
print("stage table B   %.4f ms (%d jobs, %d blocks)" % (timeit(eng._stage_table.launch), eng._stage_table.n, eng._stage_table.total))
for op in eng.ops_pgrad:
    print("pgrad %-10s %.4f ms" % (op.name, timeit(op)))
print("zero arena      %.4f ms (%d MB)" % (timeit(lambda: eng.zero_arena.zero_()), eng.zero_arena.numel() * 4 // 2**20))
print("stem stage      %.4f ms" % timeit(eng._stage_input))
print("update          %.4f ms" % timeit(eng._phase_update))
