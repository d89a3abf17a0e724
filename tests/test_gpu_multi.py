"""Gradient accumulation and the N-rank data-parallel step of the ENGINE (not of the oracle) against one rank on the
concatenated batch.  Frozen BatchNorm means no cross-sample statistics, so for the per-sample-mean losses (soft-label
CE, 1-|q.q|; loc_weight = 0 removes the batch-global rel_loss) the mean of the shard gradients IS the full-batch
gradient; what remains is fp32 atomic summation order.  The loss scale differs by an exact power of two between the
shard and the full batch, so the bf16 rounding points agree.  Tolerance: 2e-3 relative (Frobenius)."""
import os
import socket

import pytest
import torch

from oracle import ursonet_oracle as O
from tests.test_gpu_model import load_oracle_weights, make_batch, make_cfg

pytestmark = pytest.mark.gpu


def _cfg(classify=True):
    cfg = make_cfg("resnet18", classify)
    cfg.LOSS_WEIGHTS = {"loc_loss": 0.0, "ori_loss": 1.0}
    return cfg


def _relfro(a, b):
    return (a - b).norm().item() / max(b.norm().item(), 1e-30)


@pytest.mark.parametrize("classify", [True, False])
def test_gradient_accumulation_equals_full_batch(classify):
    from ursonet_b200.engine import Engine
    cfg = _cfg(classify)
    p64 = O.init_weights(cfg, seed=3, pretrained_like=True)
    img, gt_loc, gt_ori = make_batch(cfg, 4, seed=4)
    full = Engine(cfg, 4, training=True)
    load_oracle_weights(full, p64)
    full.img_u8.copy_(img); full.gt_loc.copy_(gt_loc); full.gt_ori.copy_(gt_ori)
    full.accumulate(0, 1, use_graph=False)
    torch.cuda.synchronize()
    g_full = full.grads.clone()
    for use_graph in (False, True):
        mic = Engine(cfg, 2, training=True)
        load_oracle_weights(mic, p64)
        for rep in range(2):              # the second pass runs through the captured graph and must overwrite, not add
            for m in range(2):
                sl = slice(2 * m, 2 * m + 2)
                mic.img_u8.copy_(img[sl]); mic.gt_loc.copy_(gt_loc[sl]); mic.gt_ori.copy_(gt_ori[sl])
                mic.accumulate(m, 2, use_graph=use_graph)
            torch.cuda.synchronize()
            assert _relfro(mic.grads, g_full) <= 2e-3, (use_graph, rep, _relfro(mic.grads, g_full))
    # and the update that follows is the one-batch update
    full.apply_update(1e-2, use_graph=False)
    mic.apply_update(1e-2, use_graph=True)
    torch.cuda.synchronize()
    assert _relfro(mic.params.flat, full.params.flat) <= 1e-5


def _rank_main(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from ursonet_b200.engine import Engine
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    cfg = _cfg(True)
    p64 = O.init_weights(cfg, seed=3, pretrained_like=True)
    img, gt_loc, gt_ori = make_batch(cfg, 2 * world, seed=4)
    eng = Engine(cfg, 2, training=True, world_size=world)
    load_oracle_weights(eng, p64)
    sl = slice(2 * rank, 2 * rank + 2)
    eng.img_u8.copy_(img[sl]); eng.gt_loc.copy_(gt_loc[sl]); eng.gt_ori.copy_(gt_ori[sl])
    for _ in range(2):                    # second step through the captured graphs
        eng.train_step(1e-2, allreduce=lambda g: dist.all_reduce(g), use_graph=True)
    torch.cuda.synchronize()
    q.put((rank, eng.params.flat.cpu()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_rank_step_equals_one_rank_on_concatenated_batch():
    import torch.multiprocessing as mp
    from ursonet_b200.engine import Engine
    world = 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, q)) for r in range(world)]
    for p_ in procs:
        p_.start()
    got = {}
    for _ in range(world):
        r, flat = q.get(timeout=600)
        got[r] = flat
    for p_ in procs:
        p_.join(120)
        assert p_.exitcode == 0
    assert torch.equal(got[0], got[1]), "every rank must apply the identical update"
    cfg = _cfg(True)
    p64 = O.init_weights(cfg, seed=3, pretrained_like=True)
    img, gt_loc, gt_ori = make_batch(cfg, 2 * world, seed=4)
    one = Engine(cfg, 2 * world, training=True)
    load_oracle_weights(one, p64)
    one.img_u8.copy_(img); one.gt_loc.copy_(gt_loc); one.gt_ori.copy_(gt_ori)
    w0 = one.params.flat.clone()
    for _ in range(2):
        one.train_step(1e-2, use_graph=True)
    torch.cuda.synchronize()
    d_one, d_two = one.params.flat.cpu() - w0.cpu(), got[0] - w0.cpu()
    # two steps through a chaotic random net: the second step amplifies the first step's atomic-order noise
    assert _relfro(d_two, d_one) <= 2e-2, _relfro(d_two, d_one)
