"""world_size-2 gloo test (CPU) of the data-parallel semantics: summing per-shard gradient arenas and scaling by
1/world reproduces the full-batch gradient for the per-sample-mean losses, shard_indices partitions an epoch, and the
update after the all-reduce is identical on both ranks."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ursonet_oracle as O
from ursonet_b200 import dp
from ursonet_b200.config import Config


def _cfg(regress_ori):
    c = Config()
    c.BACKBONE, c.BOTTLENECK_WIDTH, c.BRANCH_SIZE, c.NR_DENSE_LAYERS = "resnet18", 32, 32, 1
    c.ORI_BINS_PER_DIM, c.REGRESS_ORI = 4, regress_ori
    c.IMAGE_MIN_DIM, c.IMAGE_MAX_DIM = 64, 64
    c.LOSS_WEIGHTS = {"loc_loss": 0.0, "ori_loss": 1.0}     # isolate the per-sample-mean loss
    c.update()
    return c


def _batch(cfg, B):
    g = torch.Generator().manual_seed(7)
    img = torch.randint(0, 256, (B, 64, 64, 3), generator=g, dtype=torch.uint8)
    loc = torch.randn(B, 3, generator=g, dtype=torch.float64) + 5
    if cfg.REGRESS_ORI:
        ori = torch.nn.functional.normalize(torch.randn(B, 4, generator=g, dtype=torch.float64), dim=-1)
    else:
        ori = torch.softmax(torch.randn(B, 64, generator=g, dtype=torch.float64), -1)
    return O.mold_image(img), loc, ori


def _worker(rank, world, port, regress_ori, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    r, w = dp.init_distributed("gloo")
    assert (r, w) == (rank, world)
    cfg = _cfg(regress_ori)
    p = O.init_weights(cfg, seed=1)
    img, loc, ori = _batch(cfg, 4)
    sl = slice(rank * 2, rank * 2 + 2)
    grads, _, _ = O.gradients(p, (img[sl], loc[sl], ori[sl]), cfg)
    names = sorted(grads)
    flat = torch.cat([(grads[n] - 2 * cfg.WEIGHT_DECAY * p[n] / p[n].numel() * O.is_regularised(n)).reshape(-1) for n in names])
    dp.make_allreduce(world)(flat)            # SUM over ranks
    flat = flat / world                        # the 1/world of urso_add_reg_sumsq
    ms = dp.max_over_ranks_ms(float(rank + 1), device="cpu")
    q.put((rank, flat.numpy().tolist(), ms))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("regress_ori", [False, True])
def test_allreduced_shard_gradients_equal_full_batch(regress_ori):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, regress_ori, q)) for r in range(2)]
    for p_ in procs:
        p_.start()
    out = {}
    for _ in range(2):
        rank, flat, ms = q.get(timeout=240)
        out[rank] = (torch.tensor(flat, dtype=torch.float64), ms)
    for p_ in procs:
        p_.join(60)
        assert p_.exitcode == 0
    assert torch.equal(out[0][0], out[1][0])               # identical on both ranks -> identical update
    assert out[0][1] == 2.0 and out[1][1] == 2.0           # MAX over ranks
    cfg = _cfg(regress_ori)
    p = O.init_weights(cfg, seed=1)
    grads, _, _ = O.gradients(p, _batch(cfg, 4), cfg)
    names = sorted(grads)
    full = torch.cat([(grads[n] - 2 * cfg.WEIGHT_DECAY * p[n] / p[n].numel() * O.is_regularised(n)).reshape(-1) for n in names])
    assert torch.allclose(out[0][0], full, rtol=1e-9, atol=1e-12)


def test_rel_loss_is_tower_style_under_sharding():
    """rel_loss normalises by the shard's own ||gt|| (net.py:757): documented deviation from the full-batch loss."""
    gt = torch.tensor([[1.0, 0, 0], [0, 10.0, 0]], dtype=torch.float64)
    pr = gt + 0.5
    full = O.rel_loss(gt, pr)
    tower = (O.rel_loss(gt[:1], pr[:1]) + O.rel_loss(gt[1:], pr[1:])) / 2
    assert abs(full.item() - tower.item()) > 1e-2


def test_shard_indices_partition_an_epoch():
    a, b = dp.shard_indices(11, 0, 2, seed=3, epoch=5), dp.shard_indices(11, 1, 2, seed=3, epoch=5)
    assert len(a) == len(b) == 5 and not set(a) & set(b)
    assert dp.shard_indices(11, 0, 2, seed=3, epoch=6) != a
