"""Data-parallel plumbing: one process per GPU (torchrun), torch.distributed for rendezvous, ONE all-reduce of the
flat fp32 gradient arena per step.  The reference has no multi-GPU path (net.py:694-697 is a commented
stub; GPU_COUNT stays 1, pose_estimator.py:870).

Semantics: every rank runs forward/backward on its own shard of the global batch; gradients are summed by the
all-reduce and scaled by 1/world inside the fused regulariser+norm kernel (urso_add_reg_sumsq), the global-norm clip is
then computed on the AVERAGED gradient, so every rank applies the identical update (no broadcast needed).
Losses that are per-sample means (soft-label cross-entropy, 1-|q.q|) give exactly the full-batch gradient; rel_loss
normalises by the Frobenius norm of the shard's own gt_loc (net.py:757), i.e. tower-style semantics.
"""
import os

import torch
import torch.distributed as dist


def init_distributed(backend=None):
    """Initialise torch.distributed from the torchrun environment (no-op for a single process)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1 or dist.is_initialized():
        return int(os.environ.get("RANK", "0")), world
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group(backend)
    return dist.get_rank(), dist.get_world_size()


def make_allreduce(world):
    """Callable for Engine.train_step(allreduce=...): in-place SUM over ranks of the flat gradient arena."""
    if world <= 1:
        return None

    def allreduce(flat_grads):
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM)
    return allreduce


def shard_indices(n_items, rank, world, seed=0, epoch=0):
    """A rank's share of a shuffled epoch: identical permutation on all ranks, strided split, equal length
    (drops the remainder so that every rank runs the same number of steps)."""
    g = torch.Generator().manual_seed(seed * 100003 + epoch)
    perm = torch.randperm(n_items, generator=g).tolist()
    per = n_items // world
    return perm[rank:per * world:world]


def max_over_ranks_ms(ms, device="cuda"):
    """Device-timed duration reduced with MAX over ranks (how bench.py reports multi-GPU time)."""
    t = torch.tensor([ms], dtype=torch.float64, device=device if torch.cuda.is_available() else "cpu")
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()
