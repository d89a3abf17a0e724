"""HBM ceilings for different read:write mixes (torch elementwise kernels, CUDA events, best of 5).
Calibrates what the write-heavy conv epilogues can hope for.  Run on the GPU box: python scripts/hbm_mix.py"""
import json
import torch

n = 1 << 30  # bf16 elements = 2 GiB per tensor
a = torch.empty(n, dtype=torch.bfloat16, device="cuda").normal_()
b = torch.empty_like(a).normal_()
c = torch.empty_like(a)


def t(fn, nbytes):
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return nbytes / best / 1e6   # GB/s


out = {
    "write_only_fill": t(lambda: c.fill_(1.0), 2 * n),
    "read_only_sum": t(lambda: a.sum(), 2 * n),
    "copy_1r1w": t(lambda: c.copy_(a), 4 * n),
    "add_2r1w": t(lambda: torch.add(a, b, out=c), 6 * n),
    "relu_inplace_1r1w": t(lambda: a.relu_(), 4 * n),
    "bcast_write_4w": t(lambda: c.view(4, -1).copy_(a[: n // 4].unsqueeze(0).expand(4, -1)), 2 * n + n // 2),
}
print(json.dumps(out, indent=1))
