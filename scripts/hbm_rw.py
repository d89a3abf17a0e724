"""HBM bandwidth by access mix (library kernels only, for the roofline denominators in DESIGN.md): pure write (fill),
pure read (sum), copy (read + write).  2 GiB buffers (>> 126 MB L2), CUDA events, best of 10."""
import torch

n = 1 << 30
a = torch.empty(n, dtype=torch.bfloat16, device="cuda")
b = torch.empty(n, dtype=torch.bfloat16, device="cuda")
a.fill_(1.0); b.fill_(2.0)


def best(fn, nbytes):
    ts = []
    for _ in range(10):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return nbytes / (min(ts) * 1e-3) / 1e9


print("write (fill_)   %.0f GB/s" % best(lambda: a.fill_(3.0), 2 * n))
print("read  (sum)     %.0f GB/s" % best(lambda: a.view(torch.int32).sum(), 2 * n))
print("copy  (copy_)   %.0f GB/s (read + write bytes)" % best(lambda: b.copy_(a), 4 * n))
