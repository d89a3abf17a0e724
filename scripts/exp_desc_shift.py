"""Experiment: does a UMMA K-major SWIZZLE_128B descriptor whose start address is shifted by whole 128-byte rows
(not a multiple of the 1024-byte swizzle atom) read the rows TMA wrote?  out_shifted[r] should equal out[r + shift]."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ursonet_b200 import lib

def run(shift, base_off):
    os.environ["URSO_DBG_ROW_SHIFT"] = str(shift)
    os.environ["URSO_DBG_BASE_OFFSET"] = str(base_off)
    M, K, N = 128, 64, 64
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 1, M, K, generator=g).to(torch.bfloat16).cuda()
    b = torch.randn(N, K, generator=g).to(torch.bfloat16).cuda()
    out = torch.zeros(1, 1, M, N, dtype=torch.float32, device="cuda")
    lib.ConvGemm([x], b, [(0, 0, 0, 1)], out, M, 1, 1, 128, 1).launch()
    torch.cuda.synchronize()
    return out[0, 0].double().cpu(), (x[0, 0].double() @ b.double().T).cpu()

ref0, ref = run(0, 0)
print("unshifted max err", (ref0 - ref).abs().max().item())
for shift in (1, 2, 3, 5, 8):
    for bo in (0, 1):
        got, _ = run(shift, bo)
        n = 128 - shift
        err = (got[:n] - ref[shift:]).abs().max().item()
        # alternative hypothesis: rows within each 8-row group permuted / wrong swizzle -> large error
        print(f"shift {shift} base_offset_field {bo}: max err vs ref[r+shift] = {err:.4f}   (scale {ref.abs().max().item():.1f})")
