"""Micro-benchmark of one Engine-F launch (CUDA events, L2-cold by rotating through several buffer sets).
usage: python scripts/bench_layer.py M K N [addend] [mask] [relu] [block_n=..] [k3] [h=.. w=..]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ursonet_b200 import lib, convplan as P

def run(M, K, N, addend=False, mask=False, relu=True, block_n=0, colsum=False, reps=20, nsets=4, conv3=None, halo=False):
    dev = "cuda"
    plans = []
    for s in range(nsets):
        if conv3:   # 3x3 s1 conv on [B,h,w,K]
            B, h, w = conv3
            x = torch.randn(B, h, w, K, device=dev).to(torch.bfloat16)
            g = P.make_geom(3, 1, "same", K, N, h, w)
            segs, idx = P.fwd_segments(g)
            bmat = torch.randn(N, len(idx), device=dev).to(torch.bfloat16)
            out = torch.empty(B, h, w, N, dtype=torch.bfloat16, device=dev)
            ad = torch.randn(B, h, w, N, device=dev).to(torch.bfloat16) if addend else None
            mk = torch.randn(B, h, w, N, device=dev).to(torch.bfloat16) if mask else None
            tw, th = (8, 16) if halo else P.pick_patch(h, w, 128)
            cs = torch.zeros(N, device=dev) if colsum else None
            plans.append(lib.ConvGemm([x], bmat, segs, out, w, h, B, tw, th, addend=ad, mask=mk, relu=relu, colsum=cs,
                                      block_n=block_n, halo=halo))
            flops = 2.0 * B * h * w * N * 9 * K
            nbytes = 2.0 * B * h * w * (K + N * (1 + addend + mask))
        else:
            x = torch.randn(1, 1, M, K, device=dev).to(torch.bfloat16)
            bmat = torch.randn(N, K, device=dev).to(torch.bfloat16)
            out = torch.empty(1, 1, M, N, dtype=torch.bfloat16, device=dev)
            ad = torch.randn(1, 1, M, N, device=dev).to(torch.bfloat16) if addend else None
            mk = torch.randn(1, 1, M, N, device=dev).to(torch.bfloat16) if mask else None
            cs = torch.zeros(N, device=dev) if colsum else None
            plans.append(lib.ConvGemm([x], bmat, [(0, 0, 0, K // 64)], out, M, 1, 1, 128, 1, addend=ad, mask=mk, relu=relu,
                                      colsum=cs, block_n=block_n))
            flops = 2.0 * M * N * K
            nbytes = 2.0 * M * (K + N * (1 + addend + mask))
    for p in plans:
        p.launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        plans[i % nsets].launch()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return ms, flops / ms / 1e9, nbytes / ms / 1e6

if __name__ == "__main__":
    cases = [
        ("2c fwd  K64 N256 +add", dict(M=1228800, K=64, N=256, addend=True)),
        ("2c fwd  K64 N256 noadd", dict(M=1228800, K=64, N=256)),
        ("2c fwd  K64 N256 +add bn64", dict(M=1228800, K=64, N=256, addend=True, block_n=64)),
        ("2c fwd  K64 N256 noadd bn256", dict(M=1228800, K=64, N=256, block_n=256)),
        ("2c fwd  K64 N256 noadd bn64", dict(M=1228800, K=64, N=256, block_n=64)),
        ("2a fwd  K256 N64", dict(M=1228800, K=256, N=64)),
        ("blockout dgrad K320 N256 +add+mask+cs", dict(M=1228800, K=320, N=256, addend=True, mask=True, relu=False, colsum=True)),
        ("st3 2c K128 N512 +add", dict(M=307200, K=128, N=512, addend=True)),
        ("st4 2c K256 N1024 +add", dict(M=76800, K=256, N=1024, addend=True)),
        ("st4 2a K1024 N256", dict(M=76800, K=1024, N=256)),
        ("st4 2a K1024 N256 bn128", dict(M=76800, K=1024, N=256, block_n=128)),
        ("st5 2c K512 N2048 +add", dict(M=19200, K=512, N=2048, addend=True)),
        ("st2 3x3 64->64", dict(M=0, K=64, N=64, conv3=(32, 160, 240))),
        ("st2 3x3 64->64 halo", dict(M=0, K=64, N=64, conv3=(32, 160, 240), halo=True)),
        ("st2 3x3 64->64 halo mask cs", dict(M=0, K=64, N=64, conv3=(32, 160, 240), halo=True, mask=True, relu=False, colsum=True)),
        ("st3 3x3 128->128", dict(M=0, K=128, N=128, conv3=(32, 80, 120))),
        ("st3 3x3 128->128 halo", dict(M=0, K=128, N=128, conv3=(32, 80, 120), halo=True)),
        ("st4 3x3 256->256 halo", dict(M=0, K=256, N=256, conv3=(32, 40, 60), halo=True)),
        ("st4 3x3 256->256", dict(M=0, K=256, N=256, conv3=(32, 40, 60))),
        ("st4 3x3 256->256 bn128", dict(M=0, K=256, N=256, conv3=(32, 40, 60), block_n=128)),
        ("st5 3x3 512->512", dict(M=0, K=512, N=512, conv3=(32, 20, 30))),
    ]
    sel = sys.argv[1:] 
    for name, kw in cases:
        if sel and not any(s in name for s in sel):
            continue
        ms, tf, gb = run(**kw)
        print(f"{name:42s} {ms:7.3f} ms {tf:8.1f} TF/s {gb:8.1f} GB/s")
