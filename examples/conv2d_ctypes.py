#!/usr/bin/env python
"""Complete, self-contained ctypes binding of ONE Conv2D of the reference through the C-ABI (include/urso_b200.h):

    x = KL.Conv2D(32, (3, 3), padding='SAME', strides=(2, 2), name='bottleneck_layer')(C5)          # net.py:639

No ursonet_b200 Python is imported: only the shared library, ctypes and torch (for device memory and the reference
result).  This is the stub INTEGRATION.md section B refers to; tests/test_gpu_cli.py runs it on the GPU.

    python examples/conv2d_ctypes.py [path/to/liburso_b200.so]
"""
import ctypes as C
import os
import sys

import torch

LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..",
                                                         "ursonet_b200", "liburso_b200.so")


class Conv2dShape(C.Structure):       # urso_conv2d_shape
    _fields_ = [(n, C.c_int32) for n in ("N", "H", "W", "C", "K", "ksize", "stride", "pad_t", "pad_l", "pad_b", "pad_r")]


class Conv2dFwdDesc(C.Structure):     # urso_conv2d_fwd_desc
    _fields_ = [("shape", Conv2dShape), ("x", C.c_void_p), ("w", C.c_void_p), ("scale", C.c_void_p),
                ("shift", C.c_void_p), ("addend", C.c_void_p), ("y", C.c_void_p), ("relu", C.c_int32),
                ("out_fp32", C.c_int32), ("workspace", C.c_void_p), ("relu_bits", C.c_void_p)]


def main():
    lib = C.CDLL(LIB)
    lib.urso_last_error.restype = C.c_char_p
    lib.urso_conv2d_fwd_workspace_bytes.restype = C.c_int64
    assert lib.urso_sizeof_conv2d_fwd_desc() == C.sizeof(Conv2dFwdDesc), "header / binding mismatch"

    def ok(rc):
        if rc != 0:
            raise RuntimeError(lib.urso_last_error().decode())

    dev = "cuda"
    N, H, W, Cin, K = 2, 20, 30, 2048, 32                      # C5 of a 640x960 frame -> [N, 10, 15, 32]
    g = torch.Generator().manual_seed(0)
    x = torch.randn(N, H, W, Cin, generator=g).to(torch.bfloat16).to(dev)           # NHWC bf16 activation
    w = (torch.randn(3, 3, Cin, K, generator=g) * 0.02).to(dev)                     # Keras HWIO fp32 kernel
    bias = torch.randn(K, generator=g).to(dev)

    # TF 'SAME' padding, computed by the library: an even map with stride 2 pads bottom / right only
    pt, pb, pl, pr = (C.c_int32() for _ in range(4))
    lib.urso_same_pad(H, 3, 2, C.byref(pt), C.byref(pb))
    lib.urso_same_pad(W, 3, 2, C.byref(pl), C.byref(pr))
    assert (pt.value, pb.value, pl.value, pr.value) == (0, 1, 0, 1)
    shape = Conv2dShape(N, H, W, Cin, K, 3, 2, pt.value, pl.value, pb.value, pr.value)
    OH, OW = (H + pt.value + pb.value - 3) // 2 + 1, (W + pl.value + pr.value - 3) // 2 + 1

    y = torch.empty(N, OH, OW, K, dtype=torch.float32, device=dev)                  # the bottleneck output stays fp32
    ws = torch.empty(lib.urso_conv2d_fwd_workspace_bytes(C.byref(shape)), dtype=torch.uint8, device=dev)
    d = Conv2dFwdDesc(shape, x.data_ptr(), w.data_ptr(), None, bias.data_ptr(), None, y.data_ptr(), 0, 1, ws.data_ptr(),
                      None)
    h = C.c_void_p()
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ok(lib.urso_conv2d_fwd_create(C.byref(d), C.byref(h)))      # plans segments / parity views / tiles, encodes tensor maps
    ok(lib.urso_conv2d_fwd_stage_weights(h, stream))            # fp32 HWIO master -> bf16 K-major GEMM operand
    ok(lib.urso_conv2d_fwd_launch(h, stream))                   # one tcgen05 implicit-GEMM launch, graph-capturable
    torch.cuda.synchronize()
    lib.urso_conv2d_fwd_destroy(h)

    # reference: the same conv in fp64 on the bf16-rounded operands (TF SAME = explicit pad bottom/right)
    xr = torch.nn.functional.pad(x.double().cpu().permute(0, 3, 1, 2), (0, 1, 0, 1))
    wr = w.to(torch.bfloat16).double().cpu().permute(3, 2, 0, 1)
    ref = torch.nn.functional.conv2d(xr, wr, bias.double().cpu(), stride=2).permute(0, 2, 3, 1)
    err = (y.double().cpu() - ref).abs().max().item() / ref.abs().max().item()
    print("conv2d 3x3/s2 SAME through the C-ABI: max rel err %.2e" % err)
    assert err < 1e-4
    return err


if __name__ == "__main__":
    main()
