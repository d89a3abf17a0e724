"""ctypes binding of liburso_b200.so (the C-ABI declared in include/urso_b200.h).

There is NO fallback: if the shared library is missing or a call fails, an exception is raised.
The library is built in-tree by `make -C ursonet_b200/csrc` (see __graft_entry__.build()).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("URSO_LIB_PATH") or os.path.join(_HERE, "liburso_b200.so")   # override: A/B experiments

MAX_AMAPS = 8
MAX_SEGS = 32


class View4(C.Structure):
    _fields_ = [("base", C.c_void_p), ("C", C.c_int32), ("W", C.c_int32), ("H", C.c_int32), ("N", C.c_int32),
                ("stride_w", C.c_int64), ("stride_h", C.c_int64), ("stride_n", C.c_int64)]


class Seg(C.Structure):
    _fields_ = [("map_id", C.c_int32), ("dh", C.c_int32), ("dw", C.c_int32), ("c_chunks", C.c_int32)]


class Pix(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("sn", C.c_int64), ("sh", C.c_int64), ("sw", C.c_int64)]


class ConvGemmDesc(C.Structure):
    _fields_ = [("a", View4 * MAX_AMAPS), ("n_a", C.c_int32), ("b", C.c_void_p), ("b_rows", C.c_int32),
                ("b_k", C.c_int32), ("seg", Seg * MAX_SEGS), ("n_seg", C.c_int32),
                ("OW", C.c_int32), ("OH", C.c_int32), ("NB", C.c_int32), ("TW", C.c_int32), ("TH", C.c_int32),
                ("out", Pix), ("out_fp32", C.c_int32), ("shift", C.c_void_p), ("addend", Pix), ("mask", Pix),
                ("relu", C.c_int32), ("colsum", C.c_void_p), ("block_n", C.c_int32), ("relu_bits", Pix),
                ("mask_bits", Pix), ("halo", C.c_int32)]


class WgradDesc(C.Structure):
    _fields_ = [("p", View4 * MAX_AMAPS), ("n_p", C.c_int32), ("q", View4), ("seg", Seg * MAX_SEGS),
                ("n_seg", C.c_int32), ("PC", C.c_int32), ("QC", C.c_int32),
                ("OW", C.c_int32), ("OH", C.c_int32), ("NB", C.c_int32), ("TW", C.c_int32), ("TH", C.c_int32),
                ("g", C.c_void_p), ("g_seg_stride", C.c_int64), ("g_sp", C.c_int64), ("g_sq", C.c_int64),
                ("split_k", C.c_int32), ("block_q", C.c_int32)]


MAX_FANIN = 4


class Conv2dShape(C.Structure):
    _fields_ = [("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32), ("K", C.c_int32),
                ("ksize", C.c_int32), ("stride", C.c_int32), ("pad_t", C.c_int32), ("pad_l", C.c_int32),
                ("pad_b", C.c_int32), ("pad_r", C.c_int32)]


class Conv2dFwdDesc(C.Structure):
    _fields_ = [("shape", Conv2dShape), ("x", C.c_void_p), ("w", C.c_void_p), ("scale", C.c_void_p),
                ("shift", C.c_void_p), ("addend", C.c_void_p), ("y", C.c_void_p), ("relu", C.c_int32),
                ("out_fp32", C.c_int32), ("workspace", C.c_void_p), ("relu_bits", C.c_void_p)]


class Conv2dDgradDesc(C.Structure):
    _fields_ = [("n_convs", C.c_int32), ("shape", Conv2dShape * MAX_FANIN), ("dy", C.c_void_p * MAX_FANIN),
                ("w", C.c_void_p * MAX_FANIN), ("scale", C.c_void_p * MAX_FANIN), ("dy_sparse", C.c_int32),
                ("mask", C.c_void_p), ("addend", C.c_void_p), ("dx", C.c_void_p), ("colsum", C.c_void_p),
                ("workspace", C.c_void_p), ("mask_bits", C.c_void_p)]


class Conv2dWgradDesc(C.Structure):
    _fields_ = [("shape", Conv2dShape), ("x", C.c_void_p), ("dy", C.c_void_p), ("dy_sparse", C.c_int32),
                ("G", C.c_void_p)]


class BnJob(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("gamma", "beta", "mean", "var", "bias", "scale", "shift")] + \
               [("C", C.c_int32), ("pad_", C.c_int32)]


class StageJob(C.Structure):
    _fields_ = [("w", C.c_void_p), ("scale", C.c_void_p), ("out", C.c_void_p), ("index", C.c_void_p),
                ("kind", C.c_int32), ("K", C.c_int32), ("CI", C.c_int32), ("CO", C.c_int32), ("COp", C.c_int32),
                ("rows_out", C.c_int32), ("ld_out", C.c_int64), ("part", C.c_int32), ("block_begin", C.c_int32)]


class PgradJob(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("G", "g_row_map", "w", "colsum", "scale", "gamma", "mean", "var", "bias",
                                          "dW", "dbias", "dgamma", "dbeta", "S")] + \
               [(n, C.c_int32) for n in ("R", "CO", "rows_per_slab", "cblocks", "block_begin", "pad_")]


_i32, _i64, _f32, _vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p

# name -> argtypes (all return int unless listed in _RESTYPES); kept in sync with include/urso_b200.h
SIGNATURES = {
    "urso_version": [],
    "urso_last_error": [],
    "urso_num_sms": [],
    "urso_set_max_ctas": [_i32],
    "urso_set_dry_run": [_i32],
    "urso_set_pdl": [_i32],
    "urso_set_residual_mma": [_i32],
    "urso_set_wgrad_halo": [_i32],
    "urso_set_tail_split": [_i32],
    "urso_convgemm_tail_split": [_vp],
    "urso_conv2d_fwd_tail_split": [_vp],
    "urso_conv2d_dgrad_tail_split": [_vp, _i32],
    "urso_sizeof_convgemm_desc": [],
    "urso_sizeof_wgrad_desc": [],
    "urso_convgemm_create": [C.POINTER(ConvGemmDesc), C.POINTER(_vp)],
    "urso_convgemm_launch": [_vp, _vp],
    "urso_convgemm_destroy": [_vp],
    "urso_convgemm_plan_info": [_vp, C.POINTER(_i32)],
    "urso_wgrad_create": [C.POINTER(WgradDesc), C.POINTER(_vp)],
    "urso_wgrad_launch": [_vp, _vp],
    "urso_wgrad_destroy": [_vp],
    "urso_sizeof_conv2d_fwd_desc": [],
    "urso_sizeof_conv2d_dgrad_desc": [],
    "urso_sizeof_conv2d_wgrad_desc": [],
    "urso_same_pad": [_i32, _i32, _i32, C.POINTER(_i32), C.POINTER(_i32)],
    "urso_conv2d_fwd_workspace_bytes": [C.POINTER(Conv2dShape)],
    "urso_conv2d_fwd_create": [C.POINTER(Conv2dFwdDesc), C.POINTER(_vp)],
    "urso_conv2d_fwd_stage_weights": [_vp, _vp],
    "urso_conv2d_fwd_launch": [_vp, _vp],
    "urso_conv2d_fwd_destroy": [_vp],
    "urso_conv2d_fwd_plan_info": [_vp, C.POINTER(_i32)],
    "urso_conv2d_dgrad_workspace_bytes": [C.POINTER(Conv2dDgradDesc)],
    "urso_conv2d_dgrad_create": [C.POINTER(Conv2dDgradDesc), C.POINTER(_vp)],
    "urso_conv2d_dgrad_stage_weights": [_vp, _vp],
    "urso_conv2d_dgrad_launch": [_vp, _vp],
    "urso_conv2d_dgrad_untouched_phases": [_vp],
    "urso_conv2d_dgrad_num_launches": [_vp],
    "urso_conv2d_dgrad_plan_info": [_vp, _i32, C.POINTER(_i32)],
    "urso_conv2d_dgrad_destroy": [_vp],
    "urso_conv2d_wgrad_create": [C.POINTER(Conv2dWgradDesc), C.POINTER(_vp)],
    "urso_conv2d_wgrad_launch": [_vp, _vp],
    "urso_conv2d_wgrad_plan_info": [_vp, C.POINTER(_i32)],
    "urso_wgrad_plan_info": [_vp, C.POINTER(_i32)],
    "urso_conv2d_wgrad_destroy": [_vp],
    "urso_stem_grad_row_map": [C.POINTER(_i32)],
    "urso_stem_stage": [_vp, _i32, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _vp],
    "urso_maxpool_fwd": [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp],
    "urso_maxpool_bwd": [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp],
    "urso_dense_fwd": [_vp, _vp, _vp, _i32, _i32, _i32, _vp],
    "urso_dense_bias_act": [_vp, _vp, _i32, _i32, _i32, _vp],
    "urso_dense_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp],
    "urso_softmax_xent": [_vp, _vp, _vp, _vp, _i32, _i32, _f32, _vp],
    "urso_rel_loss": [_vp, _vp, _vp, _vp, _i32, _i32, _f32, _vp],
    "urso_quat_head": [_vp, _vp, _vp, _vp, _vp, _i32, _f32, _vp],
    "urso_bn_fold": [_vp, _vp, _vp, _vp, _vp, _f32, _vp, _vp, _i32, _vp],
    "urso_stage_weight_rows": [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i64, _i32, _vp],
    "urso_stage_weight_cols": [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i64, _vp],
    "urso_conv_param_grads": [_vp] * 9 + [_f32] + [_vp] * 5 + [_i32, _i32, _vp],
    "urso_sizeof_bn_job": [],
    "urso_sizeof_stage_job": [],
    "urso_sizeof_pgrad_job": [],
    "urso_bn_fold_multi": [_vp, _i32, _i32, _f32, _vp],
    "urso_stage_jobs_finalize": [_vp, _i32, _vp],
    "urso_stage_weights_multi": [_vp, _vp, _i32, _i32, _vp],
    "urso_pgrad_jobs_finalize": [_vp, _i32, _vp],
    "urso_conv_param_grads_multi": [_vp, _vp, _i32, _i32, _i32, _f32, _vp],
    "urso_conv2d_fwd_stage_job": [_vp, _vp],
    "urso_conv2d_dgrad_stage_jobs": [_vp, _vp, _i32],
    "urso_grad_accumulate": [_vp, _vp, _vp, _f32, _f32, _i64, _vp],
    "urso_add_reg_sumsq": [_vp, _vp, _vp, _vp, _f32, _vp, _i64, _vp],
    "urso_sgd_step": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp],
    "urso_amsgrad_step": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp],
    "urso_split_f32": [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp],
    "urso_maxpool_fwd_f32": [_vp, _vp, _i32, _i32, _i32, _i32, _vp],
    "urso_encode_ori": [_vp, _vp, _vp, _vp, _i32, _i32, _f32, _vp],
    "urso_decode_ori_moments": [_vp, _vp, _vp, _i32, _i32, _vp],
    "urso_sizeof_aug_params": [],
    "urso_sim2real_aug": [_vp, _vp, _vp, _i32, _i32, _i32, _vp],
    "urso_cast_f32_to_bf16": [_vp, _vp, _i64, _vp],
    "urso_cast_bf16_to_f32": [_vp, _vp, _i64, _vp],
    "urso_pad_cast_rows": [_vp, _vp, _vp, _i64, _i32, _i32, _vp],
    "urso_colsum_bf16": [_vp, _vp, _i64, _i32, _vp],
}
_RESTYPES = {"urso_last_error": C.c_char_p, "urso_convgemm_destroy": None, "urso_wgrad_destroy": None,
             "urso_same_pad": None, "urso_set_max_ctas": None, "urso_set_dry_run": None, "urso_set_pdl": None, "urso_set_residual_mma": None, "urso_set_wgrad_halo": None, "urso_set_tail_split": None, "urso_stem_grad_row_map": None, "urso_conv2d_fwd_destroy": None,
             "urso_conv2d_dgrad_destroy": None, "urso_conv2d_wgrad_destroy": None,
             "urso_conv2d_fwd_workspace_bytes": C.c_int64, "urso_conv2d_dgrad_workspace_bytes": C.c_int64}

_lib = None


class UrsoError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built -- there is no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise UrsoError(f"{LIB_PATH} not found: build it with `make -C {os.path.join(_HERE, 'csrc')}` "
                        "(or __graft_entry__.build()); this package has no CPU / eager fallback")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the header and the library diverge
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    if lib.urso_sizeof_convgemm_desc() != C.sizeof(ConvGemmDesc) or lib.urso_sizeof_wgrad_desc() != C.sizeof(WgradDesc):
        raise UrsoError("ctypes struct layout does not match include/urso_b200.h (rebuild the library)")
    if (lib.urso_sizeof_conv2d_fwd_desc() != C.sizeof(Conv2dFwdDesc)
            or lib.urso_sizeof_conv2d_dgrad_desc() != C.sizeof(Conv2dDgradDesc)
            or lib.urso_sizeof_conv2d_wgrad_desc() != C.sizeof(Conv2dWgradDesc)):
        raise UrsoError("ctypes conv2d operator structs do not match include/urso_b200.h (rebuild the library)")
    if (lib.urso_sizeof_bn_job() != C.sizeof(BnJob) or lib.urso_sizeof_stage_job() != C.sizeof(StageJob)
            or lib.urso_sizeof_pgrad_job() != C.sizeof(PgradJob)):
        raise UrsoError("ctypes job-table structs do not match include/urso_b200.h (rebuild the library)")
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().urso_last_error()
        raise UrsoError(f"{what} failed (rc={rc}): {msg.decode() if msg else '?'}")


def call(name, *args):
    """Call an int-returning entry point and raise on error."""
    check(getattr(load(), name)(*args), name)


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


def view4(t):
    """urso_view4 of a bf16 NHWC torch tensor/view [N,H,W,C] (channel stride must be 1)."""
    import torch
    assert t.dtype == torch.bfloat16 and t.dim() == 4 and t.stride(3) == 1, (t.dtype, t.shape, t.stride())
    n, h, w, c = t.shape
    return View4(t.data_ptr(), c, w, h, n, t.stride(2), t.stride(1), t.stride(0))


def pix(t):
    """urso_pix of an NHWC tensor/view (any dtype)."""
    if t is None:
        return Pix(None, 0, 0, 0)
    assert t.dim() == 4 and t.stride(3) == 1
    return Pix(t.data_ptr(), t.stride(0), t.stride(1), t.stride(2))


def stem_view(s):
    """The overlapping operand view E[N, H/2+3, W/2, 64] (pixel stride 16 elements) over the compact staged stem tensor
    S[N, H/2+3, W/2+3, 16] of urso_stem_stage -- what the ksize-7 operators build internally (see include/urso_b200.h)."""
    n, h2, w2, c = s.shape
    assert c == 16 and s.is_contiguous()
    return s.as_strided((n, h2, w2 - 3, 64), (h2 * w2 * 16, w2 * 16, 16, 1))


def bits_pix(t):
    """urso_pix (BYTE strides) of a bit-packed mask tensor/view: int32 [N,H,W,C/32]."""
    import torch
    if t is None:
        return Pix(None, 0, 0, 0)
    assert t.dtype == torch.int32 and t.dim() == 4 and t.stride(3) == 1
    return Pix(t.data_ptr(), 4 * t.stride(0), 4 * t.stride(1), 4 * t.stride(2))


class ConvGemm:
    """Owning wrapper of a urso_convgemm_t plan (tensor maps are encoded once; launch is graph-capturable)."""

    def __init__(self, a_views, b, segs, out, OW, OH, NB, TW, TH, shift=None, addend=None, mask=None, relu=False,
                 colsum=None, block_n=0, halo=False, relu_bits=None, mask_bits=None):
        import torch
        d = ConvGemmDesc()
        assert 1 <= len(a_views) <= MAX_AMAPS and 1 <= len(segs) <= MAX_SEGS
        for i, v in enumerate(a_views):
            d.a[i] = view4(v)
        d.n_a = len(a_views)
        assert b.dtype == torch.bfloat16 and b.dim() == 2 and b.is_contiguous()
        d.b, d.b_rows, d.b_k = b.data_ptr(), b.shape[0], b.shape[1]
        for i, (m, dh, dw, ch) in enumerate(segs):
            d.seg[i] = Seg(m, dh, dw, ch)
        d.n_seg = len(segs)
        d.OW, d.OH, d.NB, d.TW, d.TH = OW, OH, NB, TW, TH
        d.out = pix(out)
        d.out_fp32 = 1 if out.dtype == torch.float32 else 0
        d.shift = ptr(shift)
        d.addend, d.mask = pix(addend), pix(mask)
        d.relu = int(relu)
        d.colsum = ptr(colsum)
        d.block_n = block_n
        d.halo = int(halo)
        d.relu_bits, d.mask_bits = bits_pix(relu_bits), bits_pix(mask_bits)
        self._keep = (a_views, b, out, shift, addend, mask, colsum, relu_bits, mask_bits)   # keep tensors alive
        h = _vp()
        check(load().urso_convgemm_create(C.byref(d), C.byref(h)), "urso_convgemm_create")
        self._h = h

    def launch(self):
        check(load().urso_convgemm_launch(self._h, stream_ptr()), "urso_convgemm_launch")

    def plan_info(self):
        """dict(block_n, npipe, stages, kpack, halo, bres, a_stages, smem_bytes, grid) of the planned launch."""
        v = (_i32 * 9)()
        check(load().urso_convgemm_plan_info(self._h, v), "urso_convgemm_plan_info")
        return dict(zip(("block_n", "npipe", "stages", "kpack", "halo", "bres", "a_stages", "smem_bytes", "grid"), v))

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.urso_convgemm_destroy(self._h)
            self._h = None


class Wgrad:
    """Owning wrapper of a urso_wgrad_t plan."""

    def __init__(self, p_views, q_view, segs, PC, QC, OW, OH, NB, TW, TH, g, g_seg_stride, g_sp, g_sq, split_k=0,
                 block_q=0):
        import torch
        d = WgradDesc()
        for i, v in enumerate(p_views):
            d.p[i] = view4(v)
        d.n_p = len(p_views)
        d.q = view4(q_view)
        for i, (m, dh, dw) in enumerate(segs):
            d.seg[i] = Seg(m, dh, dw, 0)
        d.n_seg = len(segs)
        d.PC, d.QC, d.OW, d.OH, d.NB, d.TW, d.TH = PC, QC, OW, OH, NB, TW, TH
        assert g.dtype == torch.float32
        d.g, d.g_seg_stride, d.g_sp, d.g_sq = g.data_ptr(), g_seg_stride, g_sp, g_sq
        d.split_k, d.block_q = split_k, block_q
        self._keep = (p_views, q_view, g)
        h = _vp()
        check(load().urso_wgrad_create(C.byref(d), C.byref(h)), "urso_wgrad_create")
        self._h = h

    def launch(self):
        check(load().urso_wgrad_launch(self._h, stream_ptr()), "urso_wgrad_launch")

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.urso_wgrad_destroy(self._h)
            self._h = None


# ------------------------------------------------------------------------------------------------ Conv2D operators
def same_pad(n, k, s):
    """TF 'SAME' (before, after) padding of one dimension, computed by the library (urso_same_pad)."""
    b, a = _i32(), _i32()
    load().urso_same_pad(n, k, s, C.byref(b), C.byref(a))
    return b.value, a.value


def conv_shape(N, H, W, Cin, K, ksize, stride, padding):
    """urso_conv2d_shape from a Keras-style padding spec: 'same' | 'valid' | int (explicit symmetric ZeroPadding2D)."""
    if padding == "same":
        (pt, pb), (pl, pr) = same_pad(H, ksize, stride), same_pad(W, ksize, stride)
    elif padding == "valid":
        pt = pb = pl = pr = 0
    else:
        pt = pb = pl = pr = int(padding)
    return Conv2dShape(N, H, W, Cin, K, ksize, stride, pt, pl, pb, pr)


def out_hw(s):
    return (s.H + s.pad_t + s.pad_b - s.ksize) // s.stride + 1, (s.W + s.pad_l + s.pad_r - s.ksize) // s.stride + 1


def _workspace(nbytes, device):
    import torch
    if nbytes < 0:
        check(2, "workspace query")
    return torch.zeros(max(int(nbytes), 256), dtype=torch.uint8, device=device)


class Conv2dFwd:
    """urso_conv2d_fwd_t: y = relu?( conv(x, w * scale) + shift + addend ).  All planning happens in the library."""

    def __init__(self, shape, x, w, scale, shift, y, addend=None, relu=False, relu_bits=None):
        import torch
        d = Conv2dFwdDesc()
        d.relu_bits = ptr(relu_bits)
        assert relu_bits is None or (relu_bits.dtype == torch.int32 and relu_bits.is_contiguous())
        d.shape = shape
        self.ws = _workspace(load().urso_conv2d_fwd_workspace_bytes(C.byref(shape)), x.device)
        d.x, d.w, d.scale, d.shift, d.addend, d.y = ptr(x), ptr(w), ptr(scale), ptr(shift), ptr(addend), ptr(y)
        d.relu, d.out_fp32, d.workspace = int(relu), int(y.dtype == torch.float32), self.ws.data_ptr()
        assert x.is_contiguous() and y.is_contiguous() and (addend is None or addend.is_contiguous())
        self._keep = (x, w, scale, shift, y, addend, relu_bits)
        h = _vp()
        check(load().urso_conv2d_fwd_create(C.byref(d), C.byref(h)), "urso_conv2d_fwd_create")
        self._h = h

    def stage(self):
        check(load().urso_conv2d_fwd_stage_weights(self._h, stream_ptr()), "urso_conv2d_fwd_stage_weights")

    def plan_info(self):
        """Planned Engine-F launch: dict(block_n, npipe, stages, kpack, halo, bres, a_stages, smem_bytes, grid)."""
        v = (_i32 * 9)()
        check(load().urso_conv2d_fwd_plan_info(self._h, v), "urso_conv2d_fwd_plan_info")
        d = dict(zip(("block_n", "npipe", "stages", "kpack", "halo", "bres", "a_stages", "smem_bytes", "grid"), v))
        d["tail_split"] = int(load().urso_conv2d_fwd_tail_split(self._h))
        return d

    def launch(self):
        check(load().urso_conv2d_fwd_launch(self._h, stream_ptr()), "urso_conv2d_fwd_launch")

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.urso_conv2d_fwd_destroy(self._h)
            self._h = None


class Conv2dDgrad:
    """urso_conv2d_dgrad_t: dx = mask( sum_i dgrad_i(dy_i, w_i * scale_i) + addend ) with fused fan-in."""

    def __init__(self, shapes, dys, ws, scales, dx, mask=None, addend=None, colsum=None, dy_sparse=False,
                 mask_bits=None):
        d = Conv2dDgradDesc()
        d.mask_bits = ptr(mask_bits)
        assert mask_bits is None or (mask is None and mask_bits.is_contiguous())
        d.n_convs = len(shapes)
        for i, (s, dy, w, sc) in enumerate(zip(shapes, dys, ws, scales)):
            assert dy.is_contiguous()
            d.shape[i], d.dy[i], d.w[i], d.scale[i] = s, ptr(dy), ptr(w), ptr(sc)
        d.dy_sparse = int(dy_sparse)
        d.mask, d.addend, d.dx, d.colsum = ptr(mask), ptr(addend), ptr(dx), ptr(colsum)
        assert dx.is_contiguous() and (mask is None or mask.is_contiguous()) and (addend is None or addend.is_contiguous())
        self.ws = _workspace(load().urso_conv2d_dgrad_workspace_bytes(C.byref(d)), dx.device)
        d.workspace = self.ws.data_ptr()
        self._keep = (dys, ws, scales, dx, mask, addend, colsum, mask_bits)
        h = _vp()
        check(load().urso_conv2d_dgrad_create(C.byref(d), C.byref(h)), "urso_conv2d_dgrad_create")
        self._h = h
        self.n_launches = load().urso_conv2d_dgrad_num_launches(h)
        self.untouched = load().urso_conv2d_dgrad_untouched_phases(h)

    def stage(self):
        check(load().urso_conv2d_dgrad_stage_weights(self._h, stream_ptr()), "urso_conv2d_dgrad_stage_weights")

    def launch(self):
        check(load().urso_conv2d_dgrad_launch(self._h, stream_ptr()), "urso_conv2d_dgrad_launch")

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.urso_conv2d_dgrad_destroy(self._h)
            self._h = None


class Conv2dWgrad:
    """urso_conv2d_wgrad_t: G[k*k*C, K] += x (*) dy  (raw fp32 weight gradient of the unscaled conv; caller zeroes G)."""

    def __init__(self, shape, x, dy, G, dy_sparse=False):
        d = Conv2dWgradDesc()
        d.shape, d.x, d.dy, d.dy_sparse, d.G = shape, ptr(x), ptr(dy), int(dy_sparse), ptr(G)
        assert x.is_contiguous() and dy.is_contiguous()
        self._keep = (x, dy, G)
        h = _vp()
        check(load().urso_conv2d_wgrad_create(C.byref(d), C.byref(h)), "urso_conv2d_wgrad_create")
        self._h = h

    def launch(self):
        check(load().urso_conv2d_wgrad_launch(self._h, stream_ptr()), "urso_conv2d_wgrad_launch")

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.urso_conv2d_wgrad_destroy(self._h)
            self._h = None


def stem_grad_row_map():
    m = (_i32 * 147)()
    load().urso_stem_grad_row_map(m)
    return list(m)


# ------------------------------------------------------------------------------------------------ multi-tensor job tables
def _table_to_device(arr, device):
    """ctypes array of structs -> uint8 device tensor holding the same bytes."""
    import torch
    if C.sizeof(arr) == 0:
        return torch.empty(0, dtype=torch.uint8, device=device)
    buf = (C.c_char * C.sizeof(arr)).from_buffer(arr)
    return torch.frombuffer(bytearray(buf), dtype=torch.uint8).to(device)


class BnFoldTable:
    """urso_bn_fold_multi over a list of (gamma, beta, mean, var, bias, scale, shift, C) tensor tuples (None allowed)."""

    def __init__(self, jobs, device):
        arr = (BnJob * len(jobs))()
        for i, (gm, bt, mu, vr, bias, sc, sh, c) in enumerate(jobs):
            arr[i] = BnJob(ptr(gm), ptr(bt), ptr(mu), ptr(vr), ptr(bias), ptr(sc), ptr(sh), int(c), 0)
        self.n, self.max_c = len(jobs), max(int(j[7]) for j in jobs)
        self.dev = _table_to_device(arr, device)
        self._keep = jobs

    def launch(self, eps):
        call("urso_bn_fold_multi", self.dev.data_ptr(), self.n, self.max_c, eps, stream_ptr())


class StageTable:
    """urso_stage_weights_multi over the staging jobs of a list of Conv2dFwd / Conv2dDgrad operators."""

    def __init__(self, fwd_ops, dgrad_ops, device):
        jobs = []
        for op in fwd_ops:
            j = StageJob()
            check(load().urso_conv2d_fwd_stage_job(op._h, C.byref(j)), "urso_conv2d_fwd_stage_job")
            jobs.append(j)
        for op in dgrad_ops:
            tmp = (StageJob * 64)()
            n = load().urso_conv2d_dgrad_stage_jobs(op._h, tmp, 64)
            if n < 0 or n > 64:
                raise UrsoError("urso_conv2d_dgrad_stage_jobs failed")
            jobs.extend(tmp[i] for i in range(n))
        arr = (StageJob * len(jobs))()
        for i, j in enumerate(jobs):
            arr[i] = j
        begins = (_i32 * len(jobs))()
        self.total = load().urso_stage_jobs_finalize(arr, len(jobs), begins)
        self.n = len(jobs)
        self.dev = _table_to_device(arr, device)
        self.begins = _table_to_device(begins, device)
        self._keep = (fwd_ops, dgrad_ops)

    def launch(self):
        if self.n == 0:
            return
        call("urso_stage_weights_multi", self.dev.data_ptr(), self.begins.data_ptr(), self.n, self.total, stream_ptr())


class PgradTable:
    """urso_conv_param_grads_multi over a list of dicts with the tensors of urso_conv_param_grads."""

    def __init__(self, jobs, device):
        arr = (PgradJob * len(jobs))()
        for i, j in enumerate(jobs):
            arr[i] = PgradJob(ptr(j["G"]), ptr(j.get("row_map")), ptr(j["w"]), ptr(j.get("colsum")), ptr(j.get("scale")),
                              ptr(j.get("gamma")), ptr(j.get("mean")), ptr(j.get("var")), ptr(j.get("bias")),
                              ptr(j["dW"]), ptr(j.get("dbias")), ptr(j.get("dgamma")), ptr(j.get("dbeta")), ptr(j.get("S")),
                              int(j["R"]), int(j["CO"]), 0, 0, 0, 0)
        begins = (_i32 * len(jobs))()
        self.total = load().urso_pgrad_jobs_finalize(arr, len(jobs), begins)
        self.n, self.max_co = len(jobs), max(int(j["CO"]) for j in jobs)
        self.dev = _table_to_device(arr, device)
        self.begins = _table_to_device(begins, device)
        self._keep = jobs

    def launch(self, eps):
        call("urso_conv_param_grads_multi", self.dev.data_ptr(), self.begins.data_ptr(), self.n, self.total, self.max_co,
             eps, stream_ptr())
