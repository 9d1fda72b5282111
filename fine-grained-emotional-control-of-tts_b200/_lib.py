"""ctypes binding of libfs2_b200.so (the C ABI declared in include/fs2_b200.h).

There is no CPU fallback: if the shared library is missing or a call returns a
non-zero status, a RuntimeError is raised.  Tensors are passed as raw device
pointers; every call runs on torch's current CUDA stream.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# FS2_B200_LIB: measurement tools load the probe build (make -C csrc probe) instead; never a fallback -- a missing file raises
LIB_PATH = os.environ.get("FS2_B200_LIB") or os.path.join(_HERE, "libfs2_b200.so")

PAD = 4  # FS2_PAD: halo rows either side of every batch item in the padded row space

_lib = None

# argument codes of every entry point in include/fs2_b200.h that returns a status
# (p = pointer, i = int, f = float, q = long long, Q = unsigned long long); the trailing p is the stream
SIGNATURES = {
    "fs2_gemm_simt": "pp",
    "fs2_gemm_tc": "pp",
    "fs2_gemm_ln_tc": "pp",
    "fs2_embed_posenc": "pppiiiippipp",
    "fs2_embedding_bwd": "ppiiiipp",
    "fs2_ln_fwd": "pp",
    "fs2_ln_bwd": "pp",
    "fs2_softmax_fwd": "ppiiiiffQpppip",
    "fs2_softmax_bwd": "pppiiiiffQppip",
    "fs2_counter_add": "pQp",
    "fs2_cond_finish": "ppppppiiipppiip",
    "fs2_cond_bwd": "pppppiiipppp",
    "fs2_avg_over_durations": "ppiiippppp",
    "fs2_embed_add": "ppppipiiippiip",
    "fs2_embed_add_bwd": "ppiiiippp",
    "fs2_frames_to_rows": "piiiiipip",
    "fs2_gelu": "pqip",
    "fs2_prototype_buckets": "pppppiiiiipp",
    "fs2_collate": "pppppppppiiiipppppp" + "p",
    "fs2_intensity_head": "ppppppiiiipp",
    "fs2_flash_attn_fwd": "ppiiiiffQpppip",
    "fs2_flash_attn_bwd": "pppppiiiiffQpppip",
    "fs2_flash_attn_mask": "iifQppp",
    "fs2_dur_decode": "pqpp",
    "fs2_lr_prepare": "ppfiippp",
    "fs2_lr_finalize": "piippp",
    "fs2_lr_expand": "piipppiiiippiiipp",
    "fs2_lr_bwd": "ppiippiiiipiip",
    "fs2_fold_halo": "piiiipppppip",
    "fs2_colsum": "piqiqpp",
    "fs2_unpad_mask": "ppiiippiip",
    "fs2_pad_rows": "ppiiifppip",
    "fs2_pack_weights": "pippip",
    "fs2_cast_bf16": "ppqp",
    "fs2_add_": "ppqp",
    "fs2_memset": "piqp",
    "fs2_mse_losses": "pppppppppppiiiippppppppp",
    "fs2_ssim_loss": "pppiiifpppp",
    "fs2_loss_fused": "ppppppppppp" + "iiii" + "pppppppp" + "p",
    "fs2_loss_scale_grads": "pppqpppqp",
    "fs2_adamw": "ppppqfffffifp",
    "fs2_adamw_fused": "ppppqfffffifpqqqiiqip",
    "fs2_intensity_segment_mean": "pppiiiipp",
}


class Fs2Gemm(C.Structure):
    _fields_ = [
        ("mode", C.c_int), ("M", C.c_int), ("N", C.c_int), ("K", C.c_int), ("taps", C.c_int),
        ("batch1", C.c_int), ("batch2", C.c_int),
        ("A", C.c_void_p), ("lda", C.c_longlong), ("a_s1", C.c_longlong), ("a_s2", C.c_longlong),
        ("a_rows", C.c_int), ("a_inner", C.c_int), ("a_row_off", C.c_int), ("a_tap_step", C.c_int),
        ("B", C.c_void_p), ("ldb", C.c_longlong), ("b_s1", C.c_longlong), ("b_s2", C.c_longlong),
        ("b_rows", C.c_int), ("b_inner", C.c_int), ("b_row_off", C.c_int), ("b_tap_step", C.c_int),
        ("c_tap_stride", C.c_longlong), ("c_col_stride", C.c_longlong),
        ("C", C.c_void_p), ("c_bf16", C.c_int), ("ldc", C.c_longlong), ("c_s1", C.c_longlong),
        ("c_s2", C.c_longlong), ("c_row_off", C.c_int), ("c_col_off", C.c_int),
        ("accumulate", C.c_int), ("split_k", C.c_int),
        ("bias", C.c_void_p), ("alpha", C.c_float), ("relu", C.c_int),
        ("relu_aux", C.c_void_p), ("aux_bf16", C.c_int),
        ("rs_T", C.c_int), ("rs_Tp", C.c_int), ("lens", C.c_void_p), ("halo", C.c_int),
        ("ab_bf16", C.c_int), ("c_split_stride", C.c_longlong), ("a_colsum", C.c_void_p),
    ]


class Fs2LnFwd(C.Structure):
    _fields_ = [
        ("B", C.c_int), ("T", C.c_int), ("C", C.c_int),
        ("x", C.c_void_p), ("branch", C.c_void_p),
        ("drop_b_p", C.c_float), ("drop_b_seed", C.c_ulonglong),
        ("gamma", C.c_void_p), ("beta", C.c_void_p), ("eps", C.c_float),
        ("tanh_act", C.c_int),
        ("drop_a_p", C.c_float), ("drop_a_seed", C.c_ulonglong),
        ("lens", C.c_void_p), ("post_add", C.c_void_p),
        ("out_f32", C.c_void_p), ("out_act", C.c_void_p), ("act_bf16", C.c_int), ("halo", C.c_int),
        ("mean", C.c_void_p), ("rstd", C.c_void_p),
        ("head_w", C.c_void_p), ("head_b", C.c_void_p), ("head_out", C.c_void_p),
        ("head_scale", C.c_float),
        ("seed_dev", C.c_void_p),
    ]


class Fs2LnBwd(C.Structure):
    _fields_ = [
        ("B", C.c_int), ("T", C.c_int), ("C", C.c_int),
        ("dy", C.c_void_p), ("dy2", C.c_void_p), ("dy2_fold", C.c_int),
        ("dhead", C.c_void_p), ("head_w", C.c_void_p), ("head_scale", C.c_float),
        ("x", C.c_void_p), ("branch", C.c_void_p),
        ("drop_b_p", C.c_float), ("drop_b_seed", C.c_ulonglong),
        ("gamma", C.c_void_p), ("beta", C.c_void_p), ("eps", C.c_float), ("tanh_act", C.c_int),
        ("drop_a_p", C.c_float), ("drop_a_seed", C.c_ulonglong),
        ("lens", C.c_void_p), ("mean", C.c_void_p), ("rstd", C.c_void_p),
        ("relu_x", C.c_int),
        ("dx_f32", C.c_void_p), ("dact", C.c_void_p), ("act_bf16", C.c_int),
        ("dgamma", C.c_void_p), ("dbeta", C.c_void_p), ("dhead_w", C.c_void_p), ("dhead_b", C.c_void_p),
        ("seed_dev", C.c_void_p), ("dact_colsum", C.c_void_p), ("dy3", C.c_void_p), ("y", C.c_void_p),
    ]


class Fs2GemmLn(C.Structure):
    _fields_ = [
        ("B", C.c_int), ("T", C.c_int), ("K", C.c_int), ("lda", C.c_longlong), ("ldw", C.c_longlong),
        ("A", C.c_void_p), ("W", C.c_void_p), ("bias", C.c_void_p), ("x", C.c_void_p),
        ("drop_p", C.c_float), ("drop_seed", C.c_ulonglong), ("seed_dev", C.c_void_p),
        ("gamma", C.c_void_p), ("beta", C.c_void_p), ("eps", C.c_float),
        ("out_f32", C.c_void_p), ("out_act", C.c_void_p), ("halo", C.c_int),
        ("mean", C.c_void_p), ("rstd", C.c_void_p),
    ]


class Fs2PackItem(C.Structure):
    _fields_ = [("src_off", C.c_longlong), ("dst_off", C.c_longlong), ("src_ld", C.c_longlong),
                ("cout", C.c_int), ("cin", C.c_int), ("k", C.c_int), ("pad_", C.c_int)]


def load():
    """Load the C-ABI library (once).  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"fs2_b200: {LIB_PATH} is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C fine-grained-emotional-control-of-tts_b200/csrc`).  There is no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    lib.fs2_last_error.restype = C.c_char_p
    lib.fs2_launch_count.restype = C.c_longlong
    lib.fs2_ssim_ws_floats.restype = C.c_longlong
    lib.fs2_ssim_ws_floats.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.fs2_loss_ws_floats.restype = C.c_longlong
    lib.fs2_loss_ws_floats.argtypes = [C.c_int]
    codes = {"p": C.c_void_p, "i": C.c_int, "f": C.c_float, "q": C.c_longlong, "Q": C.c_ulonglong}
    for name, sig in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch: fail loudly
        fn.argtypes = [codes[c] for c in sig]
        fn.restype = C.c_int
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().fs2_last_error()
        raise RuntimeError(f"fs2_b200: {what} failed (status {rc}): {msg.decode() if msg else ''}")


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  Refuses non-CUDA tensors: no CPU path."""
    if t is None:
        return C.c_void_p(0)
    if not t.is_cuda:
        raise RuntimeError("fs2_b200: tensor is not on a CUDA device (there is no CPU fallback)")
    return C.c_void_p(t.data_ptr())


def launch_count():
    return int(load().fs2_launch_count())


def call(name, *args):
    """Call an entry point on torch's current stream.  Tensors become device pointers, None becomes NULL."""
    lib = load()
    conv = [a.data_ptr() if isinstance(a, torch.Tensor) else a for a in args]
    rc = getattr(lib, name)(*conv, torch.cuda.current_stream().cuda_stream)
    if rc != 0:
        check(rc, name)


# ---------------------------------------------------------------------------- GEMM
def gemm(*, mode, M, N, K, A, lda, a_rows, a_inner, B, ldb, b_rows, b_inner, Cout, ldc, c_bf16,
         ab_bf16, taps=1, batch1=1, batch2=1, a_s1=0, a_s2=0, a_row_off=0, a_tap_step=0,
         b_s1=0, b_s2=0, b_row_off=0, b_tap_step=0, c_tap_stride=0, c_col_stride=0, c_s1=0, c_s2=0, c_row_off=0,
         c_col_off=0, accumulate=0, split_k=1, bias=None, alpha=1.0, relu=0, relu_aux=None,
         aux_bf16=0, rs_T=0, rs_Tp=0, lens=None, halo=0, use_tc=None, A_off=0, B_off=0, C_off=0, c_split_stride=0,
         a_colsum=None):
    """Thin wrapper over fs2_gemm_tc / fs2_gemm_simt.  A_off/B_off/C_off are element offsets
    added to the base pointers (in the operand's own element size)."""
    g = Fs2Gemm()
    g.mode, g.M, g.N, g.K, g.taps, g.batch1, g.batch2 = mode, M, N, K, taps, batch1, batch2
    es_ab = 2 if ab_bf16 else 4
    g.A = A.data_ptr() + A_off * es_ab
    g.lda, g.a_s1, g.a_s2 = lda, a_s1, a_s2
    g.a_rows, g.a_inner, g.a_row_off, g.a_tap_step = a_rows, a_inner, a_row_off, a_tap_step
    g.B = B.data_ptr() + B_off * es_ab
    g.ldb, g.b_s1, g.b_s2 = ldb, b_s1, b_s2
    g.b_rows, g.b_inner, g.b_row_off, g.b_tap_step = b_rows, b_inner, b_row_off, b_tap_step
    g.c_tap_stride = c_tap_stride
    g.c_col_stride = c_col_stride
    g.C = Cout.data_ptr() + C_off * (2 if c_bf16 else 4)
    g.c_bf16, g.ldc, g.c_s1, g.c_s2 = int(c_bf16), ldc, c_s1, c_s2
    g.c_row_off, g.c_col_off = c_row_off, c_col_off
    g.accumulate, g.split_k = accumulate, split_k
    g.bias = bias.data_ptr() if bias is not None else None
    g.alpha, g.relu = alpha, relu
    g.relu_aux = relu_aux.data_ptr() if relu_aux is not None else None
    g.aux_bf16 = aux_bf16
    g.rs_T, g.rs_Tp = rs_T, rs_Tp
    g.lens = lens.data_ptr() if lens is not None else None
    g.halo, g.ab_bf16 = halo, int(ab_bf16)
    g.c_split_stride = c_split_stride
    g.a_colsum = a_colsum.data_ptr() if a_colsum is not None else None
    if use_tc is None:
        use_tc = bool(ab_bf16)
    fn = "fs2_gemm_tc" if use_tc else "fs2_gemm_simt"
    call(fn, C.addressof(g))


def gemm_tc_error_flag():
    return int(load().fs2_gemm_tc_error_flag())


def gemm_tc_tune(pair=2, cfg=-1):
    """Measurement hook: kernel / tile selection of fs2_gemm_tc (see include/fs2_b200.h)."""
    lib = load()
    lib.fs2_gemm_tc_tune.argtypes = [C.c_int, C.c_int]
    lib.fs2_gemm_tc_tune.restype = C.c_int
    check(lib.fs2_gemm_tc_tune(int(pair), int(cfg)), "fs2_gemm_tc_tune")


def gemm_tc_set_debug(buf):
    """buf: int64 CUDA tensor of >= 2 elements, or None to switch the clock probe off."""
    lib = load()
    lib.fs2_gemm_tc_set_debug.argtypes = [C.c_void_p]
    lib.fs2_gemm_tc_set_debug.restype = C.c_int
    check(lib.fs2_gemm_tc_set_debug(C.c_void_p(buf.data_ptr()) if buf is not None else None), "fs2_gemm_tc_set_debug")
