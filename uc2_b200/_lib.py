"""ctypes binding of libuc2_b200.so (the C ABI declared in include/uc2_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails, a
RuntimeError is raised with the library's own message.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libuc2_b200.so")
_lib = None

ACT_NONE, ACT_GELU, ACT_DGELU, ACT_TANH = 0, 1, 2, 3


class GemmArgs(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("lda", C.c_longlong), ("a_mn", C.c_int),
        ("b", C.c_void_p), ("ldb", C.c_longlong), ("b_mn", C.c_int),
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("bias", C.c_void_p),
        ("residual", C.c_void_p), ("ld_res", C.c_longlong),
        ("aux", C.c_void_p), ("ld_aux", C.c_longlong),
        ("act", C.c_int),
        ("out_bf16", C.c_void_p), ("ld_out", C.c_longlong),
        ("out_pre", C.c_void_p), ("ld_pre", C.c_longlong),
        ("out_f32", C.c_void_p), ("ld_f32", C.c_longlong),
        ("accumulate", C.c_int), ("split_k", C.c_int), ("block_n", C.c_int),
    ]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"uc2_b200: {LIB_PATH} not found. Build it with `python -m uc2_b200.build` "
                "(there is no CPU / PyTorch fallback for this path).")
        _lib = C.CDLL(LIB_PATH)
        _lib.uc2_last_error.restype = C.c_char_p
        _lib.uc2_launch_count.restype = C.c_longlong
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().uc2_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"uc2_b200 {what} failed (code {rc}): {msg}")


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def launch_count():
    return int(lib().uc2_launch_count())


def gemm(a, b, M, N, K, *, a_mn=False, b_mn=False, lda=None, ldb=None, bias=None, residual=None, aux=None,
         act=ACT_NONE, out_bf16=None, out_pre=None, out_f32=None, accumulate=False, split_k=1, block_n=0):
    """D[M,N] = A[M,K] @ B[N,K]^T with the fused epilogue described in include/uc2_b200.h."""
    g = GemmArgs()
    g.a, g.lda, g.a_mn = a.data_ptr(), (lda if lda is not None else a.stride(0)), int(a_mn)
    g.b, g.ldb, g.b_mn = b.data_ptr(), (ldb if ldb is not None else b.stride(0)), int(b_mn)
    g.M, g.N, g.K = M, N, K
    g.bias = bias.data_ptr() if bias is not None else None
    if residual is not None:
        g.residual, g.ld_res = residual.data_ptr(), residual.stride(0)
    if aux is not None:
        g.aux, g.ld_aux = aux.data_ptr(), aux.stride(0)
    g.act = act
    if out_bf16 is not None:
        g.out_bf16, g.ld_out = out_bf16.data_ptr(), out_bf16.stride(0)
    if out_pre is not None:
        g.out_pre, g.ld_pre = out_pre.data_ptr(), out_pre.stride(0)
    if out_f32 is not None:
        g.out_f32, g.ld_f32 = out_f32.data_ptr(), out_f32.stride(0)
    g.accumulate, g.split_k, g.block_n = int(accumulate), split_k, block_n
    check(lib().uc2_gemm_bf16(C.byref(g), stream_ptr()), "gemm")
