"""ctypes binding of libdb1_sm100.so (the C ABI declared in include/db1_sm100.h).

The library is built in-tree by bdm-db1_b200/build.py. There is NO fallback: if the shared object is missing or a
call fails, the caller gets an exception (the product path never routes through PyTorch ops or the CPU oracle).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libdb1_sm100.so")
HOSTLIB_PATH = os.path.join(os.path.dirname(_HERE), "libdb1_host.so")


class Db1Error(RuntimeError):
    pass


class GemmDesc(C.Structure):
    """Mirror of `struct db1_gemm_desc` (include/db1_sm100.h)."""
    _fields_ = [
        ("epilogue", C.c_int32), ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("a_mn", C.c_int32), ("b_mn", C.c_int32),
        ("A", C.c_void_p), ("B", C.c_void_p), ("C", C.c_void_p),
        ("lda", C.c_int64), ("ldb", C.c_int64), ("ldc", C.c_int64),
        ("Z1", C.c_int32), ("Z2", C.c_int32),
        ("a_z1", C.c_int64), ("a_z2", C.c_int64), ("b_z1", C.c_int64), ("b_z2", C.c_int64),
        ("c_z1", C.c_int64), ("c_z2", C.c_int64),
        ("reduce_z2", C.c_int32), ("k_mode", C.c_int32), ("skip_upper", C.c_int32),
        ("alpha", C.c_float), ("accumulate", C.c_int32),
        ("bias", C.c_void_p), ("resid", C.c_void_p), ("ldr", C.c_int64),
        ("drop_p", C.c_float), ("seed", C.c_uint64),
        ("u", C.c_void_p), ("v", C.c_void_p), ("d_model", C.c_int32),
        ("H", C.c_void_p), ("ldh", C.c_int64), ("F", C.c_int32),
        ("P", C.c_void_p), ("C2", C.c_void_p), ("Drow", C.c_void_p),
        ("window", C.c_int32), ("bn_hint", C.c_int32),
        ("dot_with", C.c_void_p), ("ld_dot", C.c_int64), ("dot_out", C.c_void_p),
        ("dot_L", C.c_int32), ("dot_H", C.c_int32),
        ("b_static", C.c_int32),
        ("ln_gamma", C.c_void_p), ("ln_beta", C.c_void_p), ("ln_eps", C.c_float), ("ln_out", C.c_void_p),
    ]


EPI_PLAIN, EPI_QKV, EPI_GEGLU, EPI_DGEGLU, EPI_DS = 0, 1, 2, 3, 4
K_FULL, K_END_BY_ROW, K_BEGIN_BY_ROW, K_BEGIN_REV = 0, 1, 2, 3

_lib = None
_hostlib = None


def lib():
    """Load (once) and return the CUDA library; raises Db1Error if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Db1Error(
                "libdb1_sm100.so not found at %s - run `python bdm-db1_b200/build.py` "
                "(there is no PyTorch/CPU fallback for the DB1 hot path)" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        l.db1_last_error.restype = C.c_char_p
        l.db1_abi_version.restype = C.c_int
        _lib = l
    return _lib


def hostlib():
    global _hostlib
    if _hostlib is None:
        if not os.path.exists(HOSTLIB_PATH):
            raise Db1Error("libdb1_host.so not found at %s - run `python bdm-db1_b200/build.py`" % HOSTLIB_PATH)
        _hostlib = C.CDLL(HOSTLIB_PATH)
    return _hostlib


def check(rc, what):
    if rc != 0:
        msg = lib().db1_last_error().decode("utf-8", "replace")
        raise Db1Error("%s failed (rc=%d): %s" % (what, rc, msg))


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def cur_stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
