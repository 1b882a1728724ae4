"""Thin Python wrappers: torch tensors in, C-ABI calls out. No numerics happen here.

Every function launches on torch's current CUDA stream and never synchronises. All tensors must be CUDA fp16 unless
stated; shapes/strides are passed explicitly so views (column slices of a fused buffer) work without copies.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import (EPI_DGEGLU, EPI_DS, EPI_GEGLU, EPI_PLAIN, EPI_QKV, K_BEGIN_BY_ROW, K_BEGIN_REV, K_END_BY_ROW,
                   K_FULL, GemmDesc, check, cur_stream, ptr)


def _need_cuda_half(*ts):
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda or t.dtype != torch.float16:
            raise _lib.Db1Error("DB1 sm_100a kernels need CUDA fp16 tensors, got %s on %s" % (t.dtype, t.device))


def gemm(A, B, C_out, M, N, K, *, lda, ldb, ldc, a_mn=False, b_mn=False, epilogue=EPI_PLAIN, alpha=1.0,
         accumulate=False, bias=None, resid=None, ldr=0, drop_p=0.0, seed=0, u=None, v=None, d_model=0, H=None,
         ldh=0, F=0, Z1=1, Z2=1, a_z=(0, 0), b_z=(0, 0), c_z=(0, 0), reduce_z2=False, k_mode=K_FULL,
         skip_upper=False, P=None, C2=None, Drow=None, window=0, bn_hint=0):
    """C[M,N] (+)= epilogue(alpha * A[M,K] @ B[N,K]^T) per batch index; see include/db1_sm100.h:db1_gemm_f16."""
    _need_cuda_half(A, B, C_out, bias, resid, u, v, H, P, C2)
    d = GemmDesc()
    d.epilogue = epilogue
    d.M, d.N, d.K = M, N, K
    d.a_mn, d.b_mn = int(a_mn), int(b_mn)
    d.A, d.B, d.C = A.data_ptr(), B.data_ptr(), C_out.data_ptr()
    d.lda, d.ldb, d.ldc = lda, ldb, ldc
    d.Z1, d.Z2 = Z1, Z2
    d.a_z1, d.a_z2 = a_z
    d.b_z1, d.b_z2 = b_z
    d.c_z1, d.c_z2 = c_z
    d.reduce_z2 = int(reduce_z2)
    d.k_mode = k_mode
    d.skip_upper = int(skip_upper)
    d.alpha = alpha
    d.accumulate = int(accumulate)
    d.bias = bias.data_ptr() if bias is not None else None
    d.resid = resid.data_ptr() if resid is not None else None
    d.ldr = ldr
    d.drop_p = drop_p
    d.seed = seed
    d.u = u.data_ptr() if u is not None else None
    d.v = v.data_ptr() if v is not None else None
    d.d_model = d_model
    d.H = H.data_ptr() if H is not None else None
    d.ldh = ldh
    d.F = F
    d.P = P.data_ptr() if P is not None else None
    d.C2 = C2.data_ptr() if C2 is not None else None
    d.Drow = Drow.data_ptr() if Drow is not None else None
    d.window = window
    d.bn_hint = bn_hint
    check(_lib.lib().db1_gemm_f16(C.byref(d), cur_stream()), "db1_gemm_f16")
    return C_out


def relattn_fwd(qkv4, r, out, lse2, B, L, H, dh, window, scale, probs=None):
    """qkv4: [B*L, 4*H*dh] = [q+u | q+v | k | v]; r: [L, H*dh]. mode 0 (probs is None): out, lse2 written;
    mode 1: probs [B,H,L,L] written from lse2. include/db1_sm100.h:db1_relattn_fwd."""
    _need_cuda_half(qkv4, r, out, probs)
    d = H * dh
    es = qkv4.element_size()
    base = qkv4.data_ptr()
    ld = qkv4.stride(0)
    mode = 0 if probs is None else 1
    check(_lib.lib().db1_relattn_fwd(
        C.c_void_p(base), C.c_void_p(base + d * es), C.c_void_p(base + 2 * d * es), C.c_void_p(base + 3 * d * es),
        C.c_longlong(ld), ptr(r), C.c_longlong(r.stride(0)), ptr(out), C.c_longlong(out.stride(0) if out is not None else 0),
        ptr(lse2), ptr(probs), B, L, H, dh, int(window), C.c_float(scale), mode, cur_stream()), "db1_relattn_fwd")
