"""Thin Python wrappers: torch tensors in, C-ABI calls out. No numerics happen here.

Every function launches on torch's current CUDA stream and never synchronises. All tensors must be CUDA fp16 unless
stated; shapes/strides are passed explicitly so views (column slices of a fused buffer) work without copies.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import (EPI_DGEGLU, EPI_DS, EPI_GEGLU, EPI_PLAIN, EPI_QKV, K_BEGIN_BY_ROW, K_BEGIN_REV, K_END_BY_ROW,
                   K_FULL, Db1Error, GemmDesc, check, cur_stream, ptr)


class Profile:
    """Launch accounting and optional per-launch CUDA-event timing (events are recorded on the launching stream)."""

    def __init__(self, timing=False, by_shape=False):
        self.timing = timing
        self.by_shape = by_shape  # GEMM records are keyed by shape as well (development view)
        self.launches = 0
        self.records = []  # (name, algorithmic flops, algorithmic bytes, start event, end event)

    def summary(self):
        """{name: dict(n, ms, flops, bytes)} — call after torch.cuda.synchronize()."""
        out = {}
        for name, fl, by, e0, e1 in self.records:
            d = out.setdefault(name, dict(n=0, ms=0.0, flops=0.0, bytes=0.0))
            d["n"] += 1
            d["ms"] += e0.elapsed_time(e1)
            d["flops"] += fl
            d["bytes"] += by
        return out


_active = None


def set_profile(prof):
    global _active
    _active = prof
    return prof


class _Launch:
    """with _Launch(name, nkernels, flops, bytes): <C call>"""
    __slots__ = ("name", "n", "flops", "bytes", "e0")

    def __init__(self, name, n=1, flops=0.0, nbytes=0.0):
        self.name, self.n, self.flops, self.bytes = name, n, flops, nbytes

    def __enter__(self):
        pr = _active
        if pr is not None:
            pr.launches += self.n
            if pr.timing:
                self.e0 = torch.cuda.Event(enable_timing=True)
                self.e0.record()
        return self

    def __exit__(self, *exc):
        pr = _active
        if pr is not None and pr.timing and exc[0] is None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            pr.records.append((self.name, self.flops, self.bytes, self.e0, e1))
        return False


_EPI_NAMES = {0: "plain", 1: "qkv", 2: "geglu", 3: "dgeglu", 4: "ds"}


def _need_cuda_half(*ts):
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda or t.dtype != torch.float16:
            raise _lib.Db1Error("DB1 sm_100a kernels need CUDA fp16 tensors, got %s on %s" % (t.dtype, t.device))


_gemm_ws = {}  # device index -> (tensor, registered)


def _ensure_gemm_workspace(device):
    """Scratch for the GEMM's stream-K tail (include/db1_sm100.h:db1_gemm_set_workspace): allocated by torch once per
    device, zero-filled, kept alive for the life of the process. The library is bound to one device per process (one
    process per GPU); a second device simply runs without the stream-K tail."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx in _gemm_ws:
        return
    l = _lib.lib()
    l.db1_gemm_workspace_bytes.restype = C.c_longlong
    if _gemm_ws:  # already bound to another device
        _gemm_ws[idx] = None
        return
    with torch.cuda.device(idx):
        n = int(l.db1_gemm_workspace_bytes())
        ws = torch.zeros(n, dtype=torch.uint8, device=device)
        check(l.db1_gemm_set_workspace(ptr(ws), C.c_longlong(n)), "db1_gemm_set_workspace")
    _gemm_ws[idx] = ws


def gemm(A, B, C_out, M, N, K, *, lda, ldb, ldc, a_mn=False, b_mn=False, epilogue=EPI_PLAIN, alpha=1.0,
         accumulate=False, bias=None, resid=None, ldr=0, drop_p=0.0, seed=0, u=None, v=None, d_model=0, H=None,
         ldh=0, F=0, Z1=1, Z2=1, a_z=(0, 0), b_z=(0, 0), c_z=(0, 0), reduce_z2=False, k_mode=K_FULL,
         skip_upper=False, P=None, C2=None, Drow=None, window=0, bn_hint=0, dot=None, b_static=False, ln=None):
    """C[M,N] (+)= epilogue(alpha * A[M,K] @ B[N,K]^T) per batch index; see include/db1_sm100.h:db1_gemm_f16."""
    _need_cuda_half(A, B, C_out, bias, resid, u, v, H, P, C2)
    _ensure_gemm_workspace(A.device)
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
    d.b_static = int(b_static)  # B = weights nobody is writing (inference): the few-row path may prefetch them early
    if ln is not None:  # (gamma, beta, eps, out or None): LayerNorm of A's rows on load (few-row path, M <= 8)
        g_, b_, eps_, o_ = ln
        _need_cuda_half(g_, b_, o_)
        if not few_row_gemm_applies(M, K, lda, a_mn, b_mn, Z1 * Z2, accumulate, drop_p, epilogue):
            raise Db1Error("gemm(ln=...): LayerNorm-on-load exists on the few-row path only (M <= 8)")
        d.ln_gamma, d.ln_beta, d.ln_eps = g_.data_ptr(), b_.data_ptr(), float(eps_)
        d.ln_out = o_.data_ptr() if o_ is not None else None
    if dot is not None:  # (X fp16 [M, ld], out fp32 [M / L, H, L], L, H): out = per-head rowsum(C * X), head dim 128
        X, dout_t, dL, dH = dot
        _need_cuda_half(X)
        d.dot_with, d.ld_dot, d.dot_out, d.dot_L, d.dot_H = X.data_ptr(), X.stride(0), _f32(dout_t), dL, dH
    flops = 2.0 * M * N * K * Z1 * Z2
    if k_mode != K_FULL or skip_upper:
        flops *= (M + 1) / (2.0 * M)  # causal: only unmasked (i, j) pairs are algorithmic work
    name = "gemm_" + _EPI_NAMES[epilogue] + ("_batched" if Z1 * Z2 > 1 else "")
    if _active is not None and _active.timing and _active.by_shape:
        name += "_%dx%dx%d%s%s" % (M, N, K, "_Amn" if a_mn else "", "_Bmn" if b_mn else "")
    with _Launch(name, 1, flops):
        check(_lib.lib().db1_gemm_f16(C.byref(d), cur_stream()), "db1_gemm_f16")
    return C_out


def few_row_gemm_applies(M, K, lda, a_mn=False, b_mn=False, batch=1, accumulate=False, drop_p=0.0, epilogue=EPI_PLAIN):
    """Host-side mirror of skinny_gemm_applies (csrc/skinny.cu): whether db1_gemm_f16 will take the weight-streaming path."""
    import os
    return (1 <= M <= 8 and batch == 1 and not a_mn and not b_mn and not accumulate and drop_p == 0.0 and K % 8 == 0
            and lda % 8 == 0 and 16 * K <= 200 * 1024 and epilogue in (EPI_PLAIN, EPI_QKV, EPI_GEGLU)
            and not os.environ.get("DB1_NO_SKINNY"))


def gemm_dot_supported(M, N, H, dh):
    """Whether db1_gemm_f16 can emit the per-head row-dot next to C (include/db1_sm100.h: dot_with / dot_out)."""
    return dh == 128 and N == H * dh and M > 384 and 2 * ((M + 127) // 128) * ((N + 255) // 256) > 148


def relattn_fwd(qkv4, r, out, lse2, B, L, H, dh, window, scale, probs=None):
    """qkv4: [B*L, 4*H*dh] = [q+u | q+v | k | v]; r: [L, H*dh]. mode 0 (probs is None): out, lse2 written;
    mode 1: probs [B,H,L,L] written from lse2. include/db1_sm100.h:db1_relattn_fwd."""
    _need_cuda_half(qkv4, r, out, probs)
    d = H * dh
    es = qkv4.element_size()
    base = qkv4.data_ptr()
    ld = qkv4.stride(0)
    mode = 0 if probs is None else 1
    pairs = L * (L + 1) / 2.0 if window >= L else (window * (window + 1) / 2.0 + (L - window) * window)
    flops = B * H * pairs * dh * (6.0 if mode == 0 else 4.0)
    nbytes = B * H * (4.0 * L * dh * 2 + 4 * L) + H * L * dh * 2.0
    with _Launch("relattn_fwd" if mode == 0 else "relattn_probs", 1, flops, nbytes):
      check(_lib.lib().db1_relattn_fwd(
        C.c_void_p(base), C.c_void_p(base + d * es), C.c_void_p(base + 2 * d * es), C.c_void_p(base + 3 * d * es),
        C.c_longlong(ld), ptr(r), C.c_longlong(r.stride(0)), ptr(out), C.c_longlong(out.stride(0) if out is not None else 0),
        ptr(lse2), ptr(probs), B, L, H, dh, int(window), C.c_float(scale), mode, cur_stream()), "db1_relattn_fwd")


def relattn_mem_fwd(qkv4, r, out, lse2, B, K, H, dh, window, scale, mlen):
    """Attention over cat(mem, w) (K = mlen + qlen rows per sequence); only query rows >= mlen are produced.
    include/db1_sm100.h:db1_relattn_mem_fwd."""
    _need_cuda_half(qkv4, r, out)
    d = H * dh
    es = qkv4.element_size()
    base = qkv4.data_ptr()
    with _Launch("relattn_mem_fwd", 1):
      check(_lib.lib().db1_relattn_mem_fwd(
        C.c_void_p(base), C.c_void_p(base + d * es), C.c_void_p(base + 2 * d * es), C.c_void_p(base + 3 * d * es),
        C.c_longlong(qkv4.stride(0)), ptr(r), C.c_longlong(r.stride(0)), ptr(out), C.c_longlong(out.stride(0)),
        _f32(lse2), B, K, H, dh, int(window), C.c_float(scale), int(mlen), cur_stream()), "db1_relattn_mem_fwd")


def relattn_bwd_ds(qkv4, r, dout, lse2, drow, probs, ds, B, L, H, dh, window, scale, o=None):
    """P and dS = P * (dO V^T - D) * scale for the attention backward; include/db1_sm100.h:db1_relattn_bwd_ds.
    D = rowsum(dO * O) comes from `drow` (db1_rowdot) or, with drow=None, is formed in the kernel from `o`."""
    _need_cuda_half(qkv4, r, dout, probs, ds, o)
    d = H * dh
    es = qkv4.element_size()
    base = qkv4.data_ptr()
    pairs = L * (L + 1) / 2.0 if window >= L else (window * (window + 1) / 2.0 + (L - window) * window)
    if drow is None:
        with _Launch("relattn_bwd_ds", 1, B * H * pairs * dh * 6.0):
          check(_lib.lib().db1_relattn_bwd_ds_o(
            C.c_void_p(base), C.c_void_p(base + d * es), C.c_void_p(base + 2 * d * es), C.c_void_p(base + 3 * d * es),
            C.c_longlong(qkv4.stride(0)), ptr(r), C.c_longlong(r.stride(0)), ptr(dout), C.c_longlong(dout.stride(0)),
            ptr(o), C.c_longlong(o.stride(0)), _f32(lse2), ptr(probs), ptr(ds), B, L, H, dh, int(window),
            C.c_float(scale), cur_stream()), "db1_relattn_bwd_ds_o")
        return
    with _Launch("relattn_bwd_ds", 1, B * H * pairs * dh * 6.0):
      check(_lib.lib().db1_relattn_bwd_ds(
        C.c_void_p(base), C.c_void_p(base + d * es), C.c_void_p(base + 2 * d * es), C.c_void_p(base + 3 * d * es),
        C.c_longlong(qkv4.stride(0)), ptr(r), C.c_longlong(r.stride(0)), ptr(dout), C.c_longlong(dout.stride(0)),
        _f32(lse2), _f32(drow), ptr(probs), ptr(ds), B, L, H, dh, int(window), C.c_float(scale), cur_stream()),
        "db1_relattn_bwd_ds")


def relattn_bwd_dkdv(probs, ds, dout, qu_view, dk_view, dv_view, B, L, H, dh, window):
    """dv = P^T dO, dk = dS^T (q+u) per (sequence, head), key-outer with TMEM-resident accumulators;
    include/db1_sm100.h:db1_relattn_bwd_dkdv. dk_view / dv_view: fp16 views [B*L, H*dh] with a common row stride."""
    _need_cuda_half(probs, ds, dout, qu_view, dk_view, dv_view)
    if dk_view.stride(0) != dv_view.stride(0):
        raise _lib.Db1Error("relattn_bwd_dkdv: dk and dv views need the same row stride")
    pairs = L * (L + 1) / 2.0 if window >= L else (window * (window + 1) / 2.0 + (L - window) * window)
    with _Launch("relattn_bwd_dkdv", 1, B * H * pairs * dh * 4.0, 4.0 * B * H * pairs + 4.0 * B * L * H * dh * 2):
      check(_lib.lib().db1_relattn_bwd_dkdv(ptr(probs), ptr(ds), ptr(dout), C.c_longlong(dout.stride(0)), ptr(qu_view),
                                            C.c_longlong(qu_view.stride(0)), ptr(dk_view), ptr(dv_view),
                                            C.c_longlong(dk_view.stride(0)), B, L, H, dh, int(window), cur_stream()),
            "db1_relattn_bwd_dkdv")


def relattn_bwd_dq(ds, k_view, r, dq_view, du, dv, B, L, H, dh, window):
    """dq (fp16 view [B*L, H*dh] with its own row stride), du / dv (fp32 [H*dh], accumulated) from dS, K and r;
    include/db1_sm100.h:db1_relattn_bwd_dq. k_view: the k columns of the fused QKV buffer."""
    _need_cuda_half(ds, k_view, r, dq_view)
    pairs = L * (L + 1) / 2.0 if window >= L else (window * (window + 1) / 2.0 + (L - window) * window)
    with _Launch("relattn_bwd_dq", 1, B * H * pairs * dh * 4.0, 2.0 * B * H * pairs + 2.0 * B * L * H * dh * 2):
      check(_lib.lib().db1_relattn_bwd_dq(ptr(ds), ptr(k_view), C.c_longlong(k_view.stride(0)), ptr(r),
                                          C.c_longlong(r.stride(0)), ptr(dq_view), C.c_longlong(dq_view.stride(0)),
                                          _f32(du), _f32(dv), B, L, H, dh, int(window), cur_stream()), "db1_relattn_bwd_dq")


def relattn_bwd_dr(ds, qv_view, dr32, B, L, H, dh, window):
    """dr32 (fp32 [L, H*dh], zero-filled by the caller) += the relative-position gradient; db1_relattn_bwd_dr."""
    _need_cuda_half(ds, qv_view)
    pairs = L * (L + 1) / 2.0 if window >= L else (window * (window + 1) / 2.0 + (L - window) * window)
    with _Launch("relattn_bwd_dr", 1, B * H * pairs * dh * 2.0, 2.0 * B * H * pairs + 2.0 * B * L * H * dh):
      check(_lib.lib().db1_relattn_bwd_dr(ptr(ds), ptr(qv_view), C.c_longlong(qv_view.stride(0)), _f32(dr32),
                                          C.c_longlong(dr32.stride(0)), B, L, H, dh, int(window), cur_stream()),
            "db1_relattn_bwd_dr")


def score_tiles_shape(B, L, H):
    """Shape of the tiled P / dS scratch (include/db1_sm100.h:db1_relattn_bwd_ds_tiled)."""
    nq = (L + 127) // 128
    return (B, H, nq, nq, 2, 128, 64)


def relattn_bwd_ds_tiled(qkv4, r, dout, lse2, drow, probs, ds, B, L, H, dh, window, scale, o=None):
    """P and dS = P * (dO V^T - D) * scale in the tiled layout read by relattn_bwd_dkdv / _dq / _dr."""
    _need_cuda_half(qkv4, r, dout, probs, ds, o)
    d = H * dh
    es = qkv4.element_size()
    base = qkv4.data_ptr()
    pairs = L * (L + 1) / 2.0 if window >= L else (window * (window + 1) / 2.0 + (L - window) * window)
    with _Launch("relattn_bwd_ds", 1, B * H * pairs * dh * 6.0):
      check(_lib.lib().db1_relattn_bwd_ds_tiled(
        C.c_void_p(base), C.c_void_p(base + d * es), C.c_void_p(base + 2 * d * es), C.c_void_p(base + 3 * d * es),
        C.c_longlong(qkv4.stride(0)), ptr(r), C.c_longlong(r.stride(0)), ptr(dout), C.c_longlong(dout.stride(0)),
        ptr(o), C.c_longlong(o.stride(0) if o is not None else 0), _f32(lse2), _f32(drow) if drow is not None else None,
        ptr(probs), ptr(ds), B, L, H, dh, int(window), C.c_float(scale), cur_stream()), "db1_relattn_bwd_ds_tiled")


def _f32(t):
    if t is not None and (not t.is_cuda or t.dtype != torch.float32):
        raise _lib.Db1Error("expected a CUDA fp32 tensor, got %s on %s" % (t.dtype, t.device))
    return ptr(t)


def _i64(t):
    if t is not None and (not t.is_cuda or t.dtype != torch.int64 or not t.is_contiguous()):
        raise _lib.Db1Error("expected a contiguous CUDA int64 tensor")
    return ptr(t)


def layernorm_fwd(y, gamma, beta, out, stats, eps):
    _need_cuda_half(y, gamma, beta, out)
    rows, d = y.shape
    with _Launch("layernorm_fwd", 1):
      check(_lib.lib().db1_layernorm_fwd(ptr(y), ptr(gamma), ptr(beta), ptr(out), _f32(stats), rows, d, C.c_float(eps),
                                       cur_stream()), "db1_layernorm_fwd")


def layernorm_bwd(dout, y, gamma, stats, dy, dz, dgamma, dbeta, dbias, drop_p=0.0, seed=0):
    _need_cuda_half(dout, y, gamma, dy, dz)
    rows, d = y.shape
    with _Launch("layernorm_bwd", 1):
      check(_lib.lib().db1_layernorm_bwd(ptr(dout), ptr(y), ptr(gamma), _f32(stats), ptr(dy), ptr(dz), _f32(dgamma),
                                       _f32(dbeta), _f32(dbias), rows, d, C.c_float(drop_p), C.c_uint64(seed),
                                       cur_stream()), "db1_layernorm_bwd")


def ce_fwd(logits, labels, mask, row_loss, row_lse, loss2, V):
    _need_cuda_half(logits)
    rows = logits.shape[0]
    with _Launch("ce_fwd", 2):
      check(_lib.lib().db1_ce_fwd(ptr(logits), C.c_longlong(logits.stride(0)), _i64(labels), _f32(mask), _f32(row_loss),
                                _f32(row_lse), _f32(loss2), rows, V, cur_stream()), "db1_ce_fwd")


def ce_bwd(logits, labels, mask, row_lse, loss2, gscale, dlogits, V):
    _need_cuda_half(logits, dlogits)
    rows = logits.shape[0]
    with _Launch("ce_bwd", 1):
      check(_lib.lib().db1_ce_bwd(ptr(logits), C.c_longlong(logits.stride(0)), _i64(labels), _f32(mask), _f32(row_lse),
                                _f32(loss2), _f32(gscale), ptr(dlogits), C.c_longlong(dlogits.stride(0)), rows, V,
                                cur_stream()), "db1_ce_bwd")


def embed_fwd(tok, pos, slot, W, T, vis, out, out_bs, B, L, d, V, drop_p=0.0, seed=0, seed_row0=0):
    """out is addressed as out.data_ptr() + b*out_bs + l*d (elements); vis [B, nvis, d] contiguous or None."""
    _need_cuda_half(W, T, vis, out)
    nvis = vis.shape[1] if vis is not None else 0
    vis_bs = vis.stride(0) if vis is not None else 0
    with _Launch("embed_fwd", 2):
      check(_lib.lib().db1_embed_fwd(_i64(tok), _i64(pos), ptr(slot), ptr(W), ptr(T), ptr(vis), C.c_longlong(vis_bs), nvis,
                                   ptr(out), C.c_longlong(out_bs), B, L, d, V, C.c_float(drop_p), C.c_uint64(seed),
                                   C.c_longlong(seed_row0), cur_stream()), "db1_embed_fwd")


def embed_bwd(tok, pos, slot, dout, dout_bs, dW, dT, dvis, B, L, d, V, drop_p=0.0, seed=0, seed_row0=0):
    _need_cuda_half(dout, dW, dT, dvis)
    nvis = dvis.shape[1] if dvis is not None else 0
    vis_bs = dvis.stride(0) if dvis is not None else 0
    with _Launch("embed_bwd", 1):
      check(_lib.lib().db1_embed_bwd(_i64(tok), _i64(pos), ptr(slot), ptr(dout), C.c_longlong(dout_bs), ptr(dW), ptr(dT),
                                   ptr(dvis), C.c_longlong(vis_bs), nvis, B, L, d, V, C.c_float(drop_p),
                                   C.c_uint64(seed), C.c_longlong(seed_row0), cur_stream()), "db1_embed_bwd")


def colsum(x, out, rows, n):
    _need_cuda_half(x)
    with _Launch("colsum", 1):
      check(_lib.lib().db1_colsum(ptr(x), C.c_longlong(x.stride(0)), _f32(out), rows, n, cur_stream()), "db1_colsum")


def dq_finalize(dqu, dqv, dq, du, dv, rows, n):
    _need_cuda_half(dqu, dqv, dq)
    with _Launch("dq_finalize", 1):
      check(_lib.lib().db1_dq_finalize(ptr(dqu), ptr(dqv), C.c_longlong(dqu.stride(0)), ptr(dq), C.c_longlong(dq.stride(0)),
                                     _f32(du), _f32(dv), rows, n, cur_stream()), "db1_dq_finalize")


def rowdot(a, b, out, B, L, H, dh):
    _need_cuda_half(a, b)
    with _Launch("rowdot", 1):
      check(_lib.lib().db1_rowdot(ptr(a), ptr(b), C.c_longlong(a.stride(0)), _f32(out), B, L, H, dh, cur_stream()),
          "db1_rowdot")


def posemb(out, inv_freq, klen, d, clamp_len, drop_p=0.0, seed=0, half_phase=False):
    _need_cuda_half(out)
    fn = _lib.lib().db1_posemb_half_phase if half_phase else _lib.lib().db1_posemb
    with _Launch("posemb", 1):
      check(fn(ptr(out), _f32(inv_freq), klen, d, clamp_len, C.c_float(drop_p), C.c_uint64(seed), cur_stream()),
            "db1_posemb")


def f32_to_f16(src, dst, accumulate=False):
    _need_cuda_half(dst)
    with _Launch("f32_to_f16", 1):
      check(_lib.lib().db1_f32_to_f16(_f32(src), ptr(dst), C.c_longlong(src.numel()), int(accumulate), cur_stream()),
          "db1_f32_to_f16")


def rel_unshift(ds, dsr, Z, L):
    """dsr[z,i,c] = ds[z,i,c-(L-1-i)] (0 where c < L-1-i): relative-position re-layout of dS."""
    _need_cuda_half(ds, dsr)
    with _Launch("rel_unshift", 1, 0.0, 2.0 * Z * L * (L + 1)):
      check(_lib.lib().db1_rel_unshift(ptr(ds), ptr(dsr), Z, L, cur_stream()), "db1_rel_unshift")


def transpose(x, out, batch, rows, cols):
    """out[b][c][r] = x[b][r][c] (contiguous fp16)."""
    _need_cuda_half(x, out)
    with _Launch("transpose", 1):
      check(_lib.lib().db1_transpose_f16(ptr(x), ptr(out), batch, rows, cols, cur_stream()), "db1_transpose_f16")
    return out


def patch_conv1_fwd(pixels, W1, b1, xs, y1):
    """pixels: CUDA fp16 or fp32 [N,C,H,W] (fp32 frames are standardised from their unrounded values)."""
    _need_cuda_half(W1, b1, xs, y1)
    N, Cc, Hh, Ww = pixels.shape
    if pixels.dtype == torch.float32 and pixels.is_cuda:
        with _Launch("patch_conv1_fwd", 1):
          check(_lib.lib().db1_patch_conv1_fwd_f32(_f32(pixels), ptr(W1), ptr(b1), ptr(xs), ptr(y1), N, Cc, Hh, Ww,
                                                   cur_stream()), "db1_patch_conv1_fwd_f32")
        return
    _need_cuda_half(pixels)
    with _Launch("patch_conv1_fwd", 1):
      check(_lib.lib().db1_patch_conv1_fwd(ptr(pixels), ptr(W1), ptr(b1), ptr(xs), ptr(y1), N, Cc, Hh, Ww, cur_stream()),
          "db1_patch_conv1_fwd")


def patch_conv1_bwd(xs, dy1, dW1, P, Cc):
    _need_cuda_half(xs, dy1)
    with _Launch("patch_conv1_bwd", 1):
      check(_lib.lib().db1_patch_conv1_bwd(ptr(xs), ptr(dy1), _f32(dW1), P, Cc, cur_stream()), "db1_patch_conv1_bwd")


def gn_gelu_im2col(x, gamma, beta, col, stats, P, eps):
    _need_cuda_half(x, gamma, beta, col)
    with _Launch("gn_gelu_im2col", 1, 0.0, P * 256 * (64 + 576) * 2.0):
      check(_lib.lib().db1_gn_gelu_im2col(ptr(x), ptr(gamma), ptr(beta), ptr(col), _f32(stats), P, C.c_float(eps),
                                        cur_stream()), "db1_gn_gelu_im2col")


def col2im_gn_gelu_bwd(dcol, x, stats, gamma, beta, dres, dx, dgamma, dbeta, P):
    _need_cuda_half(dcol, x, gamma, beta, dres, dx)
    with _Launch("col2im_gn_gelu_bwd", 1, 0.0, P * 256 * (576 + 64 * 3) * 2.0):
      check(_lib.lib().db1_col2im_gn_gelu_bwd(ptr(dcol), ptr(x), _f32(stats), ptr(gamma), ptr(beta), ptr(dres), ptr(dx),
                                            _f32(dgamma), _f32(dbeta), P, cur_stream()), "db1_col2im_gn_gelu_bwd")


def dropout(x, out, p, seed):
    _need_cuda_half(x, out)
    with _Launch("dropout", 1):
      check(_lib.lib().db1_dropout_f16(ptr(x), ptr(out), C.c_longlong(x.numel()), C.c_float(p), C.c_uint64(seed),
                                     cur_stream()), "db1_dropout_f16")
    return out


def f32_to_f16_multi(src, segs):
    """segs: list of (dst fp16 tensor (contiguous), src offset, n, accumulate) - at most 8 per launch."""
    for i in range(0, len(segs), 8):
        part = segs[i:i + 8]
        k = len(part)
        _need_cuda_half(*[sg[0] for sg in part])
        dst = (C.c_void_p * k)(*[sg[0].data_ptr() for sg in part])
        off = (C.c_longlong * k)(*[sg[1] for sg in part])
        n = (C.c_longlong * k)(*[sg[2] for sg in part])
        acc = (C.c_int * k)(*[int(sg[3]) for sg in part])
        with _Launch("f32_to_f16", 1):
          check(_lib.lib().db1_f32_to_f16_multi(_f32(src), k, dst, off, n, acc, cur_stream()), "db1_f32_to_f16_multi")


def grad_sumsq(g16, out):
    _need_cuda_half(g16)
    with _Launch("grad_sumsq", 1, 0.0, 2.0 * g16.numel()):
      check(_lib.lib().db1_grad_sumsq(ptr(g16), C.c_longlong(g16.numel()), _f32(out), cur_stream()), "db1_grad_sumsq")


def clip_coef(sumsq, inv_scale, clip, gcoef2, flag):
    with _Launch("clip_coef", 1):
      check(_lib.lib().db1_clip_coef(_f32(sumsq), C.c_float(inv_scale), C.c_float(clip), _f32(gcoef2), ptr(flag),
                                   cur_stream()), "db1_clip_coef")


def adam_step(g16, p16, master, m, v, gcoef, lr, beta1, beta2, eps, weight_decay, step, adamw=True):
    _need_cuda_half(g16, p16)
    with _Launch("adam_step", 1, 0.0, 28.0 * g16.numel()):
      check(_lib.lib().db1_adam_step(ptr(g16), ptr(p16), _f32(master), _f32(m), _f32(v), C.c_longlong(g16.numel()),
                                   _f32(gcoef), C.c_float(lr), C.c_float(beta1), C.c_float(beta2), C.c_float(eps),
                                   C.c_float(weight_decay), int(step), int(adamw), cur_stream()), "db1_adam_step")


def decode_splits(B, Q, H):
    return int(_lib.lib().db1_decode_splits(B, Q, H))


def relattn_decode(qkv4, kcache, vcache, head, r, out, ws, B, Q, H, dh, window, scale, head_dev=None):
    """Few-query attention over [ring cache | new rows]; include/db1_sm100.h:db1_relattn_decode. qkv4: the new rows' fused
    [q+u | q+v | k | v] buffer [B*Q, 4*H*dh]; kcache / vcache [B, cap, H*dh]; head_dev: int32 device scalar or None."""
    _need_cuda_half(qkv4, kcache, vcache, r, out)
    d = H * dh
    es = qkv4.element_size()
    base = qkv4.data_ptr()
    cap = kcache.shape[1]
    nbytes = 2.0 * B * (cap + Q) * d * 2 + (cap + Q) * d * 2.0
    with _Launch("relattn_decode", 2, 0.0, nbytes):
      check(_lib.lib().db1_relattn_decode(
        C.c_void_p(base), C.c_void_p(base + d * es), C.c_void_p(base + 2 * d * es), C.c_void_p(base + 3 * d * es),
        C.c_longlong(qkv4.stride(0)), ptr(kcache), ptr(vcache), cap, int(head), ptr(head_dev), ptr(r),
        C.c_longlong(r.stride(0)), ptr(out), C.c_longlong(out.stride(0)), _f32(ws), C.c_longlong(ws.numel()), B, Q, H, dh,
        int(window), C.c_float(scale), cur_stream()), "db1_relattn_decode")


def ring_append(pairs, head, B, Q, head_dev=None):
    """pairs: up to three (src [B*Q, >= n] with its own row stride, ring [B, cap, n]); one launch."""
    k = len(pairs)
    _need_cuda_half(*[t for pr in pairs for t in pr])
    srcs = (C.c_void_p * k)(*[s_.data_ptr() for s_, _r in pairs])
    lds = (C.c_longlong * k)(*[s_.stride(0) for s_, _r in pairs])
    rings = (C.c_void_p * k)(*[r_.data_ptr() for _s, r_ in pairs])
    ring0 = pairs[0][1]
    with _Launch("ring_append", 1):
      check(_lib.lib().db1_ring_append(srcs, lds, rings, k, ring0.shape[1], int(head), ptr(head_dev), B, Q, ring0.shape[2],
                                       cur_stream()), "db1_ring_append")


def masked_argmax(logits2d, lo, hi, add_mask=None):
    """argmax over token columns [lo, hi) of every row of logits2d (fp16 [rows, ld]); returns int64 [rows]."""
    _need_cuda_half(logits2d)
    out = torch.empty(logits2d.shape[0], dtype=torch.int64, device=logits2d.device)
    with _Launch("masked_argmax", 1):
      check(_lib.lib().db1_masked_argmax(ptr(logits2d), C.c_longlong(logits2d.stride(0)), logits2d.shape[0], int(lo), int(hi),
                                         _f32(add_mask) if add_mask is not None else None, ptr(out), cur_stream()),
            "db1_masked_argmax")
    return out
