"""GPU parity of the tcgen05 GEMM (db1_gemm_f16) against torch fp32 matmuls on the same fp16 inputs."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return (a.float() - b.float()).abs().max().item() / (b.float().abs().max().item() + 1e-12)


def _mk(shape, dev, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).half().to(dev)


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 520, 200), (4096, 2048, 2048), (1024, 6144, 2048)])
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_gemm_layouts(cuda, M, N, K, a_mn, b_mn):
    from db1_sm100 import ops
    if (a_mn or b_mn) and (M % 8 or N % 8):
        pytest.skip("MN-major operands need 16-byte aligned rows")
    A = _mk((M, K), cuda, 1.0, 1)
    B = _mk((N, K), cuda, 1.0, 2)
    As = A.t().contiguous() if a_mn else A
    Bs = B.t().contiguous() if b_mn else B
    Cc = torch.empty(M, N, dtype=torch.half, device=cuda)
    ops.gemm(As, Bs, Cc, M, N, K, lda=As.stride(0), ldb=Bs.stride(0), ldc=N, a_mn=a_mn, b_mn=b_mn)
    ref = A.float() @ B.float().t()
    assert _rel(Cc, ref) < 2e-3


def test_gemm_bn128_and_ragged_n(cuda):
    from db1_sm100 import ops
    M, N, K = 500, 1001, 136
    A = _mk((M, K), cuda, 1.0, 3)
    B = _mk((N, K), cuda, 1.0, 4)
    ldc = 1008
    Cc = torch.zeros(M, ldc, dtype=torch.half, device=cuda)
    ops.gemm(A, B, Cc, M, N, K, lda=K, ldb=K, ldc=ldc)
    ref = A.float() @ B.float().t()
    assert _rel(Cc[:, :N], ref) < 2e-3
    assert Cc[:, N:].abs().max().item() == 0
    C2 = torch.zeros(M, ldc, dtype=torch.half, device=cuda)
    ops.gemm(A, B, C2, M, N, K, lda=K, ldb=K, ldc=ldc, bn_hint=128)
    assert _rel(C2[:, :N], ref) < 2e-3


def test_gemm_bias_resid_accumulate(cuda):
    from db1_sm100 import ops
    M, N, K = 384, 512, 320
    A = _mk((M, K), cuda, 1.0, 5)
    B = _mk((N, K), cuda, 0.1, 6)
    bias = _mk((N,), cuda, 1.0, 7)
    resid = _mk((M, N), cuda, 1.0, 8)
    C0 = _mk((M, N), cuda, 1.0, 9)
    Cc = C0.clone()
    ops.gemm(A, B, Cc, M, N, K, lda=K, ldb=K, ldc=N, alpha=0.5, bias=bias, resid=resid, ldr=N, accumulate=True)
    ref = 0.5 * (A.float() @ B.float().t()) + bias.float() + resid.float() + C0.float()
    assert _rel(Cc, ref) < 2e-3


def test_gemm_dropout_mask_is_reproducible(cuda):
    from db1_sm100 import ops
    M, N, K = 256, 512, 128
    A = _mk((M, K), cuda, 1.0, 10)
    B = _mk((N, K), cuda, 1.0, 11)
    C1 = torch.empty(M, N, dtype=torch.half, device=cuda)
    C2 = torch.empty_like(C1)
    C3 = torch.empty_like(C1)
    ops.gemm(A, B, C1, M, N, K, lda=K, ldb=K, ldc=N, drop_p=0.25, seed=1234)
    ops.gemm(A, B, C2, M, N, K, lda=K, ldb=K, ldc=N, drop_p=0.25, seed=1234)
    ops.gemm(A, B, C3, M, N, K, lda=K, ldb=K, ldc=N)
    assert torch.equal(C1, C2)
    keep = (C1 != 0)
    frac = keep.float().mean().item()
    assert abs(frac - 0.75) < 0.01
    assert _rel(C1[keep], (C3.float() / 0.75)[keep]) < 2e-3


def test_gemm_qkv_epilogue(cuda):
    from db1_sm100 import ops
    M, d, K = 512, 256, 256
    A = _mk((M, K), cuda, 1.0, 12)
    W = _mk((3 * d, K), cuda, 0.1, 13)
    u = _mk((d,), cuda, 1.0, 14)
    v = _mk((d,), cuda, 1.0, 15)
    Cc = torch.empty(M, 4 * d, dtype=torch.half, device=cuda)
    ops.gemm(A, W, Cc, M, 3 * d, K, lda=K, ldb=K, ldc=4 * d, epilogue=ops.EPI_QKV, u=u, v=v, d_model=d)
    ref = A.float() @ W.float().t()
    q, k, vv = ref[:, :d], ref[:, d:2 * d], ref[:, 2 * d:]
    refc = torch.cat([q + u.float(), q + v.float(), k, vv], 1)
    assert _rel(Cc, refc) < 2e-3


def test_gemm_geglu_fwd_bwd(cuda):
    from db1_sm100 import ops
    M, F, K = 384, 256, 192
    A = _mk((M, K), cuda, 1.0, 16)
    W1 = _mk((2 * F, K), cuda, 0.1, 17)
    b1 = _mk((2 * F,), cuda, 0.5, 18)
    H = torch.empty(M, 2 * F, dtype=torch.half, device=cuda)
    Y = torch.empty(M, F, dtype=torch.half, device=cuda)
    ops.gemm(A, W1, Y, M, 2 * F, K, lda=K, ldb=K, ldc=F, epilogue=ops.EPI_GEGLU, bias=b1, H=H, ldh=2 * F, F=F)
    h = A.float() @ W1.float().t() + b1.float()
    a, g = h[:, :F], h[:, F:]
    y = a * torch.nn.functional.gelu(g)
    assert _rel(H, h) < 2e-3
    assert _rel(Y, y) < 3e-3
    # backward epilogue: dH = [dY*gelu(g) | dY*a*gelu'(g)] with dY = G @ W2 (W2 stored [N2, F] -> MN-major B)
    N2 = 320
    G = _mk((M, N2), cuda, 1.0, 19)
    W2 = _mk((N2, F), cuda, 0.1, 20)
    dH = torch.empty(M, 2 * F, dtype=torch.half, device=cuda)
    ops.gemm(G, W2, dH, M, F, N2, lda=N2, ldb=F, ldc=2 * F, b_mn=True, epilogue=ops.EPI_DGEGLU, H=H, ldh=2 * F, F=F)
    dY = G.float() @ W2.float()
    hh = H.float().requires_grad_(True)
    yy = hh[:, :F] * torch.nn.functional.gelu(hh[:, F:])
    yy.backward(dY)
    assert _rel(dH, hh.grad) < 3e-3


def _attn_ref(L, dh, B, Hh, dev, window):
    torch.manual_seed(0)
    dO = _mk((B, L, Hh, dh), dev, 1.0, 21)
    V = _mk((B, L, Hh, dh), dev, 1.0, 22)
    P = torch.softmax(torch.randn(B, Hh, L, L, device=dev) + torch.triu(torch.full((L, L), -1e30, device=dev), 1), -1)
    P = P.half()
    D = torch.randn(B, Hh, L, device=dev) * 0.1
    return dO, V, P, D


@pytest.mark.parametrize("L,dh,window", [(256, 128, 1 << 30), (384, 64, 100), (200, 32, 1 << 30)])
def test_gemm_ds_epilogue_and_causal_kmodes(cuda, L, dh, window):
    """The attention-backward GEMM chain on materialised P: dS/dS_rel (DS epilogue), then the causal contractions."""
    from db1_sm100 import ops
    B, Hh = 2, 3
    d = Hh * dh
    dO, V, P, D = _attn_ref(L, dh, B, Hh, cuda, window)
    scale = 1.0 / math.sqrt(dh)
    dS = torch.zeros(B, Hh, L, L, dtype=torch.half, device=cuda)
    dSr = torch.zeros(B, Hh, L, L, dtype=torch.half, device=cuda)
    ops.gemm(dO, V, dS, L, L, dh, lda=d, ldb=d, ldc=L, epilogue=ops.EPI_DS, alpha=scale, Z1=Hh, Z2=B,
             a_z=(dh, L * d), b_z=(dh, L * d), c_z=(L * L, Hh * L * L), skip_upper=True, P=P, Drow=D,
             window=window)
    ops.rel_unshift(dS, dSr, B * Hh, L)
    i = torch.arange(L, device=cuda)[:, None]
    j = torch.arange(L, device=cuda)[None, :]
    ok = (j <= i) & ((i - j) < window)
    dP = torch.einsum("bihd,bjhd->bhij", dO.float(), V.float())
    ref = torch.where(ok, P.float() * (dP - D[..., None]) * scale, torch.zeros((), device=cuda))
    assert _rel(dS, ref) < 3e-3
    # relative-position order: dSr[i, j + L-1-i] = dS[i, j]
    ref_r = torch.zeros_like(ref)
    ii, jj = torch.nonzero(j <= i, as_tuple=True)
    ref_r[:, :, ii, jj + L - 1 - ii] = ref[:, :, ii, jj]
    assert _rel(dSr, ref_r) < 3e-3

    # dV[j] = sum_{i>=j} P[i,j] dO[i]   (A = P^T MN-major, B = dO MN-major, k begins at the row tile)
    dV = torch.empty(B, L, Hh, dh, dtype=torch.half, device=cuda)
    ops.gemm(P, dO, dV, L, dh, L, lda=L, ldb=d, ldc=d, a_mn=True, b_mn=True, Z1=Hh, Z2=B,
             a_z=(L * L, Hh * L * L), b_z=(dh, L * d), c_z=(dh, L * d), k_mode=ops.K_BEGIN_BY_ROW)
    refV = torch.einsum("bhij,bihd->bjhd", P.float(), dO.float())
    assert _rel(dV, refV) < 3e-3
    # dQ[i] = sum_{j<=i} dS[i,j] K[j]   (A = dS K-major, B = K MN-major, k ends at the row tile)
    Kt = V
    dQ = torch.empty(B, L, Hh, dh, dtype=torch.half, device=cuda)
    ops.gemm(dS, Kt, dQ, L, dh, L, lda=L, ldb=d, ldc=d, b_mn=True, Z1=Hh, Z2=B,
             a_z=(L * L, Hh * L * L), b_z=(dh, L * d), c_z=(dh, L * d), k_mode=ops.K_END_BY_ROW)
    refQ = torch.einsum("bhij,bjhd->bihd", dS.float(), Kt.float())
    assert _rel(dQ, refQ) < 3e-3
    # dQv[i] = sum_c dSr[i,c] R[c]   (R shared over the batch: broadcast z2; k begins at K-(mt+1)*128)
    R = _mk((L, Hh, dh), cuda, 1.0, 23)
    dQv = torch.empty(B, L, Hh, dh, dtype=torch.half, device=cuda)
    ops.gemm(dSr, R, dQv, L, dh, L, lda=L, ldb=d, ldc=d, b_mn=True, Z1=Hh, Z2=B,
             a_z=(L * L, Hh * L * L), b_z=(dh, 0), c_z=(dh, L * d), k_mode=ops.K_BEGIN_REV)
    refQv = torch.einsum("bhic,chd->bihd", dSr.float(), R.float())
    assert _rel(dQv, refQv) < 3e-3
    # dR[c] = sum_b sum_i dSr[b,i,c] Qv[b,i]   (contraction over z2 as well)
    dR = torch.empty(L, Hh, dh, dtype=torch.half, device=cuda)
    ops.gemm(dSr, dO, dR, L, dh, L, lda=L, ldb=d, ldc=d, a_mn=True, b_mn=True, Z1=Hh, Z2=B,
             a_z=(L * L, Hh * L * L), b_z=(dh, L * d), c_z=(dh, 0), reduce_z2=True, k_mode=ops.K_BEGIN_REV)
    refR = torch.einsum("bhic,bihd->chd", dSr.float(), dO.float())
    assert _rel(dR, refR) < 3e-3


# ---------------------------------------------------------------------------------------------------------------------
# Stream-K tail (CTA-pair launches whose pair-tile count is not a multiple of 74): same results as the data-parallel
# schedule up to fp32 summation order; repeated launches must re-arm the arrival counters.
# ---------------------------------------------------------------------------------------------------------------------
def _sk_expected(M, N, K):
    import ctypes as C
    from db1_sm100 import _lib
    R, GS = C.c_int(0), C.c_int(0)
    tiles = ((M + 127) // 128 + 1) // 2 * ((N + 255) // 256)
    return _lib.lib().db1_gemm_sk_choose(C.c_longlong(tiles), 74, (K + 63) // 64, C.byref(R), C.byref(GS)), R.value, GS.value


@pytest.mark.parametrize("M,N,K,a_mn,b_mn", [
    (4096, 2048, 2048, 0, 0),   # o_net: 128 pair-tiles, 54 in the tail, every pair both contributes and owns
    (4096, 2048, 4096, 0, 1),   # dgrad layout
    (2048, 2048, 4096, 1, 1),   # wgrad: 64 tiles on 74 pairs (no data-parallel part at all)
    (6144, 2048, 4096, 1, 1),   # 192 tiles: 2 DP waves + 44-tile tail
    (4096, 2048, 6144, 0, 1),   # KB = 96
    (2560, 2048, 512, 0, 0),    # 80 tiles, KB = 8: the 6 tail tiles are cut in two, most pairs only do their DP tile
    (4000, 2040, 1096, 0, 0),   # ragged M / N / K
])
def test_gemm_stream_k_tail(cuda, M, N, K, a_mn, b_mn, monkeypatch):
    from db1_sm100 import ops
    monkeypatch.setenv("DB1_GEMM_SK", "1")  # opt-in: measured slower than the data-parallel schedule on B200
    on, R, GS = _sk_expected(M, N, K)
    assert on == 1, "shape does not exercise the stream-K tail"
    A = _mk((M, K), cuda, 1.0, 31)
    B = _mk((N, K), cuda, 1.0, 32)
    As = A.t().contiguous() if a_mn else A
    Bs = B.t().contiguous() if b_mn else B
    ref = A.float() @ B.float().t()
    outs = []
    for _rep in range(3):  # back-to-back launches: the counters re-arm themselves
        Cc = torch.empty(M, N, dtype=torch.half, device=cuda)
        ops.gemm(As, Bs, Cc, M, N, K, lda=As.stride(0), ldb=Bs.stride(0), ldc=N, a_mn=a_mn, b_mn=b_mn)
        outs.append(Cc)
    torch.cuda.synchronize()
    for Cc in outs:
        assert _rel(Cc, ref) < 2e-3
        assert torch.equal(Cc, outs[0]), "stream-K sums must be deterministic"
    ws = ops._gemm_ws[cuda.index if cuda.index is not None else 0]
    assert ws[:4096].view(torch.int32).abs().max().item() == 0, "arrival counters not re-armed"


@pytest.mark.parametrize("sk", [0, 1])
def test_gemm_stream_k_fused_epilogues(cuda, sk, monkeypatch):
    """CTA-pair kernels at model size, with and without the stream-K tail (owner tiles run the normal fused epilogues
    after the fix-up): bias + dropout + residual (TMA epilogue), QKV, GeGLU backward (TMA epilogue)."""
    from db1_sm100 import ops
    if sk:
        monkeypatch.setenv("DB1_GEMM_SK", "1")
    M, N, K = 4096, 2048, 2048
    assert _sk_expected(M, N, K)[0] == 1
    A = _mk((M, K), cuda, 1.0, 33)
    B = _mk((N, K), cuda, 0.05, 34)
    bias = _mk((N,), cuda, 1.0, 35)
    resid = _mk((M, N), cuda, 1.0, 36)
    C1 = torch.empty(M, N, dtype=torch.half, device=cuda)
    ops.gemm(A, B, C1, M, N, K, lda=K, ldb=K, ldc=N, bias=bias, resid=resid, ldr=N)
    ref = A.float() @ B.float().t() + bias.float() + resid.float()
    assert _rel(C1, ref) < 2e-3
    C2 = torch.empty_like(C1)
    ops.gemm(A, B, C2, M, N, K, lda=K, ldb=K, ldc=N, bias=bias, resid=resid, ldr=N, drop_p=0.1, seed=77)
    keep = (C2 != resid)
    assert abs(keep.float().mean().item() - 0.9) < 0.01
    refd = (A.float() @ B.float().t() + bias.float()) / 0.9 + resid.float()
    assert _rel(C2[keep], refd[keep]) < 2e-3
    # QKV epilogue: 3*d columns, d = 1024 -> 16 x 12 = 192 pair-tiles
    d = 1024
    assert _sk_expected(M, 3 * d, d)[0] == 1
    X = _mk((M, d), cuda, 1.0, 37)
    W = _mk((3 * d, d), cuda, 0.05, 38)
    u = _mk((d,), cuda, 1.0, 39)
    v = _mk((d,), cuda, 1.0, 40)
    Q = torch.empty(M, 4 * d, dtype=torch.half, device=cuda)
    ops.gemm(X, W, Q, M, 3 * d, d, lda=d, ldb=d, ldc=4 * d, epilogue=ops.EPI_QKV, u=u, v=v, d_model=d)
    r = X.float() @ W.float().t()
    refq = torch.cat([r[:, :d] + u.float(), r[:, :d] + v.float(), r[:, d:2 * d], r[:, 2 * d:]], 1)
    assert _rel(Q, refq) < 2e-3
    # GeGLU backward epilogue: M x F accumulator, F = 4096 -> 256 pair-tiles (34-tile tail)
    F, N2 = 4096, 1024
    assert _sk_expected(M, F, N2)[0] == 1
    Hs = _mk((M, 2 * F), cuda, 1.0, 41)
    G = _mk((M, N2), cuda, 1.0, 42)
    W2 = _mk((N2, F), cuda, 0.05, 43)
    dH = torch.empty(M, 2 * F, dtype=torch.half, device=cuda)
    ops.gemm(G, W2, dH, M, F, N2, lda=N2, ldb=F, ldc=2 * F, b_mn=True, epilogue=ops.EPI_DGEGLU, H=Hs, ldh=2 * F, F=F)
    dY = G.float() @ W2.float()
    hh = Hs.float().requires_grad_(True)
    (hh[:, :F] * torch.nn.functional.gelu(hh[:, F:])).backward(dY)
    assert _rel(dH, hh.grad) < 3e-3


def test_gemm_tma_epilogue_ragged_and_accumulate(cuda):
    """CTA-pair PLAIN / GeGLU-backward kernels (TMA epilogue) on shapes whose last tiles are ragged in M, N and K, and
    the C += path (old C fetched by TMA)."""
    from db1_sm100 import ops
    M, N, K = 2000, 2072, 520
    A = _mk((M, K), cuda, 1.0, 51)
    B = _mk((N, K), cuda, 0.05, 52)
    C0 = _mk((M, N), cuda, 1.0, 53)
    Cc = C0.clone()
    ops.gemm(A, B, Cc, M, N, K, lda=K, ldb=K, ldc=N, accumulate=True)
    assert _rel(Cc, A.float() @ B.float().t() + C0.float()) < 2e-3
    # odd N (no 16-byte row alignment of the logical width; ldc padded): the head shape in miniature
    N2, ld2 = 2057, 2064
    B2 = _mk((N2, K), cuda, 0.05, 54)
    C2 = torch.zeros(M, ld2, dtype=torch.half, device=cuda)
    ops.gemm(A, B2, C2, M, N2, K, lda=K, ldb=K, ldc=ld2)
    assert _rel(C2[:, :N2], A.float() @ B2.float().t()) < 2e-3
    assert C2[:, N2:].abs().max().item() == 0
    # GeGLU backward, F not a multiple of 32
    F, N3 = 2072, 328
    Hs = _mk((M, 2 * F), cuda, 1.0, 55)
    G = _mk((M, N3), cuda, 1.0, 56)
    W2 = _mk((N3, F), cuda, 0.05, 57)
    dH = torch.full((M, 2 * F), 7.0, dtype=torch.half, device=cuda)
    ops.gemm(G, W2, dH, M, F, N3, lda=N3, ldb=F, ldc=2 * F, b_mn=True, epilogue=ops.EPI_DGEGLU, H=Hs, ldh=2 * F, F=F)
    dY = G.float() @ W2.float()
    hh = Hs.float().requires_grad_(True)
    (hh[:, :F] * torch.nn.functional.gelu(hh[:, F:])).backward(dY)
    assert _rel(dH, hh.grad) < 3e-3


def test_gemm_rowdot_side_output(cuda):
    """dO = dZ . W_o with D[b,h,i] = sum over head h of dO[b*L+i, :] * O[b*L+i, :] from the same (TMA) epilogue."""
    from db1_sm100 import ops
    Bq, L, Hh, dh, K = 4, 512, 16, 128, 320
    M, N = Bq * L, Hh * dh
    assert ops.gemm_dot_supported(M, N, Hh, dh)
    A = _mk((M, K), cuda, 1.0, 61)
    W = _mk((K, N), cuda, 0.1, 62)  # stored [K][N]: MN-major B, as W_o is in the dgrad
    O = _mk((M, N), cuda, 1.0, 63)
    Cc = torch.empty(M, N, dtype=torch.half, device=cuda)
    D = torch.full((Bq, Hh, L), float("nan"), dtype=torch.float32, device=cuda)
    ops.gemm(A, W, Cc, M, N, K, lda=K, ldb=N, ldc=N, b_mn=True, dot=(O, D, L, Hh))
    ref = A.float() @ W.float()
    assert _rel(Cc, ref) < 2e-3
    refD = (ref * O.float()).view(Bq, L, Hh, dh).sum(-1).permute(0, 2, 1)
    assert _rel(D, refD) < 2e-3


@pytest.mark.parametrize("M", [1, 3, 8])
def test_few_row_gemm_matches_fp32_and_the_tensor_core_path(cuda, M):
    """db1_gemm_f16 with M <= 8 (the decode step) runs the weight-streaming kernel of csrc/skinny.cu: plain (+bias,
    +residual, alpha, ragged N as the tied head has), the QKV split (+u / +v) and GeGLU (H and a * gelu(g)) against fp32
    torch on the same fp16 inputs; M = 9 of the same problem goes through the tcgen05 kernel and must agree."""
    import torch.nn.functional as Fn
    from db1_sm100 import ops
    d, K, F = 256, 512, 384
    A9 = _mk((9, K), cuda, 1.0, 21)
    A = A9[:M].contiguous()
    # plain, ragged N, strided C
    N = 1001
    B = _mk((N, K), cuda, 0.2, 22)
    bias = _mk((N + 7,), cuda, 1.0, 23)[:N]
    resid = _mk((9, 1008), cuda, 1.0, 24)
    C1 = torch.zeros(M, 1008, dtype=torch.half, device=cuda)
    ops.gemm(A, B, C1, M, N, K, lda=K, ldb=K, ldc=1008, alpha=0.5)
    assert _rel(C1[:, :N], 0.5 * (A.float() @ B.float().t())) < 2e-3
    assert C1[:, N:].abs().max().item() == 0
    Nb = 1000
    C2 = torch.zeros(M, 1008, dtype=torch.half, device=cuda)
    ops.gemm(A, B, C2, M, Nb, K, lda=K, ldb=K, ldc=1008, bias=bias[:Nb], resid=resid[:M], ldr=1008)
    ref2 = A.float() @ B[:Nb].float().t() + bias[:Nb].float() + resid[:M, :Nb].float()
    assert _rel(C2[:, :Nb], ref2) < 2e-3
    C2t = torch.zeros(9, 1008, dtype=torch.half, device=cuda)
    ops.gemm(A9, B, C2t, 9, Nb, K, lda=K, ldb=K, ldc=1008, bias=bias[:Nb], resid=resid, ldr=1008)
    assert _rel(C2[:, :Nb], C2t[:M, :Nb]) < 2e-3
    # QKV split
    Wqkv = _mk((3 * d, K), cuda, 0.2, 25)
    u, v = _mk((d,), cuda, 1.0, 26), _mk((d,), cuda, 1.0, 27)
    Q4 = torch.empty(M, 4 * d, dtype=torch.half, device=cuda)
    ops.gemm(A, Wqkv, Q4, M, 3 * d, K, lda=K, ldb=K, ldc=4 * d, epilogue=ops.EPI_QKV, u=u, v=v, d_model=d)
    qkv = A.float() @ Wqkv.float().t()
    refq = torch.cat([qkv[:, :d] + u.float(), qkv[:, :d] + v.float(), qkv[:, d:]], 1)
    assert _rel(Q4, refq) < 2e-3
    # GeGLU
    W1 = _mk((2 * F, K), cuda, 0.1, 28)
    b1 = _mk((2 * F,), cuda, 0.5, 29)
    Hb = torch.empty(M, 2 * F, dtype=torch.half, device=cuda)
    G = torch.empty(M, F, dtype=torch.half, device=cuda)
    ops.gemm(A, W1, G, M, 2 * F, K, lda=K, ldb=K, ldc=F, epilogue=ops.EPI_GEGLU, bias=b1, H=Hb, ldh=2 * F, F=F)
    h = A.float() @ W1.float().t() + b1.float()
    assert _rel(Hb, h) < 2e-3
    hh = Hb.float()
    assert _rel(G, hh[:, :F] * Fn.gelu(hh[:, F:])) < 2e-3


@pytest.mark.parametrize("M", [1, 5])
def test_few_row_gemm_layernorm_on_load(cuda, M):
    """db1_gemm_desc.ln_*: the few-row path LayerNorms A's rows while staging them (what lets the decode step drop its
    LayerNorm launches); product and the written-out normalised rows against torch on the same fp16 inputs, for the
    plain, QKV and GeGLU epilogues; the tensor-core path (M > 8) refuses the option."""
    import torch.nn.functional as Fn
    from db1_sm100 import ops
    from db1_sm100._lib import Db1Error
    K, N, d, F = 512, 1000, 256, 384
    A = _mk((M, K), cuda, 2.0, 31) + 0.5
    gamma, beta = _mk((K,), cuda, 0.3, 32) + 1.0, _mk((K,), cuda, 0.3, 33)
    ref_ln = Fn.layer_norm(A.float(), (K,), gamma.float(), beta.float(), 1e-5)
    ln16 = ref_ln.half()
    B = _mk((N, K), cuda, 0.1, 34)
    C1 = torch.empty(M, N, dtype=torch.half, device=cuda)
    lnout = torch.zeros(M, K, dtype=torch.half, device=cuda)
    ops.gemm(A, B, C1, M, N, K, lda=K, ldb=K, ldc=N, ln=(gamma, beta, 1e-5, lnout), b_static=True)
    assert _rel(lnout, ref_ln) < 2e-3
    assert _rel(C1, ln16.float() @ B.float().t()) < 3e-3
    Wqkv = _mk((3 * d, K), cuda, 0.1, 35)
    u, v = _mk((d,), cuda, 1.0, 36), _mk((d,), cuda, 1.0, 37)
    Q4 = torch.empty(M, 4 * d, dtype=torch.half, device=cuda)
    ops.gemm(A, Wqkv, Q4, M, 3 * d, K, lda=K, ldb=K, ldc=4 * d, epilogue=ops.EPI_QKV, u=u, v=v, d_model=d,
             ln=(gamma, beta, 1e-5, None))
    qkv = ln16.float() @ Wqkv.float().t()
    assert _rel(Q4, torch.cat([qkv[:, :d] + u.float(), qkv[:, :d] + v.float(), qkv[:, d:]], 1)) < 3e-3
    W1, b1 = _mk((2 * F, K), cuda, 0.1, 38), _mk((2 * F,), cuda, 0.5, 39)
    G = torch.empty(M, F, dtype=torch.half, device=cuda)
    ops.gemm(A, W1, G, M, 2 * F, K, lda=K, ldb=K, ldc=F, epilogue=ops.EPI_GEGLU, bias=b1, F=F, ln=(gamma, beta, 1e-5, None))
    h = (ln16.float() @ W1.float().t() + b1.float()).half().float()
    assert _rel(G, h[:, :F] * Fn.gelu(h[:, F:])) < 3e-3
    A9 = _mk((9, K), cuda, 1.0, 40)
    with pytest.raises(Db1Error):
        ops.gemm(A9, B, torch.empty(9, N, dtype=torch.half, device=cuda), 9, N, K, lda=K, ldb=K, ldc=N,
                 ln=(gamma, beta, 1e-5, None))
