"""GPU parity of the fused relative-position attention forward against the CPU oracle (oracle/db1_oracle.py)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return (a.float() - b.float()).abs().max().item() / (b.float().abs().max().item() + 1e-12)


@pytest.mark.parametrize("B,L,H,dh,window", [(1, 128, 1, 128, 1 << 20), (2, 256, 2, 128, 1 << 20),
                                             (1, 512, 2, 128, 200), (2, 256, 4, 32, 1 << 20), (1, 200, 2, 64, 77),
                                             (1, 1024, 2, 128, 1024)])
def test_relattn_fwd_matches_oracle(cuda, B, L, H, dh, window):
    from db1_sm100 import ops
    from oracle import db1_oracle as orc
    d = H * dh
    g = torch.Generator().manual_seed(1)
    q = torch.randn(B, L, H, dh, generator=g)
    k = torch.randn(B, L, H, dh, generator=g)
    v = torch.randn(B, L, H, dh, generator=g)
    rk = torch.randn(L, H, dh, generator=g)
    u = torch.randn(H, dh, generator=g) * 0.5
    vb = torch.randn(H, dh, generator=g) * 0.5
    qu = (q.half() + u.half()).half()
    qv = (q.half() + vb.half()).half()
    qkv4 = torch.cat([qu.reshape(B * L, d), qv.reshape(B * L, d), k.half().reshape(B * L, d),
                      v.half().reshape(B * L, d)], 1).contiguous().to(cuda)
    r = rk.half().reshape(L, d).contiguous().to(cuda)
    out = torch.zeros(B * L, d, dtype=torch.half, device=cuda)
    lse2 = torch.zeros(B, H, L, dtype=torch.float32, device=cuda)
    scale = 1.0 / math.sqrt(dh)
    ops.relattn_fwd(qkv4, r, out, lse2, B, L, H, dh, window, scale)
    torch.cuda.synchronize()
    # oracle on the same fp16-rounded operands (u, v already folded into qu / qv)
    ok = orc.attention_mask_ok(L, L, window, True)
    zero = torch.zeros(H, dh)
    o_ac, _, s_ac = orc.rel_attention_core(qu.float(), k.half().float(), v.half().float(), rk.half().float() * 0, zero, zero, ok, scale)
    # scores = AC(qu) + BD(qv): evaluate the two halves with their own query copies
    _, _, s_bd = orc.rel_attention_core(qv.float(), k.half().float() * 0, v.half().float(), rk.half().float(), zero, zero, ok, scale)
    s = torch.where(ok[None, None], s_ac + s_bd, torch.full((), -1e30))
    p = torch.softmax(s, -1)
    o = torch.einsum("bhij,bjhd->bihd", p, v.half().float()).reshape(B * L, d)
    assert _rel(out.cpu(), o) < 3e-3
    lse_ref = torch.logsumexp(s, -1) / math.log(2.0)
    assert (lse2.cpu() - lse_ref).abs().max().item() < 2e-3
    # recompute mode: normalised probabilities on the causal tiles
    probs = torch.zeros(B, H, L, L, dtype=torch.half, device=cuda)
    ops.relattn_fwd(qkv4, r, None, lse2, B, L, H, dh, window, scale, probs=probs)
    torch.cuda.synchronize()
    assert (probs.cpu().float() - p).abs().max().item() < 2e-3


@pytest.mark.parametrize("B,L,H,dh,window", [(2, 256, 2, 128, 1 << 20), (1, 512, 2, 128, 200), (2, 256, 4, 32, 1 << 20),
                                             (1, 200, 2, 64, 77), (1, 1024, 1, 128, 1024)])
def test_relattn_bwd_ds_matches_autograd(cuda, B, L, H, dh, window):
    """P and dS = dLoss/dS (pre-softmax scores, scale folded in) from the fused recompute kernel vs fp32 autograd of
    the oracle's score/softmax/value restatement on the same fp16-rounded operands."""
    from db1_sm100 import ops
    from oracle import db1_oracle as orc
    d = H * dh
    g = torch.Generator().manual_seed(2)
    qu = torch.randn(B, L, H, dh, generator=g).half()
    qv = torch.randn(B, L, H, dh, generator=g).half()
    k = torch.randn(B, L, H, dh, generator=g).half()
    v = torch.randn(B, L, H, dh, generator=g).half()
    rk = torch.randn(L, H, dh, generator=g).half()
    do = torch.randn(B, L, H, dh, generator=g).half()
    scale = 1.0 / math.sqrt(dh)
    ok = orc.attention_mask_ok(L, L, window, True)
    zero = torch.zeros(H, dh)
    _, _, s_ac = orc.rel_attention_core(qu.float(), k.float(), v.float(), rk.float() * 0, zero, zero, ok, scale)
    _, _, s_bd = orc.rel_attention_core(qv.float(), k.float() * 0, v.float(), rk.float(), zero, zero, ok, scale)
    s = torch.where(ok[None, None], s_ac + s_bd, torch.full((), -1e30)).requires_grad_(True)  # already scaled
    p = torch.softmax(s, -1)
    o = torch.einsum("bhij,bjhd->bihd", p, v.float())
    (o * do.float()).sum().backward()
    ds_ref = s.grad * scale  # the kernel returns dLoss/d(unscaled score) = dLoss/dS * scale
    qkv4 = torch.cat([x.reshape(B * L, d) for x in (qu, qv, k, v)], 1).contiguous().to(cuda)
    r = rk.reshape(L, d).contiguous().to(cuda)
    do_d = do.reshape(B * L, d).contiguous().to(cuda)
    out = torch.zeros(B * L, d, dtype=torch.half, device=cuda)
    lse2 = torch.zeros(B, H, L, dtype=torch.float32, device=cuda)
    ops.relattn_fwd(qkv4, r, out, lse2, B, L, H, dh, window, scale)
    drow = torch.empty(B, H, L, dtype=torch.float32, device=cuda)
    ops.rowdot(do_d, out, drow, B, L, H, dh)
    P = torch.zeros(B, H, L, L, dtype=torch.half, device=cuda)
    dS = torch.zeros(B, H, L, L, dtype=torch.half, device=cuda)
    ops.relattn_bwd_ds(qkv4, r, do_d, lse2, drow, P, dS, B, L, H, dh, window, scale)
    torch.cuda.synchronize()
    assert (P.cpu().float() - p.detach()).abs().max().item() < 2e-3
    assert _rel(dS.cpu(), ds_ref) < 5e-3  # D comes from the fp16-rounded O; dS itself is stored in fp16
    # same with D = rowsum(dO * O) formed inside the kernel (the path the model uses): identical inputs -> same dS
    P2 = torch.zeros_like(P)
    dS2 = torch.zeros_like(dS)
    ops.relattn_bwd_ds(qkv4, r, do_d, lse2, None, P2, dS2, B, L, H, dh, window, scale, o=out)
    torch.cuda.synchronize()
    assert torch.equal(P2, P)
    assert _rel(dS2.cpu(), ds_ref) < 5e-3
    assert (dS2.float() - dS.float()).abs().max().item() <= 2e-3 * dS.float().abs().max().item()
    # tiled producer (what the model's backward uses): same values on every visited tile, zeros where masked
    shp = ops.score_tiles_shape(B, L, H)
    Pt = torch.full(shp, float("nan"), dtype=torch.half, device=cuda)
    dSt = torch.full(shp, float("nan"), dtype=torch.half, device=cuda)
    ops.relattn_bwd_ds_tiled(qkv4, r, do_d, lse2, None, Pt, dSt, B, L, H, dh, window, scale, o=out)
    torch.cuda.synchronize()
    nq = (L + 127) // 128
    for I in range(nq):
        jlo = max(0, I * 128 - window + 1) // 128
        for J in range(nq):
            rows = slice(I * 128, min(L, (I + 1) * 128))
            cols = slice(J * 128, min(L, (J + 1) * 128))
            if jlo <= J <= I:
                assert not torch.isnan(Pt[:, :, I, J]).any() and not torch.isnan(dSt[:, :, I, J]).any()
                assert torch.equal(_from_tiles(Pt, L)[:, :, rows, cols], P2[:, :, rows, cols])
                assert torch.equal(_from_tiles(dSt, L)[:, :, rows, cols], dS2[:, :, rows, cols])
            else:
                assert torch.isnan(Pt[:, :, I, J]).all()  # never touched
    # rows / columns beyond the sequence inside visited tiles are zeros
    full = _from_tiles(torch.nan_to_num(Pt, nan=0.0), nq * 128)
    assert full[:, :, L:, :].abs().max().item() == 0 if nq * 128 > L else True


def _to_tiles(x):
    """[B,H,L,L] -> the tiled scratch layout [B,H,nq,nq,2,128,64] (zero padded to whole tiles)."""
    B, H, L, _ = x.shape
    nq = (L + 127) // 128
    xp = torch.zeros(B, H, nq * 128, nq * 128, dtype=x.dtype, device=x.device)
    xp[:, :, :L, :L] = x
    return xp.view(B, H, nq, 128, nq, 2, 64).permute(0, 1, 2, 4, 5, 3, 6).contiguous()


def _from_tiles(t, L):
    B, H, nq = t.shape[:3]
    return t.permute(0, 1, 2, 5, 3, 4, 6).reshape(B, H, nq * 128, nq * 128)[:, :, :L, :L]


def _unshift_ref(ds):
    """dsr[..., i, c] = ds[..., i, c - (L-1-i)] (0 where the source column is negative): adjoint of _rel_shift
    (transformer_xl.py:98-110) for qlen == klen, restated with an index gather."""
    L = ds.shape[-1]
    i = torch.arange(L, device=ds.device)[:, None]
    c = torch.arange(L, device=ds.device)[None, :]
    j = c - (L - 1 - i)
    valid = j >= 0
    g = ds.gather(-1, j.clamp(min=0).expand(ds.shape))
    return torch.where(valid, g, torch.zeros((), device=ds.device, dtype=ds.dtype))


@pytest.mark.parametrize("B,L,H,dh,window", [(1, 128, 1, 128, 1 << 20), (2, 256, 2, 128, 1 << 20), (3, 384, 2, 128, 1 << 20),
                                             (1, 512, 2, 128, 200), (2, 256, 4, 32, 1 << 20), (1, 200, 2, 64, 77),
                                             (2, 1024, 2, 128, 1024), (1, 1024, 1, 128, 300),
                                             (2, 1024, 16, 128, 1024), (3, 640, 24, 64, 1 << 20), (2, 331, 3, 64, 1 << 20)])
def test_relattn_bwd_dq_dr_match_dense_formulas(cuda, B, L, H, dh, window):
    """db1_relattn_bwd_dq / db1_relattn_bwd_dr against the dense fp32 formulas:
    dq = dS K + unshift(dS) R, du = sum dS K, dv = sum unshift(dS) R, dR = sum_b unshift(dS)^T (q+v)."""
    from db1_sm100 import ops
    d = H * dh
    g = torch.Generator().manual_seed(11)
    i = torch.arange(L)[:, None]
    j = torch.arange(L)[None, :]
    ok = (j <= i) & (i - j < window)
    ds = (torch.randn(B, H, L, L, generator=g) * ok).half().to(cuda)
    ds_t = _to_tiles(ds)
    # tiles the recompute kernel never visits hold garbage in production: poison them
    nq = (L + 127) // 128
    unvisited = [(I, J) for I in range(nq) for J in range(nq) if J > I or J < max(0, I * 128 - window + 1) // 128]
    for I, J in unvisited:
        ds_t[:, :, I, J] = float("nan")
    qkv4 = (torch.randn(B * L, 4 * d, generator=g) * 0.7).half().to(cuda)
    r = (torch.randn(L, d, generator=g) * 0.7).half().to(cuda)
    qv = qkv4[:, d:2 * d]
    kk = qkv4[:, 2 * d:3 * d]
    dqkv = torch.full((B * L, 3 * d), 7.0, dtype=torch.half, device=cuda)
    du = torch.zeros(d, dtype=torch.float32, device=cuda)
    dv = torch.zeros(d, dtype=torch.float32, device=cuda)
    dr = torch.zeros(L, d, dtype=torch.float32, device=cuda)
    ops.relattn_bwd_dq(ds_t, kk, r, dqkv[:, 0:d], du, dv, B, L, H, dh, window)
    ops.relattn_bwd_dr(ds_t, qv, dr, B, L, H, dh, window)
    torch.cuda.synchronize()
    dsf = ds.float() * ok.to(cuda)
    dsr = _unshift_ref(dsf)
    K4 = kk.float().reshape(B, L, H, dh)
    Qv4 = qv.float().reshape(B, L, H, dh)
    R3 = r.float().reshape(L, H, dh)
    dqu = torch.einsum("bhij,bjhd->bihd", dsf, K4)
    dqv = torch.einsum("bhic,chd->bihd", dsr, R3)
    dq_ref = (dqu + dqv).reshape(B * L, d)
    assert _rel(dqkv[:, 0:d], dq_ref) < 2e-3
    assert torch.all(dqkv[:, d:] == 7.0)  # neighbours of the strided output view untouched
    assert _rel(du, dqu.sum((0, 1)).reshape(d)) < 1e-3
    assert _rel(dv, dqv.sum((0, 1)).reshape(d)) < 1e-3
    dr_ref = torch.einsum("bhic,bihd->chd", dsr, Qv4).reshape(L, d)
    assert _rel(dr, dr_ref) < 1e-3
    # accumulation contract: a second call adds
    ops.relattn_bwd_dr(ds_t, qv, dr, B, L, H, dh, window)
    torch.cuda.synchronize()
    assert _rel(dr, 2 * dr_ref) < 1e-3
    # key-outer kernel: dv = P^T dO, dk = dS^T (q+u)
    probs = (torch.rand(B, H, L, L, generator=g) * ok).half().to(cuda)
    probs_t = _to_tiles(probs)
    for I, J in unvisited:
        probs_t[:, :, I, J] = float("nan")
    do = (torch.randn(B * L, d, generator=g) * 0.7).half().to(cuda)
    qu = qkv4[:, 0:d]
    dqkv.fill_(7.0)
    ops.relattn_bwd_dkdv(probs_t, ds_t, do, qu, dqkv[:, d:2 * d], dqkv[:, 2 * d:], B, L, H, dh, window)
    torch.cuda.synchronize()
    pf = probs.float() * ok.to(cuda)
    dv_ref = torch.einsum("bhij,bihd->bjhd", pf, do.float().reshape(B, L, H, dh)).reshape(B * L, d)
    dk_ref = torch.einsum("bhij,bihd->bjhd", dsf, qu.float().reshape(B, L, H, dh)).reshape(B * L, d)
    assert _rel(dqkv[:, 2 * d:], dv_ref) < 2e-3
    assert _rel(dqkv[:, d:2 * d], dk_ref) < 2e-3
    assert torch.all(dqkv[:, 0:d] == 7.0)
