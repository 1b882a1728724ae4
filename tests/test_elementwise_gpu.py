"""GPU parity of the HBM-bound kernels (LayerNorm fwd/bwd, column sums, masked CE, embedding assembly) against fp32
torch restatements of the reference ops (nn.LayerNorm transformer_xl.py:238/:290, CrossEntropyLoss :602-613,
embedding assembly :621-649). Floating-point kernels: tolerance = fp16 rounding of the outputs (stated per test)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(shape, dev, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).half().to(dev)


def _rel(a, b):
    return (a.float() - b.float()).abs().max().item() / (b.float().abs().max().item() + 1e-12)


@pytest.mark.parametrize("rows,d", [(4096, 2048), (1000, 2048), (515, 1024), (64, 256), (300, 4096), (37, 128), (50, 520)])
@pytest.mark.parametrize("with_bias", [False, True])
def test_layernorm_fwd_bwd(cuda, rows, d, with_bias):
    from db1_sm100 import ops
    y = _mk((rows, d), cuda, 1.5, 1)
    gamma = (1.0 + 0.1 * _mk((d,), cuda, 1.0, 2).float()).half()
    beta = _mk((d,), cuda, 0.1, 3)
    dout = _mk((rows, d), cuda, 1.0, 4)
    out = torch.empty_like(y)
    stats = torch.empty(rows, 2, dtype=torch.float32, device=cuda)
    ops.layernorm_fwd(y, gamma, beta, out, stats, 1e-5)
    yr = y.float().requires_grad_(True)
    gr = gamma.float().requires_grad_(True)
    br = beta.float().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(yr, (d,), gr, br, 1e-5)
    assert _rel(out, ref) < 1e-3  # one fp16 rounding of an O(1) output
    ref.backward(dout.float())
    dy = torch.empty_like(y)
    small = torch.zeros(3 * d, dtype=torch.float32, device=cuda)
    dg, db, dbias = small[0:d], small[d:2 * d], small[2 * d:]
    ops.layernorm_bwd(dout, y, gamma, stats, dy, None, dg, db, dbias if with_bias else None)
    assert _rel(dy, yr.grad) < 2e-3
    assert _rel(dg, gr.grad) < 1e-3
    assert _rel(db, br.grad) < 1e-3
    if with_bias:
        assert _rel(dbias, dy.float().sum(0)) < 1e-3  # bias gradient of the producing GEMM = column sum of (fp16) dy
    else:
        assert dbias.abs().max().item() == 0


@pytest.mark.parametrize("rows,d", [(512, 2048), (100, 256), (33, 128)])
def test_layernorm_bwd_replays_the_gemm_dropout_mask(cuda, rows, d):
    """dz must be dy times exactly the mask the GEMM epilogue applied (same seed, same flat element index)."""
    from db1_sm100 import ops
    p, seed = 0.25, 987654321
    A = _mk((rows, 64), cuda, 1.0, 5)
    Bm = _mk((d, 64), cuda, 1.0, 6)
    Cd = torch.empty(rows, d, dtype=torch.half, device=cuda)
    Cn = torch.empty_like(Cd)
    ops.gemm(A, Bm, Cd, rows, d, 64, lda=64, ldb=64, ldc=d, drop_p=p, seed=seed)
    ops.gemm(A, Bm, Cn, rows, d, 64, lda=64, ldb=64, ldc=d)
    keep = (Cd != 0) | (Cn == 0)
    y = _mk((rows, d), cuda, 1.0, 7)
    gamma = torch.ones(d, dtype=torch.half, device=cuda)
    beta = torch.zeros(d, dtype=torch.half, device=cuda)
    out = torch.empty_like(y)
    stats = torch.empty(rows, 2, dtype=torch.float32, device=cuda)
    ops.layernorm_fwd(y, gamma, beta, out, stats, 1e-5)
    dout = _mk((rows, d), cuda, 1.0, 8)
    dy = torch.empty_like(y)
    dz = torch.empty_like(y)
    small = torch.zeros(3 * d, dtype=torch.float32, device=cuda)
    ops.layernorm_bwd(dout, y, gamma, stats, dy, dz, small[0:d], small[d:2 * d], small[2 * d:], p, seed)
    expect = torch.where(keep, dy.float() / (1 - p), torch.zeros((), device=cuda))
    assert _rel(dz, expect) < 1e-3
    assert ((dz == 0) | keep).all()
    assert _rel(small[2 * d:], dz.float().sum(0)) < 1e-3


@pytest.mark.parametrize("rows,n,ld", [(4096, 8192, 8192), (777, 264, 272), (5, 8, 8), (1030, 2048, 4096)])
def test_colsum(cuda, rows, n, ld):
    from db1_sm100 import ops
    x = _mk((rows, ld), cuda, 1.0, 9)
    out = torch.full((n,), 2.0, dtype=torch.float32, device=cuda)
    ops.colsum(x, out, rows, n)
    ref = x[:, :n].float().sum(0) + 2.0
    assert (out - ref).abs().max().item() <= 1e-3 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("rows,V", [(64, 33025), (17, 1000), (300, 515)])
def test_masked_cross_entropy_fwd_bwd(cuda, rows, V):
    from db1_sm100 import ops
    Vp = (V + 127) // 128 * 128
    g = torch.Generator().manual_seed(11)
    logits = torch.zeros(rows, Vp, dtype=torch.half, device=cuda)
    logits[:, :V] = (torch.randn(rows, V, generator=g) * 3).half().to(cuda)
    logits[:, V:] = 77.0  # padding columns must never be read as logits
    labels = torch.randint(0, V, (rows,), generator=g).to(cuda)
    mask = (torch.rand(rows, generator=g) < 0.4).float().to(cuda)
    mask[0] = 1.0
    row_loss = torch.empty(rows, dtype=torch.float32, device=cuda)
    row_lse = torch.empty_like(row_loss)
    loss2 = torch.empty(2, dtype=torch.float32, device=cuda)
    ops.ce_fwd(logits, labels, mask, row_loss, row_lse, loss2, V)
    lr = logits[:, :V].float().requires_grad_(True)
    ce = torch.nn.functional.cross_entropy(lr, labels, reduction="none")
    ref = (ce * mask).sum() / mask.sum()
    assert abs(loss2[0].item() - ref.item()) <= 1e-5 * abs(ref.item()) + 1e-6  # fp32 reduction of the same fp16 logits
    assert loss2[1].item() == mask.sum().item()
    (ref * 512.0).backward()
    dl = torch.empty(rows, Vp, dtype=torch.half, device=cuda)
    gs = torch.full((1,), 512.0, dtype=torch.float32, device=cuda)
    ops.ce_bwd(logits, labels, mask, row_lse, loss2, gs, dl, V)
    assert _rel(dl[:, :V], lr.grad) < 1e-3
    assert dl[:, V:(V + 7) // 8 * 8].abs().max().item() == 0 if V % 8 else True


def test_cross_entropy_out_of_range_labels_are_ignored_rows(cuda):
    """Labels outside [0, V) (ignore_index -100, image-slot -1, garbage padding) are never dereferenced: the row gives
    zero loss and zero gradient like nn.CrossEntropyLoss's ignore_index, the denominator stays sum(mask)
    (transformer_xl.py:602-609); the loss stays finite even when such a row is masked in."""
    from db1_sm100 import ops
    rows, V = 64, 1000
    Vp = (V + 127) // 128 * 128
    g = torch.Generator().manual_seed(21)
    logits = torch.zeros(rows, Vp, dtype=torch.half, device=cuda)
    logits[:, :V] = (torch.randn(rows, V, generator=g) * 3).half().to(cuda)
    labels = torch.randint(0, V, (rows,), generator=g)
    labels[3], labels[10], labels[11], labels[40] = -100, -1, V, 1 << 40
    labels = labels.to(cuda)
    mask = torch.ones(rows, dtype=torch.float32, device=cuda)
    mask[10] = 0.0
    row_loss = torch.empty(rows, dtype=torch.float32, device=cuda)
    row_lse = torch.empty_like(row_loss)
    loss2 = torch.empty(2, dtype=torch.float32, device=cuda)
    ops.ce_fwd(logits, labels, mask, row_loss, row_lse, loss2, V)
    bad = torch.tensor([3, 10, 11, 40], device=cuda)
    safe = labels.clone()
    safe[bad] = -100
    lr = logits[:, :V].float().requires_grad_(True)
    ce = torch.nn.functional.cross_entropy(lr, safe, reduction="none", ignore_index=-100)
    ref = (ce * mask).sum() / mask.sum()
    assert torch.isfinite(loss2).all()
    assert abs(loss2[0].item() - ref.item()) <= 1e-5 * abs(ref.item())
    assert row_loss[bad].abs().max().item() == 0.0
    ref.backward()
    dl = torch.empty(rows, Vp, dtype=torch.half, device=cuda)
    gs = torch.ones(1, dtype=torch.float32, device=cuda)
    ops.ce_bwd(logits, labels, mask, row_lse, loss2, gs, dl, V)
    assert dl[bad][:, :V].abs().max().item() == 0.0  # (columns >= V are padding, never written)
    assert _rel(dl[:, :V], lr.grad) < 1e-3


def test_embedding_assembly_fwd_bwd(cuda):
    from db1_sm100 import ops
    B, L, d, V, nT, nvis = 3, 40, 256, 500, 33, 6
    g = torch.Generator().manual_seed(12)
    tok = torch.randint(0, V, (B, L), generator=g)
    for b in range(B):
        tok[b, torch.randperm(L, generator=g)[:nvis]] = -1
    pos = torch.randint(0, nT, (B, L), generator=g)
    W, T = _mk((V, d), cuda, 1.0, 13), _mk((nT, d), cuda, 1.0, 14)
    vis = _mk((B, nvis, d), cuda, 1.0, 15)
    tok_d, pos_d = tok.to(cuda), pos.to(cuda)
    out = torch.empty(B, L, d, dtype=torch.half, device=cuda)
    slot = torch.empty(B, L, dtype=torch.int32, device=cuda)
    ops.embed_fwd(tok_d, pos_d, slot, W, T, vis, out, L * d, B, L, d, V)
    ref = torch.zeros(B, L, d, device=cuda)
    for b in range(B):
        k = 0
        for l in range(L):
            if tok[b, l] >= 0:
                ref[b, l] = W[tok[b, l]].float()
            else:
                ref[b, l] = vis[b, k].float()
                assert slot[b, l].item() == k  # integer bookkeeping: bit-exact
                k += 1
            ref[b, l] += T[pos[b, l]].float()
    assert torch.equal(out, ref.half())
    dout = _mk((B, L, d), cuda, 1.0, 16)
    dW = torch.zeros(V, d, dtype=torch.half, device=cuda)
    dT = torch.zeros(nT, d, dtype=torch.half, device=cuda)
    dvis = torch.zeros_like(vis)
    ops.embed_bwd(tok_d, pos_d, slot, dout, L * d, dW, dT, dvis, B, L, d, V)
    rW = torch.zeros(V, d, device=cuda)
    rT = torch.zeros(nT, d, device=cuda)
    flat = dout.float().reshape(-1, d)
    tk = tok_d.reshape(-1)
    rW.index_add_(0, tk.clamp(min=0), flat * (tk >= 0).float()[:, None])
    rT.index_add_(0, pos_d.reshape(-1), flat)
    assert _rel(dW, rW) < 4e-3 and _rel(dT, rT) < 1e-2  # fp16 atomic accumulation order
    for b in range(B):
        assert torch.equal(dvis[b], dout[b][tok_d[b] == -1])
