import math, os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "bdm-db1_b200"))
from db1_sm100 import ops
dev = torch.device("cuda")
B, mlen, qlen, H, dh = 2, 64, 1, 4, 32
d = H * dh; K = mlen + qlen
g = torch.Generator(device="cuda").manual_seed(3)
Wqkv = (torch.randn(3 * d, d, generator=g, device=dev) * 0.05).half()
Wr = (torch.randn(d, d, generator=g, device=dev) * 0.05).half()
u = (torch.randn(d, generator=g, device=dev) * 0.1).half(); v = (torch.randn(d, generator=g, device=dev) * 0.1).half()
def rel(a, b): return ((a.float() - b.float()).abs().max() / b.float().abs().max()).item()
for rep in range(4):
    mem = torch.randn(B, mlen, d, generator=g, device=dev).half()
    w = torch.randn(B, qlen, d, generator=g, device=dev).half()
    r = torch.randn(K, d, generator=g, device=dev).half()
    # allocator churn like the model's
    junk = [torch.empty(1024 * (i + 1), device=dev) for i in range(8)]; del junk
    cat = torch.cat([mem, w], dim=1).reshape(B * K, d)
    qkv4 = torch.empty(B * K, 4 * d, dtype=torch.half, device=dev)
    ops.gemm(cat, Wqkv, qkv4, B * K, 3 * d, d, lda=d, ldb=d, ldc=4 * d, epilogue=ops.EPI_QKV, u=u, v=v, d_model=d)
    rk = torch.empty(K, d, dtype=torch.half, device=dev)
    ops.gemm(r, Wr, rk, K, d, d, lda=d, ldb=d, ldc=d)
    o = torch.empty(B, K, d, dtype=torch.half, device=dev)
    lse2 = torch.empty(B, H, K, dtype=torch.float32, device=dev)
    ops.relattn_mem_fwd(qkv4, rk, o.view(B * K, d), lse2, B, K, H, dh, 64, 1 / math.sqrt(dh), mlen)
    oq = o[:, mlen:].reshape(B * qlen, d)
    torch.cuda.synchronize()
    y = cat.float() @ Wqkv.float().t()
    ref4 = torch.cat([y[:, :d] + u.float(), y[:, :d] + v.float(), y[:, d:]], 1)
    rkr = r.float() @ Wr.float().t()
    x = qkv4.float().view(B, K, 4, H, dh); R = rk.float().view(K, H, dh)
    i = torch.arange(K, device=dev)[:, None]; j = torch.arange(K, device=dev)[None, :]
    ok = (j <= i) & (i - j < 64)
    ac = torch.einsum("bihd,bjhd->bhij", x[:, :, 0], x[:, :, 2])
    idx = (j + K - 1 - i).clamp(0, K - 1)
    bd = torch.einsum("bihd,chd->bhic", x[:, :, 1], R).gather(-1, idx[None, None].expand(B, H, K, K))
    s = torch.where(ok[None, None], (ac + bd) / math.sqrt(dh), torch.full((), -1e30, device=dev))
    ref = torch.einsum("bhij,bjhd->bihd", torch.softmax(s, -1), x[:, :, 3]).reshape(B, K, d)
    print(rep, "qkv %.1e rk %.1e attn(b0) %.1e attn(b1) %.1e" % (rel(qkv4, ref4), rel(rk, rkr), rel(oq[0], ref[0, mlen]), rel(oq[1], ref[1, mlen])))
