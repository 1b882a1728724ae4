"""Development: db1_relattn_mem_fwd for small / odd K against a dense torch reference."""
import math, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bdm-db1_b200"))
from db1_sm100 import ops
dev = torch.device("cuda")
H, dh = 4, 32
d = H * dh
for (B, K, mlen, window) in [(1, 65, 64, 64), (2, 65, 64, 64), (2, 66, 64, 64), (2, 73, 64, 64), (3, 65, 64, 64), (2, 65, 64, 1 << 20),
                             (2, 129, 128, 128), (2, 64, 63, 64), (2, 72, 71, 64)]:
    g = torch.Generator(device="cuda").manual_seed(K * 7 + B)
    qkv4 = torch.randn(B * K, 4 * d, generator=g, device=dev).half()
    r = torch.randn(K, d, generator=g, device=dev).half()
    o = torch.zeros(B, K, d, dtype=torch.half, device=dev)
    lse2 = torch.zeros(B, H, K, dtype=torch.float32, device=dev)
    scale = 1 / math.sqrt(dh)
    ops.relattn_mem_fwd(qkv4, r, o.view(B * K, d), lse2, B, K, H, dh, window, scale, mlen)
    torch.cuda.synchronize()
    x = qkv4.float().view(B, K, 4, H, dh)
    qu, qv, k, v = x[:, :, 0], x[:, :, 1], x[:, :, 2], x[:, :, 3]
    R = r.float().view(K, H, dh)
    i = torch.arange(K, device=dev)[:, None]
    j = torch.arange(K, device=dev)[None, :]
    ok = (j <= i) & (i - j < window)
    ac = torch.einsum("bihd,bjhd->bhij", qu, k)
    idx = (j + K - 1 - i).clamp(0, K - 1)
    bd = torch.einsum("bihd,chd->bhic", qv, R).gather(-1, idx[None, None].expand(B, H, K, K))
    s = torch.where(ok[None, None], (ac + bd) * scale, torch.full((), -1e30, device=dev))
    ref = torch.einsum("bhij,bjhd->bihd", torch.softmax(s, -1), v).reshape(B, K, d)
    err = (o.float() - ref)[:, mlen:].abs().amax(dim=(1, 2)) / ref.abs().max()
    print("B=%d K=%d mlen=%d window=%d: per-sequence rel err %s" % (B, K, mlen, window, ["%.2e" % e for e in err.tolist()]))
