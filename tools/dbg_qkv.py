"""Development: QKV-epilogue GEMM with ragged M against torch."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bdm-db1_b200"))
from db1_sm100 import ops
dev = torch.device("cuda")
d = int(sys.argv[1]) if len(sys.argv) > 1 else 128
g = torch.Generator(device="cuda").manual_seed(0)
W = (torch.randn(3 * d, d, generator=g, device=dev) * 0.05).half()
u = (torch.randn(d, generator=g, device=dev) * 0.1).half()
v = (torch.randn(d, generator=g, device=dev) * 0.1).half()
for M in (1, 2, 7, 128, 129, 130, 131, 136, 146, 255, 257, 258, 386):
    x = torch.randn(M, d, generator=g, device=dev).half()
    out = torch.full((M, 4 * d), 9.0, dtype=torch.half, device=dev)
    ops.gemm(x, W, out, M, 3 * d, d, lda=d, ldb=d, ldc=4 * d, epilogue=ops.EPI_QKV, u=u, v=v, d_model=d)
    torch.cuda.synchronize()
    y = x.float() @ W.float().t()
    ref = torch.cat([y[:, :d] + u.float(), y[:, :d] + v.float(), y[:, d:]], 1)
    err = (out.float() - ref).abs().max(dim=1).values
    bad = (err > 2e-2).nonzero().flatten().tolist()
    print("M=%d max err %.3e bad rows %s" % (M, err.max().item(), bad[:10]))
    out2 = torch.full((M, d), 9.0, dtype=torch.half, device=dev)
    ops.gemm(x, W[:d].contiguous(), out2, M, d, d, lda=d, ldb=d, ldc=d)
    torch.cuda.synchronize()
    e2 = (out2.float() - y[:, :d]).abs().max(dim=1).values
    print("      plain max err %.3e bad rows %s" % (e2.max().item(), (e2 > 2e-2).nonzero().flatten().tolist()[:10]))
