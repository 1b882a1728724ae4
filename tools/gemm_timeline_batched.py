"""Development: per-CTA timeline (DB1_GEMM_DBG=8) of one causal batched attention-backward GEMM (dQu = dS . K)."""
import os
import sys

os.environ["DB1_GEMM_DBG"] = "8"
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bdm-db1_b200"))
from db1_sm100 import ops  # noqa: E402

dev = torch.device("cuda")
B, H, dh, L = 4, 16, 128, 1024
d = H * dh
which = sys.argv[1] if len(sys.argv) > 1 else "dqu"
dS = torch.randn(B, H, L, L, device=dev).half().tril_()
P2 = torch.randn(B, H, L, L, device=dev).half().tril_()
qkv4 = torch.randn(B * L, 4 * d, device=dev).half()
do = torch.randn(B * L, d, device=dev).half()
kk = qkv4[:, 2 * d:3 * d]
qu = qkv4[:, 0:d]
out = torch.empty(B * L, d, dtype=torch.half, device=dev)
LL = L * L
sz = (LL, H * LL)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run():
    if which == "dqu":
        ops.gemm(dS, kk, out, L, dh, L, lda=L, ldb=4 * d, ldc=d, b_mn=True, Z1=H, Z2=B, a_z=sz, b_z=(dh, L * 4 * d),
                 c_z=(dh, L * d), k_mode=ops.K_END_BY_ROW)
    else:  # dV = P^T dO
        ops.gemm(P2, do, out, L, dh, L, lda=L, ldb=d, ldc=d, a_mn=True, b_mn=True, Z1=H, Z2=B, a_z=sz, b_z=(dh, L * d),
                 c_z=(dh, L * d), k_mode=ops.K_BEGIN_BY_ROW)


for cold in (0, 1):
    for _ in range(2):
        if cold:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
    print("%s %s: %.1f us" % (which, "cold (L2 flushed)" if cold else "warm", e0.elapsed_time(e1) * 1e3))
    ws = ops._gemm_ws[0]
    tl = ws[4096:4096 + 148 * 64 * 8].view(torch.int64).view(148, 64).cpu()
    for c in (0, 1, 73, 74, 146, 147):
        t0 = tl[c, 0].item()
        row = ["%3d %7d |" % (c, tl[c, 1].item() - t0)]
        for i in range(5):
            v = [tl[c, 4 + 4 * i + j].item() for j in range(4)]
            if v[0] == 0:
                break
            row.append("[%6d %6d | %6d %6d]" % tuple(x - t0 if x else -1 for x in v))
        print(" ".join(row))
    tot = (tl[:, 1] - tl[:, 0]).float()
    print("kernel span per CTA: mean %.0f min %.0f max %.0f clk" % (tot.mean().item(), tot.min().item(), tot.max().item()))
    ws[4096:4096 + 148 * 64 * 8].zero_()
