"""Development: per-CTA clock64 timeline of one GEMM launch (DB1_GEMM_DBG=8). python tools/gemm_timeline.py M N K [resid] [drop]"""
import os
import sys

os.environ["DB1_GEMM_DBG"] = str(8 | int(os.environ.get("DBG_EXTRA", "0")))
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bdm-db1_b200"))
from db1_sm100 import ops  # noqa: E402

dev = torch.device("cuda")
M, N, K = (int(x) for x in sys.argv[1:4])
resid = "resid" in sys.argv
drop = 0.1 if "drop" in sys.argv else 0.0
bmn = "bmn" in sys.argv
dgeglu = "dgeglu" in sys.argv
geglu = "geglu" in sys.argv and not dgeglu
A = (torch.randn(M, K, device=dev) * 0.05).half()
B = (torch.randn(K, N, device=dev) * 0.05).half() if bmn else (torch.randn(N, K, device=dev) * 0.05).half()
Cc = torch.empty(M, N, dtype=torch.half, device=dev)
R = torch.randn(M, N, device=dev).half() if resid else None
if dgeglu:
    Hs = torch.randn(M, 2 * N, device=dev).half()
    Cc = torch.empty(M, 2 * N, dtype=torch.half, device=dev)
if geglu:  # N = 2F
    Hs = torch.empty(M, N, dtype=torch.half, device=dev)
    Cc = torch.empty(M, N // 2, dtype=torch.half, device=dev)
    bias = torch.randn(N, device=dev).half()
for _ in range(3):
    if geglu:
        ops.gemm(A, B, Cc, M, N, K, lda=K, ldb=K, ldc=N // 2, epilogue=ops.EPI_GEGLU, bias=bias, H=Hs, ldh=N, F=N // 2)
    elif dgeglu:
        ops.gemm(A, B, Cc, M, N, K, lda=K, ldb=B.stride(0), ldc=2 * N, b_mn=bmn, epilogue=ops.EPI_DGEGLU, H=Hs, ldh=2 * N, F=N)
    else:
        ops.gemm(A, B, Cc, M, N, K, lda=K, ldb=B.stride(0), ldc=N, b_mn=bmn, resid=R, ldr=N if resid else 0, drop_p=drop, seed=5)
torch.cuda.synchronize()
ws = ops._gemm_ws[0]
tl = ws[4096:4096 + 148 * 64 * 8].view(torch.int64).view(148, 64).cpu()
print("cta  total | per item: mma_first_full->issue_end  tfull_seen  epi_end (clk, relative to CTA start)")
for c in list(range(0, 8)) + list(range(140, 148)):
    t0 = tl[c, 0].item()
    row = ["%3d %7d |" % (c, tl[c, 1].item() - t0)]
    for i in range(5):
        v = [tl[c, 4 + 4 * i + j].item() for j in range(4)]
        if v[0] == 0:
            break
        row.append("[%6d %6d | %6d %6d]" % tuple(x - t0 if x else -1 for x in v))
    print(" ".join(row))
tot = (tl[:, 1] - tl[:, 0]).float()
print("kernel span per CTA: mean %.0f max %.0f clk" % (tot.mean().item(), tot.max().item()))
