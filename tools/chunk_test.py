"""Development: what bounds the causal batched attention-backward GEMMs? Variants of dQu = dS . K at B=4, H=16, L=1024."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bdm-db1_b200"))
from db1_sm100 import ops
dev = torch.device("cuda")
B, H, dh, L = 4, 16, 128, 1024
d = H * dh
dS = torch.randn(B, H, L, L, device=dev).half().tril_()
qkv4 = torch.randn(B * L, 4 * d, device=dev).half()
kk = qkv4[:, 2 * d:3 * d]
kc = kk.reshape(B, L, H, dh).permute(0, 2, 1, 3).contiguous()  # [B, H, L, dh] compact per head
out = torch.empty(B * L, d, dtype=torch.half, device=dev)
LL = L * L; sz = (LL, H * LL)

def t(name, fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    print("%-46s %.1f us" % (name, e0.elapsed_time(e1) / n * 1e3))

t("causal, K strided in qkv4 (model layout)", lambda: ops.gemm(dS, kk, out, L, dh, L, lda=L, ldb=4 * d, ldc=d, b_mn=True, Z1=H, Z2=B, a_z=sz, b_z=(dh, L * 4 * d), c_z=(dh, L * d), k_mode=ops.K_END_BY_ROW))
t("full K range, K strided", lambda: ops.gemm(dS, kk, out, L, dh, L, lda=L, ldb=4 * d, ldc=d, b_mn=True, Z1=H, Z2=B, a_z=sz, b_z=(dh, L * 4 * d), c_z=(dh, L * d)))
t("causal, K compact [B,H,L,dh]", lambda: ops.gemm(dS, kc, out, L, dh, L, lda=L, ldb=dh, ldc=d, b_mn=True, Z1=H, Z2=B, a_z=sz, b_z=(L * dh, H * L * dh), c_z=(dh, L * d), k_mode=ops.K_END_BY_ROW))
t("full, K compact", lambda: ops.gemm(dS, kc, out, L, dh, L, lda=L, ldb=dh, ldc=d, b_mn=True, Z1=H, Z2=B, a_z=sz, b_z=(L * dh, H * L * dh), c_z=(dh, L * d)))
os.environ["DB1_GEMM_NO_SNAKE"] = "1"
t("causal, K strided, no snake order", lambda: ops.gemm(dS, kk, out, L, dh, L, lda=L, ldb=4 * d, ldc=d, b_mn=True, Z1=H, Z2=B, a_z=sz, b_z=(dh, L * 4 * d), c_z=(dh, L * d), k_mode=ops.K_END_BY_ROW))
