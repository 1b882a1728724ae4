"""Per-shape TFLOP/s of db1_gemm_f16 on the GEMM shapes of one DB1-1.3B step (B*L = 4096), with torch.matmul (cuBLAS)
beside it as a yardstick. Development tool: python tools/bench_gemm.py [filter]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bdm-db1_b200"))
from db1_sm100 import ops  # noqa: E402

dev = torch.device("cuda")
R, d, F, V, L = 4096, 2048, 4096, 33025, 1024
Vp = 33152


def t(*shape):
    return (torch.randn(*shape, device=dev) * 0.05).half()


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


cases = []


def case(name, M, N, K, a_mn=False, b_mn=False, **kw):
    cases.append((name, M, N, K, a_mn, b_mn, kw))


case("qkv      fwd", R, 3 * d, d, epilogue=ops.EPI_QKV)
case("r_net    fwd", L, d, d)
case("o_net    fwd", R, d, d, resid=True, drop_p=0.1)
case("ff1geglu fwd", R, 2 * F, d, epilogue=ops.EPI_GEGLU)
case("ff2      fwd", R, d, F, bias=True, resid=True, drop_p=0.1)
case("head     fwd", R, V, d)
case("dWo    wgrad", d, d, R, True, True)
case("do     dgrad", R, d, d, False, True)
case("dWqkv  wgrad", 3 * d, d, R, True, True)
case("dx_qkv dgrad", R, d, 3 * d, False, True, resid=True)
case("dW2    wgrad", d, F, R, True, True)
case("dH    dgeglu", R, F, d, False, True, epilogue=ops.EPI_DGEGLU)
case("dW1    wgrad", 2 * F, d, R, True, True)
case("dx_ff  dgrad", R, d, 2 * F, False, True, resid=True)
case("dWemb  wgrad", V, d, R, True, True, lda=Vp)
case("dh_head dgrad", R, d, V, False, True, lda=Vp)

flt = sys.argv[1] if len(sys.argv) > 1 else ""
tot_us = 0.0
for name, M, N, K, a_mn, b_mn, kw in cases:
    if flt and flt not in name:
        continue
    lda = kw.pop("lda", None)
    A = t(K, lda or M) if a_mn else t(M, lda or K)
    B = t(K, N) if b_mn else t(N, K)
    epi = kw.get("epilogue", ops.EPI_PLAIN)
    ldc = N
    extra = {}
    if epi == ops.EPI_QKV:
        ldc = 4 * d
        extra = dict(u=t(d), v=t(d), d_model=d)
    elif epi == ops.EPI_GEGLU:
        ldc = F
        extra = dict(bias=t(N), H=torch.empty(M, N, dtype=torch.half, device=dev), ldh=N, F=F)
    elif epi == ops.EPI_DGEGLU:
        ldc = 2 * F
        extra = dict(H=t(M, 2 * F), ldh=2 * F, F=F)
    else:
        if kw.get("bias"):
            extra["bias"] = t(N)
        if kw.get("resid"):
            extra["resid"] = t(M, N)
            extra["ldr"] = N
        if kw.get("drop_p"):
            extra["drop_p"] = kw["drop_p"]
            extra["seed"] = 1
        if N % 8:
            ldc = (N + 127) // 128 * 128
    Cc = torch.empty(M, ldc, dtype=torch.half, device=dev)
    fn = lambda: ops.gemm(A, B, Cc, M, N, K, lda=A.stride(0), ldb=B.stride(0), ldc=ldc, a_mn=a_mn, b_mn=b_mn,  # noqa: E731
                          epilogue=epi, **extra)
    us = timeit(fn)
    Am = (A[:, :M].t() if a_mn else A[:, :K])
    Bm = (B if b_mn else B.t())
    us_ref = timeit(lambda: torch.matmul(Am, Bm))
    fl = 2.0 * M * N * K
    tot_us += us
    print("%-14s M=%5d N=%5d K=%5d  ours %8.1f us %7.1f TF | cuBLAS %8.1f us %7.1f TF" %
          (name, M, N, K, us, fl / us / 1e6, us_ref, fl / us_ref / 1e6))
print("sum ours: %.1f us (x24 layers for the per-layer shapes)" % tot_us)
