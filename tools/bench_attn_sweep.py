"""Attention sweep with the backward included (BASELINE.json config 5): B=4, H=16, dh=128, L in {512, 1024, 2048, 4096}.
Per L: the fused forward (db1_relattn_fwd) and the four backward kernels (recompute P/dS, key-outer dK/dV, query-outer dq,
diagonal-outer dR), CUDA-event timed one kernel at a time, as TFLOP/s over the unmasked pairs and as algorithmic GB/s.
    python tools/bench_attn_sweep.py [L ...]
Writes gpurun_out/attn_sweep.json when that directory exists."""
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bdm-db1_b200"))
from db1_sm100 import ops  # noqa: E402

dev = torch.device("cuda")
B, H, dh = 4, 16, 128
d = H * dh


def timeit(fn, iters=10):
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


Ls = [int(a) for a in sys.argv[1:]] or [512, 1024, 2048, 4096]
rows = []
for L in Ls:
    window = L
    scale = 1.0 / math.sqrt(dh)
    qkv4 = (torch.randn(B * L, 4 * d, device=dev) * 0.5).half()
    r = (torch.randn(L, d, device=dev) * 0.5).half()
    do = (torch.randn(B * L, d, device=dev) * 0.1).half()
    o = torch.empty(B * L, d, dtype=torch.half, device=dev)
    lse2 = torch.empty(B, H, L, dtype=torch.float32, device=dev)
    pairs = L * (L + 1) / 2
    unit = B * H * 2.0 * dh * pairs  # one [L x L] x dh causal contraction
    io = B * L * d * 2.0             # one [B*L, d] fp16 operand
    res = dict(L=L)

    def rec(name, us, n_contractions, nbytes):
        res[name] = dict(us=us, tflops=n_contractions * unit / us / 1e6, gbs=nbytes / us / 1e3)
        print("L=%5d %-12s %9.1f us  %7.1f TFLOP/s  %7.1f GB/s (algorithmic)" % (L, name, us, res[name]["tflops"], res[name]["gbs"]))

    rec("fwd", timeit(lambda: ops.relattn_fwd(qkv4, r, o, lse2, B, L, H, dh, window, scale)), 3, 5 * io + L * d * 2.0)
    tshape = ops.score_tiles_shape(B, L, H)
    P = torch.empty(tshape, dtype=torch.half, device=dev)
    dS = torch.empty(tshape, dtype=torch.half, device=dev)
    tile_bytes = 2.0 * B * H * (L // 128) * (L // 128 + 1) / 2 * 128 * 128
    dqkv = torch.empty(B * L, 3 * d, dtype=torch.half, device=dev)
    du = torch.zeros(d, dtype=torch.float32, device=dev)
    dv = torch.zeros(d, dtype=torch.float32, device=dev)
    dr32 = torch.zeros(L, d, dtype=torch.float32, device=dev)
    rec("bwd_ds", timeit(lambda: ops.relattn_bwd_ds_tiled(qkv4, r, do, lse2, None, P, dS, B, L, H, dh, window, scale, o=o)),
        3, 6 * io + 2 * tile_bytes)
    rec("bwd_dkdv", timeit(lambda: ops.relattn_bwd_dkdv(P, dS, do, qkv4[:, 0:d], dqkv[:, d:2 * d], dqkv[:, 2 * d:], B, L, H,
                                                        dh, window)), 2, 4 * io + 2 * tile_bytes)
    rec("bwd_dq", timeit(lambda: ops.relattn_bwd_dq(dS, qkv4[:, 2 * d:3 * d], r, dqkv[:, 0:d], du, dv, B, L, H, dh, window)),
        2, 2 * io + tile_bytes)
    rec("bwd_dr", timeit(lambda: ops.relattn_bwd_dr(dS, qkv4[:, d:2 * d], dr32, B, L, H, dh, window)), 1, io + tile_bytes)
    tot = sum(res[k]["us"] for k in ("bwd_ds", "bwd_dkdv", "bwd_dq", "bwd_dr"))
    res["bwd_total"] = dict(us=tot, tflops=8 * unit / tot / 1e6)
    print("L=%5d %-12s %9.1f us  %7.1f TFLOP/s" % (L, "bwd total", tot, res["bwd_total"]["tflops"]))
    rows.append(res)
    del P, dS
out = os.path.join(ROOT, "gpurun_out")
if os.path.isdir(out):
    with open(os.path.join(out, "attn_sweep.json"), "w") as f:
        json.dump(rows, f, indent=1)
