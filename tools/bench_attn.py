"""Attention-kernel sweep (BASELINE.json config 5): db1_relattn_fwd (and the backward when built) at B=4, H=16, dh=128,
L in {512, 1024, 2048, 4096}, window in {L, 1024}; prints TFLOP/s (algorithmic, unmasked pairs only) and GB/s.
Development tool: python tools/bench_attn.py [L ...]"""
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
    for window in sorted({L, min(L, 1024)}, reverse=True):
        qkv4 = (torch.randn(B * L, 4 * d, device=dev) * 0.5).half()
        r = (torch.randn(L, d, device=dev) * 0.5).half()
        o = torch.empty(B * L, d, dtype=torch.half, device=dev)
        lse2 = torch.empty(B, H, L, dtype=torch.float32, device=dev)
        us = timeit(lambda: ops.relattn_fwd(qkv4, r, o, lse2, B, L, H, dh, window, 1.0 / math.sqrt(dh)))
        W = min(window, L)
        pairs = W * (W + 1) / 2 + (L - W) * W
        fl = B * H * 6.0 * dh * pairs
        by = B * H * (4.0 * L * dh * 2 + 4 * L) + H * L * dh * 2.0
        row = dict(kernel="relattn_fwd", L=L, window=window, us=us, tflops=fl / us / 1e6, gbs=by / us / 1e3)
        rows.append(row)
        print("relattn_fwd L=%5d window=%5d  %9.1f us  %7.1f TFLOP/s  %7.1f GB/s" % (L, window, us, row["tflops"], row["gbs"]))
        if hasattr(ops, "relattn_bwd"):
            do = (torch.randn(B * L, d, device=dev) * 0.5).half()
            dqkv = torch.empty(B * L, 3 * d, dtype=torch.half, device=dev)
            drk = torch.zeros(L, d, dtype=torch.float32, device=dev)
            duv = torch.zeros(2 * d, dtype=torch.float32, device=dev)
            usb = timeit(lambda: ops.relattn_bwd(qkv4, r, o, do, lse2, dqkv, drk, duv, B, L, H, dh, window,
                                                 1.0 / math.sqrt(dh)))
            flb = B * H * 16.0 * dh * pairs
            rows.append(dict(kernel="relattn_bwd", L=L, window=window, us=usb, tflops=flb / usb / 1e6))
            print("relattn_bwd L=%5d window=%5d  %9.1f us  %7.1f TFLOP/s" % (L, window, usb, flb / usb / 1e6))
out = os.path.join(ROOT, "gpurun_out")
if os.path.isdir(out):
    with open(os.path.join(out, "attn_sweep.json"), "w") as f:
        json.dump(rows, f, indent=1)
