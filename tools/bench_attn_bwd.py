"""One launch set of the attention-backward recompute kernel (db1_relattn_bwd_ds) at B=4, H=16, dh=128, L=1024 for ncu."""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bdm-db1_b200"))
from db1_sm100 import ops  # noqa: E402

dev = torch.device("cuda")
B, H, dh, L = 4, 16, 128, int(sys.argv[1]) if len(sys.argv) > 1 else 1024
d = H * dh
qkv4 = (torch.randn(B * L, 4 * d, device=dev) * 0.5).half()
r = (torch.randn(L, d, device=dev) * 0.5).half()
do = (torch.randn(B * L, d, device=dev) * 0.5).half()
o = torch.empty(B * L, d, dtype=torch.half, device=dev)
lse2 = torch.empty(B, H, L, dtype=torch.float32, device=dev)
drow = torch.empty(B, H, L, dtype=torch.float32, device=dev)
P = torch.zeros(B, H, L, L, dtype=torch.half, device=dev)
dS = torch.zeros(B, H, L, L, dtype=torch.half, device=dev)
scale = 1.0 / math.sqrt(dh)
ops.relattn_fwd(qkv4, r, o, lse2, B, L, H, dh, L, scale)
ops.rowdot(do, o, drow, B, L, H, dh)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(6):
    if it == 3:
        e0.record()
    ops.relattn_bwd_ds(qkv4, r, do, lse2, None, P, dS, B, L, H, dh, L, scale, o=o)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 3 * 1e3
print("relattn_bwd_ds L=%d  %.1f us  %.1f TFLOP/s (S, BD, dP contractions)" % (L, us, B * H * 6.0 * dh * (L * (L + 1) / 2) / us / 1e6))
