"""Development: run the two band kernels of the attention backward at a given shape and report time / parity."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bdm-db1_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from db1_sm100 import ops  # noqa: E402

B, L, H, dh = [int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (4, 1024, 16, 128))]
window = int(sys.argv[5]) if len(sys.argv) > 5 else L
dev = torch.device("cuda")
d = H * dh
g = torch.Generator(device="cuda").manual_seed(1)
i = torch.arange(L, device=dev)[:, None]
j = torch.arange(L, device=dev)[None, :]
ok = (j <= i) & (i - j < window)
ds = (torch.randn(ops.score_tiles_shape(B, L, H), generator=g, device=dev)).half()
qkv4 = (torch.randn(B * L, 4 * d, generator=g, device=dev) * 0.7).half()
r = (torch.randn(L, d, generator=g, device=dev) * 0.7).half()
qv, kk = qkv4[:, d:2 * d], qkv4[:, 2 * d:3 * d]
dqkv = torch.zeros(B * L, 3 * d, dtype=torch.half, device=dev)
du = torch.zeros(d, dtype=torch.float32, device=dev)
dv = torch.zeros(d, dtype=torch.float32, device=dev)
dr = torch.zeros(L, d, dtype=torch.float32, device=dev)
torch.cuda.synchronize()
P = (torch.rand(ops.score_tiles_shape(B, L, H), generator=g, device=dev)).half()
do = (torch.randn(B * L, d, generator=g, device=dev) * 0.7).half()
for name, fn in (("dkdv", lambda: ops.relattn_bwd_dkdv(P, ds, do, qkv4[:, 0:d], dqkv[:, d:2 * d], dqkv[:, 2 * d:], B, L, H, dh, window)),
                 ("dq", lambda: ops.relattn_bwd_dq(ds, kk, r, dqkv[:, 0:d], du, dv, B, L, H, dh, window)),
                 ("dr", lambda: ops.relattn_bwd_dr(ds, qv, dr, B, L, H, dh, window))):
    print("launch", name, flush=True)
    t0 = time.time()
    fn()
    torch.cuda.synchronize()
    print("  first call ok %.3f s" % (time.time() - t0), flush=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print("  %s: %.1f us per call" % (name, e0.elapsed_time(e1) / 5 * 1e3), flush=True)
