"""Development: one band kernel, one shape, parity + timing. usage: dbg_band2.py dq|dr B L H dh [window] [reps]"""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bdm-db1_b200"))
from db1_sm100 import ops  # noqa: E402
which = sys.argv[1]
B, L, H, dh = [int(x) for x in sys.argv[2:6]]
window = int(sys.argv[6]) if len(sys.argv) > 6 else L
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 3
dev = torch.device("cuda")
d = H * dh
g = torch.Generator(device="cuda").manual_seed(1)
i = torch.arange(L, device=dev)[:, None]
j = torch.arange(L, device=dev)[None, :]
ok = (j <= i) & (i - j < window)
ds = (torch.randn(B, H, L, L, generator=g, device=dev) * ok).half()
qkv4 = (torch.randn(B * L, 4 * d, generator=g, device=dev) * 0.7).half()
r = (torch.randn(L, d, generator=g, device=dev) * 0.7).half()
qv, kk = qkv4[:, d:2 * d], qkv4[:, 2 * d:3 * d]
dqkv = torch.zeros(B * L, 3 * d, dtype=torch.half, device=dev)
du = torch.zeros(d, dtype=torch.float32, device=dev)
dv = torch.zeros(d, dtype=torch.float32, device=dev)
dr = torch.zeros(L, d, dtype=torch.float32, device=dev)
torch.cuda.synchronize()
fn = (lambda: ops.relattn_bwd_dq(ds, kk, r, dqkv[:, 0:d], du, dv, B, L, H, dh, window)) if which == "dq" else \
     (lambda: ops.relattn_bwd_dr(ds, qv, dr, B, L, H, dh, window))
for rep in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    print("%s B%d L%d H%d dh%d rep %d: %.1f us" % (which, B, L, H, dh, rep, e0.elapsed_time(e1) * 1e3), flush=True)
