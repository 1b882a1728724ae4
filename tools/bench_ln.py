"""Development: LayerNorm forward / backward kernel time at the model shape (4096 x 2048), inputs cold (L2 flushed)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bdm-db1_b200"))
from db1_sm100 import ops
dev = torch.device("cuda")
R, d = 4096, 2048
y = torch.randn(R, d, device=dev).half(); g = torch.randn(d, device=dev).half(); b = torch.randn(d, device=dev).half()
out = torch.empty_like(y); stats = torch.empty(R, 2, device=dev)
dout = torch.randn(R, d, device=dev).half(); dy = torch.empty_like(y); dz = torch.empty_like(y)
small = torch.zeros(3 * d, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def t(name, fn, nbytes):
    ts = []
    for _ in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / 2)
    print("%-10s %.1f us  %.2f TB/s" % (name, min(ts), nbytes / min(ts) / 1e6))
t("ln_fwd", lambda: ops.layernorm_fwd(y, g, b, out, stats, 1e-5), 2 * R * d * 2)
t("ln_bwd", lambda: ops.layernorm_bwd(dout, y, g, stats, dy, dz, small[0:d], small[d:2 * d], small[2 * d:], 0.1, 1234), 4 * R * d * 2)
B, H, L = 4, 16, 1024
dS = torch.randn(B, H, L, L, device=dev).half().tril_()
dSr = torch.zeros_like(dS)
t("rel_unshift", lambda: ops.rel_unshift(dS, dSr, B * H, L), 2.0 * B * H * L * (L + 1))
