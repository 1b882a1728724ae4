"""Per-kernel durations of ONE replayed decode step (DB1-1.3B, B=1, q=1, mem_len=1024) for a run under
    ncu --metrics gpu__time_duration.sum --profile-from-start off --graph-profiling node --csv --log-file <csv> python tools/decode_kernel_times.py
(cudaProfilerStart / Stop bracket the replay)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bdm-db1_b200"))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from db1_sm100.functions import DecodeGraph  # noqa: E402
from src.model import TransformerXL  # noqa: E402

dev = torch.device("cuda")
cfg = bench.make_config()
torch.manual_seed(0)
with torch.device(dev):
    model = TransformerXL(cfg)
model = model.half().to(dev).eval()
q = int(sys.argv[1]) if len(sys.argv) > 1 else 1
mem = model.init_mem(1, kv_cache=True)
g = DecodeGraph(model, mem, q)
tok = torch.randint(32000, 33024, (1, q), device=dev)
for _ in range(3):
    g.step(tok)
torch.cuda.synchronize()
torch.cuda.profiler.start()
g.step(tok)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
