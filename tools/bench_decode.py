"""Decode-step benchmark: DB1-1.3B, B sequences, memory length 1024 (= n_position), one new token per call
as in evaluate_rl.py:get_action. Compares forward(..., mems=<KVMemory>) (cached keys / values, db1_relattn_decode) with the
reference-format path forward(..., mems=[hidden states]) that re-projects cat(mem, w) on every call.
    python tools/bench_decode.py [B] [steps]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bdm-db1_b200"))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from db1_sm100 import ops  # noqa: E402
from src.data.input_specs import RLTaskInput  # noqa: E402
from src.model import TransformerXL  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dev = torch.device("cuda")
cfg = bench.make_config()
torch.manual_seed(0)
with torch.device(dev):
    model = TransformerXL(cfg)
model = model.half().to(dev).eval()


def inp(q):
    tok = torch.randint(32000, 33024, (B, q), device=dev)
    return [RLTaskInput(position_id=torch.zeros(B, q, dtype=torch.int64, device=dev), attention_mask=None, loss_mask=None,
                        label=None, text_seq=None, vision_seq=None, tensor_seq=tok)]


def run(mems, n, q=1):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        for _ in range(3):
            _, _, mems = model(inp(q), compute_loss=False, mems=mems)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            logits, _, mems = model(inp(q), compute_loss=False, mems=mems)
            ops.masked_argmax(logits[:, -1, :].contiguous(), cfg.text_vocab_size, logits.shape[-1] - 1)
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


prof = ops.set_profile(ops.Profile(timing=True))
ms_kv = run(model.init_mem(B, kv_cache=True), steps)
ops.set_profile(None)
s = prof.summary()
dec = s.get("relattn_decode")


def run_graph(n, q=1):
    from db1_sm100.functions import DecodeGraph
    mem = model.init_mem(B, kv_cache=True)
    g = DecodeGraph(model, mem, q)
    tok = torch.randint(32000, 33024, (B, q), device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        g.step(tok)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        logits = g.step(tok)
        ops.masked_argmax(logits[:, -1, :].contiguous(), cfg.text_vocab_size, logits.shape[-1] - 1)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


ms_graph = run_graph(steps)
ms_graph24 = run_graph(max(4, steps // 4), q=24)
ms_list = run(model.init_mem(B), max(4, steps // 4))
ms_kv24 = run(model.init_mem(B, kv_cache=True), max(4, steps // 4), q=24)
ms_list24 = run(model.init_mem(B), max(4, steps // 4), q=24)
out = {"what": "DB1-1.3B decode step, B=%d, mem_len=1024, fp16, eval; step = forward of the new token(s) + masked arg-max" % B,
       "cached_kv_graph_ms_per_step_q1": ms_graph, "cached_kv_graph_tokens_per_s_q1": B * 1e3 / ms_graph,
       "cached_kv_graph_ms_per_step_q24": ms_graph24, "speedup_graph_vs_recompute_q1": ms_list / ms_graph, "speedup_graph_vs_recompute_q24": ms_list24 / ms_graph24,
       "cached_kv_ms_per_step_q1": ms_kv, "recompute_ms_per_step_q1": ms_list, "speedup_q1": ms_list / ms_kv,
       "cached_kv_tokens_per_s_q1": B * 1e3 / ms_kv, "recompute_tokens_per_s_q1": B * 1e3 / ms_list,
       "cached_kv_ms_per_step_q24": ms_kv24, "recompute_ms_per_step_q24": ms_list24, "speedup_q24": ms_list24 / ms_kv24,
       "relattn_decode_us": dec["ms"] / dec["n"] * 1e3 if dec else None,
       "relattn_decode_GBs": (dec["bytes"] / (dec["ms"] / 1e3) / 1e9) if dec else None,
       "weights_streamed_GB_per_step": sum(p.numel() for p in model.parameters() if p.dim() == 2 and p.shape[0] != 33025) * 2 / 1e9}
print(json.dumps(out, indent=1))
