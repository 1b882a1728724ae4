"""Runs the UNMODIFIED reference (src.model.TransformerXL) installed under oracle/_ref/reference by oracle/Makefile.

TEST INFRASTRUCTURE. Used by bench.py only (`--impl reference` = the reference's CPU path on the host cores;
`gpu_eager_baseline` = the same reference module `.half().cuda()` timed on the same B200),
always in its OWN process: the reference and the product both use the top-level package name `src`, and the
reference arm must not map any product binary. Nothing under bdm-db1_b200/ is imported here; the synthetic batch comes
from the oracle's own discretiser / RL token layout (oracle/db1_oracle.py).

    python oracle/ref_runner.py --device cpu|cuda [--half] [--batch B] [--seq L] [--layers N] [--iters K] [--warmup W]
prints one JSON object: {"tokens_per_s", "ms_per_step", "iters", "threads", "device", "dtype", "loss"}.
"""
import argparse
import json
import os
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref", "reference")


def available():
    return os.path.isfile(os.path.join(REF, "src", "model", "transformer_xl.py"))


def _import_reference():
    if not available():
        raise RuntimeError("reference not installed under oracle/_ref/reference (run `make -C oracle`)")
    # keep the product package (bdm-db1_b200/src) out of this process: same top-level name `src`
    sys.path[:] = [p for p in sys.path if not p.rstrip("/").endswith("bdm-db1_b200")]
    sys.path.insert(0, REF)
    sys.path.insert(1, os.path.dirname(HERE))
    for m in ("gym", "d4rl", "tree"):  # imported (never used) by src.data.rl_dataset; absent from this image
        sys.modules.setdefault(m, types.ModuleType(m))
    from src.model import TransformerXL
    from src.data.input_specs import RLTaskInput
    return TransformerXL, RLTaskInput


def rl_batch(orc, cfg, B, L, seed, obs_len=17, act_len=6):
    """Config C2 through the oracle's discretiser and RL token layout."""
    import numpy as np
    import torch
    rng = np.random.default_rng(seed)
    cont0 = cfg.text_vocab_size if cfg.overlap_with_text else cfg.text_vocab_size + cfg.num_discrete_values
    sep = orc.total_vocab(cfg) - 1
    T = L // (obs_len + act_len + 1) + 1
    rows = []
    for _ in range(B):
        obs = orc.discretize(rng.standard_normal((T, obs_len)).astype(np.float32), False).astype(np.int64) + cont0
        act = orc.discretize(rng.uniform(-1, 1, (T, act_len)).astype(np.float32), True).astype(np.int64) + cont0
        rows.append(orc.rl_sequence(obs, act, sep, L))
    st = lambda k, dt: torch.from_numpy(np.stack([r[k] for r in rows])).to(dt)  # noqa: E731
    return dict(tensor_seq=st("tensor_seq", torch.int64), label=st("label", torch.int64),
                loss_mask=st("loss_mask", torch.float32), position_id=st("position_id", torch.int64))


def run(device="cpu", half=False, batch=1, seq=1024, layers=24, iters=2, warmup=1, train=True, budget_s=None):
    import torch
    TransformerXL, RLTaskInput = _import_reference()
    from oracle import db1_oracle as orc
    if device == "cpu":
        torch.set_num_threads(os.cpu_count() or 1)
    cfg = orc.default_config(n_layer=layers, drop=0.1 if train else 0.0, embd_pdrop=0.1 if train else 0.0,
                             vision_hidden_dropout_prob=0.1 if train else 0.0, fp16=bool(half))
    torch.manual_seed(0)
    model = TransformerXL(types.SimpleNamespace(**vars(cfg)))
    dev = torch.device(device)
    if half:
        model = model.half()
    model = model.to(dev)
    model.train(train)
    host = rl_batch(orc, cfg, batch, seq, seed=1234)

    def make_input():
        t = {k: v.clone().to(dev) for k, v in host.items()}
        return [RLTaskInput(position_id=t["position_id"], attention_mask=None, loss_mask=t["loss_mask"], label=t["label"],
                            text_seq=None, vision_seq=None, tensor_seq=t["tensor_seq"])]

    scale = 4096.0 if half else 1.0

    def step():
        _logits, loss = model(make_input())
        (loss * scale).backward()
        for p in model.parameters():
            p.grad = None
        return loss

    for _ in range(warmup):
        step()
    times = []
    loss = None
    if dev.type == "cuda":
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            loss = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        n = iters
    else:
        t_all = time.perf_counter()
        while len(times) < iters and (not times or budget_s is None or time.perf_counter() - t_all < budget_s):
            t0 = time.perf_counter()
            loss = step()
            times.append(time.perf_counter() - t0)
        ms = sum(times) / len(times) * 1e3
        n = len(times)
    return {"tokens_per_s": batch * seq / (ms / 1e3), "ms_per_step": ms, "iters": n, "warmup": warmup,
            "threads": torch.get_num_threads() if dev.type == "cpu" else None, "device": device,
            "dtype": "f16" if half else "f32", "batch": batch, "seq": seq, "layers": layers,
            "loss": float(loss.detach().float().cpu())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--half", action="store_true")
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--seq", type=int, default=1024)
    ap.add_argument("--layers", type=int, default=24)
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--eval", action="store_true")
    ap.add_argument("--budget-s", type=float, default=None)
    a = ap.parse_args()
    try:
        out = run(a.device, a.half, a.batch, a.seq, a.layers, a.iters, a.warmup, not a.eval, a.budget_s)
    except Exception as e:  # the caller records why there is no number
        out = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
    print("REF_RUNNER " + json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
