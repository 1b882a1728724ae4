"""Shared test helpers: golden fixture loading and conversion of oracle-style task dicts to the model's input types."""
import ast
import os
from types import SimpleNamespace

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False)


def golden_cfg(g):
    return SimpleNamespace(**ast.literal_eval(str(g["cfg"])))


def tasks_from_golden(g):
    tasks = []
    ti = 0
    while "task%d:type" % ti in g:
        t = {"type": str(g["task%d:type" % ti])}
        pre = "task%d:" % ti
        for k in g.files:
            if k.startswith(pre) and k != pre + "type":
                t[k[len(pre):]] = g[k]
        t.setdefault("vision_seq", None)
        tasks.append(t)
        ti += 1
    return tasks


def to_model_inputs(tasks, device, half=True):
    """Oracle task dicts -> src.data.input_specs objects on `device` (pixels in fp16 when `half`)."""
    from src.data.input_specs import ICTaskInput, NLPTaskInput, RLTaskInput
    out = []

    def T(t, k, dtype=None):
        v = t.get(k)
        if v is None:
            return None
        x = torch.as_tensor(np.asarray(v))
        if dtype is not None:
            x = x.to(dtype)
        return x.to(device)

    pix = torch.float16 if half else torch.float32
    for t in tasks:
        if t["type"] == "rl":
            out.append(RLTaskInput(position_id=T(t, "position_id", torch.int64), attention_mask=None,
                                   loss_mask=T(t, "loss_mask", torch.float32), label=T(t, "label", torch.int64),
                                   text_seq=None, vision_seq=T(t, "vision_seq", pix),
                                   tensor_seq=T(t, "tensor_seq", torch.int64)))
        elif t["type"] == "nlp":
            out.append(NLPTaskInput(position_id=None, attention_mask=None, loss_mask=T(t, "loss_mask", torch.float32),
                                    label=T(t, "label", torch.int64), text_seq=T(t, "text_seq", torch.int64),
                                    text_len=None))
        else:
            out.append(ICTaskInput(position_id=None, attention_mask=None, loss_mask=T(t, "loss_mask", torch.float32),
                                   label=T(t, "label", torch.int64), prompt_seq=T(t, "prompt_seq", torch.int64),
                                   img_seq=T(t, "img_seq", pix), text_seq=T(t, "text_seq", torch.int64),
                                   img_id_seq=None))
    return out


def rel_err(a, b):
    a = a.detach().float().cpu()
    b = b.detach().float().cpu()
    return (a - b).abs().max().item() / (b.abs().max().item() + 1e-20)


def rel_l2(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()
