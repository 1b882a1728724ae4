"""Host-side batch assembly on libdb1_host.so (include/db1_host.h): one RL sample from raw arrays
(RLFullDataset.get, src/data/rl_dataset.py:614-752) and my_collate_fn (src/data/data_samplers.py:28-42)."""
import ctypes as C
from dataclasses import fields

import numpy as np
import torch

from src.data.input_specs import RLTaskInput

from . import _lib

_p = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else C.c_void_p(0)  # noqa: E731


def rl_assemble(obs_float, act, seq_len, frames=None, transition_num=None, text_vocab=32000, n_disc=1024, n_cont=1024,
                overlap_with_text=True, prepend_trans_num=0, patch=16):
    """obs_float [T, n_float] float32 or None; act [T, act_len] float (continuous) or integer (discrete) or [T];
    frames [T, C, H, W] float32 or None (zero-padded to `transition_num` frames like the reference, :660-664).
    Returns an RLTaskInput with a leading batch dimension of 1, exactly as RLFullDataset.get does."""
    act = np.asarray(act)
    if act.ndim == 1:
        act = act[:, None]
    T, act_len = act.shape
    n_float = 0
    of = None
    if obs_float is not None:
        of = np.ascontiguousarray(obs_float, dtype=np.float32)
        n_float = of.shape[1]
    n_img, n_frames = 0, 0
    if frames is not None:
        frames = np.asarray(frames, dtype=np.float32)
        n, c, h, w = frames.shape
        n_img = (h // patch) * (w // patch)
        n_frames = n
        if transition_num is not None and n < transition_num:
            pad = np.zeros((transition_num, c, h, w), dtype=np.float32)
            pad[:n] = frames
            frames, n_frames = pad, transition_num
    af = ad = None
    if "float" in act.dtype.name:
        af = np.ascontiguousarray(act, dtype=np.float32)
    else:
        ad = np.ascontiguousarray(act, dtype=np.int64)
    ts = np.empty(seq_len, np.int64)
    lb = np.empty(seq_len, np.int64)
    lm = np.empty(seq_len, np.float32)
    ps = np.empty(seq_len, np.int64)
    rc = _lib.hostlib().db1_rl_assemble(_p(of), n_float, n_img, _p(af), _p(ad), act_len, T, n_frames, text_vocab, n_disc,
                                        n_cont, int(bool(overlap_with_text)), seq_len, prepend_trans_num, _p(ts), _p(lb),
                                        _p(lm), _p(ps))
    if rc != 0:
        raise _lib.Db1Error("db1_rl_assemble failed rc=%d" % rc)
    t = lambda a: torch.from_numpy(a)[None]  # noqa: E731
    return RLTaskInput(position_id=t(ps), attention_mask=None, loss_mask=t(lm.astype(np.int64)), label=t(lb), text_seq=None,
                       vision_seq=torch.from_numpy(frames)[None] if frames is not None else None, tensor_seq=t(ts))


def collate(task_list, pin=False):
    """my_collate_fn: one object per task type (order of first appearance), tensor fields concatenated on dim 0, None
    fields stay None. The grouping comes from db1_collate_plan, the concatenation from db1_concat_rows straight into
    (optionally pinned) output tensors."""
    names = []
    ids = np.empty(len(task_list), np.int32)
    for i, t in enumerate(task_list):
        nm = type(t).__name__
        if nm not in names:
            names.append(nm)
        ids[i] = names.index(nm)
    n = len(task_list)
    perm = np.empty(n, np.int32)
    gtype = np.empty(n, np.int32)
    gcount = np.empty(n, np.int32)
    lib = _lib.hostlib()
    ng = lib.db1_collate_plan(_p(ids), n, _p(perm), _p(gtype), _p(gcount))
    if ng <= 0:
        raise _lib.Db1Error("db1_collate_plan failed rc=%d" % ng)
    out = []
    w = 0
    for g in range(ng):
        members = [task_list[i] for i in perm[w:w + gcount[g]]]
        w += gcount[g]
        first = members[0]
        kw = {}
        for f in fields(first):
            vals = [getattr(m, f.name) for m in members]
            if not isinstance(vals[0], torch.Tensor):
                kw[f.name] = vals[0]
                continue
            vals = [v.contiguous() for v in vals]
            rows = sum(v.shape[0] for v in vals)
            dst = torch.empty((rows,) + tuple(vals[0].shape[1:]), dtype=vals[0].dtype, pin_memory=pin)
            srcs = (C.c_void_p * len(vals))(*[v.data_ptr() for v in vals])
            nb = (C.c_longlong * len(vals))(*[v.numel() * v.element_size() for v in vals])
            rc = lib.db1_concat_rows(C.c_void_p(dst.data_ptr()), srcs, nb, len(vals))
            if rc != 0:
                raise _lib.Db1Error("db1_concat_rows failed rc=%d" % rc)
            kw[f.name] = dst
        out.append(type(first)(**kw))
    return out
