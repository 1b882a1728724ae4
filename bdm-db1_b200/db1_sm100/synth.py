"""Seeded synthetic batches of the shapes named in BASELINE.json, built through the product's
own host path: mu-law discretiser (libdb1_host.so:db1_discretize) and RL token layout (db1_rl_layout)."""
import ctypes as C

import numpy as np
import torch

from src.data.input_specs import ICTaskInput, NLPTaskInput, RLTaskInput
from src.tokenizer.scalar_tokenizer import ContinuousScalarTokenizer

from . import _lib


def rl_layout(obs_tok, act_tok, sep_id, seq_len, pad_id=0, prepend_trans_num=0):
    """One sample: obs_tok [T, obs_len], act_tok [T, act_len] int64 -> (tensor_seq, label, loss_mask, position_id)."""
    obs_tok = np.ascontiguousarray(obs_tok, dtype=np.int64)
    act_tok = np.ascontiguousarray(act_tok, dtype=np.int64)
    T, ol = obs_tok.shape
    al = act_tok.shape[1]
    ts = np.empty(seq_len, np.int64)
    lb = np.empty(seq_len, np.int64)
    lm = np.empty(seq_len, np.float32)
    ps = np.empty(seq_len, np.int64)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    rc = _lib.hostlib().db1_rl_layout(p(obs_tok), p(act_tok), T, ol, al, C.c_longlong(sep_id), seq_len,
                                      C.c_longlong(pad_id), prepend_trans_num, p(ts), p(lb), p(lm), p(ps))
    if rc != 0:
        raise _lib.Db1Error("db1_rl_layout failed rc=%d" % rc)
    return ts, lb, lm, ps


def _vocab(cfg):
    text = cfg.text_vocab_size
    cont0 = text if cfg.overlap_with_text else text + cfg.num_discrete_values
    total = text + cfg.num_continuous_bin + (0 if cfg.overlap_with_text else cfg.num_discrete_values)
    return text, cont0, total  # total == separator id


def rl_continuous_batch(cfg, B, L, obs_len=17, act_len=6, seed=1234, pin=False):
    """Config C2: continuous-control trajectories; raw floats ~ N(0,1) (obs) / U(-1,1) (actions) -> discretiser."""
    rng = np.random.default_rng(seed)
    tk = ContinuousScalarTokenizer(cfg.num_continuous_bin)
    _text, cont0, sep = _vocab(cfg)
    step = obs_len + act_len + 1
    T = L // step + 1
    rows = []
    for _ in range(B):
        obs = tk.discretize(rng.standard_normal((T, obs_len)).astype(np.float32), is_action=False).numpy()
        act = tk.discretize(rng.uniform(-1, 1, (T, act_len)).astype(np.float32), is_action=True).numpy()
        rows.append(rl_layout(obs.astype(np.int64) + cont0, act.astype(np.int64) + cont0, sep, L))
    return _rl_input(rows, None, pin)


def rl_atari_batch(cfg, B, L, frame_hw=(80, 80), n_actions=18, seed=1234, pin=False, pixel_dtype=torch.float16):
    """Config C3: one image frame (patch slots = -1) + separator + one discrete action per transition. Frames are
    generated at 84x84 and centre-cropped to 80x80 because the reference's patch rearrange needs multiples of 16."""
    rng = np.random.default_rng(seed)
    ps = cfg.vision_patch_size
    h0, w0 = frame_hw[0] // ps, frame_hw[1] // ps
    npatch = h0 * w0
    _text, _cont0, sep = _vocab(cfg)
    step = npatch + 2
    T = L // step + 1
    rows = []
    for _ in range(B):
        obs = -np.ones((T, npatch), dtype=np.int64)
        act = rng.integers(0, n_actions, size=(T, 1)).astype(np.int64)
        if not cfg.overlap_with_text:
            act = act + cfg.text_vocab_size
        rows.append(rl_layout(obs, act, sep, L))
    nslots = int((rows[0][0] == -1).sum())
    nfr = (nslots + npatch - 1) // npatch
    big = rng.integers(0, 256, size=(B, nfr, 3, 84, 84)).astype(np.float32) / 255.0
    y0, x0 = (84 - frame_hw[0]) // 2, (84 - frame_hw[1]) // 2
    frames = torch.from_numpy(big[..., y0:y0 + frame_hw[0], x0:x0 + frame_hw[1]].copy()).to(pixel_dtype)
    return _rl_input(rows, frames, pin)


def _rl_input(rows, frames, pin):
    def stack(i, dt):
        t = torch.from_numpy(np.stack([r[i] for r in rows])).to(dt)
        return t.pin_memory() if pin else t
    if frames is not None and pin:
        frames = frames.pin_memory()
    return RLTaskInput(position_id=stack(3, torch.int64), attention_mask=None, loss_mask=stack(2, torch.float32),
                       label=stack(1, torch.int64), text_seq=None, vision_seq=frames, tensor_seq=stack(0, torch.int64))


def nlp_batch(cfg, B, L, seed=1234, pin=False):
    rng = np.random.default_rng(seed)
    txt = torch.from_numpy(rng.integers(0, cfg.text_vocab_size, size=(B, L + 1)).astype(np.int64))
    mk = (lambda t: t.pin_memory()) if pin else (lambda t: t)
    return NLPTaskInput(position_id=None, attention_mask=None, loss_mask=mk(torch.ones(B, L)),
                        label=mk(txt[:, 1:].contiguous()), text_seq=mk(txt[:, :-1].contiguous()), text_len=None)


def ic_batch(cfg, B, L, image_hw=(224, 224), prompt_len=10, seed=1234, pin=False, pixel_dtype=torch.float16):
    rng = np.random.default_rng(seed)
    ps = cfg.vision_patch_size
    npatch = (image_hw[0] // ps) * (image_hw[1] // ps)
    nt = L - prompt_len - npatch
    assert nt > 0, "sequence too short for the image"
    mk = (lambda t: t.pin_memory()) if pin else (lambda t: t)
    prompt = torch.from_numpy(rng.integers(0, cfg.text_vocab_size, size=(B, prompt_len)).astype(np.int64))
    text = torch.from_numpy(rng.integers(0, cfg.text_vocab_size, size=(B, nt)).astype(np.int64))
    label = torch.from_numpy(rng.integers(0, cfg.text_vocab_size, size=(B, L)).astype(np.int64))
    lm = torch.zeros(B, L)
    lm[:, prompt_len + npatch:] = 1
    img = torch.from_numpy(rng.integers(0, 256, size=(B, 3, image_hw[0], image_hw[1])).astype(np.float32) / 255.0)
    return ICTaskInput(position_id=None, attention_mask=None, loss_mask=mk(lm), label=mk(label), prompt_seq=mk(prompt),
                       img_seq=mk(img.to(pixel_dtype)), text_seq=mk(text), img_id_seq=None)


def to_device(task, device, non_blocking=True):
    """Copy of `task` with every tensor field moved to `device` (host buffers stay untouched for the next step)."""
    import copy
    from dataclasses import fields
    out = copy.copy(task)
    for f in fields(task):
        v = getattr(task, f.name)
        if isinstance(v, torch.Tensor):
            setattr(out, f.name, v.to(device, non_blocking=non_blocking))
    return out


def input_bytes(tasks):
    from dataclasses import fields
    n = 0
    for t in tasks:
        for f in fields(t):
            v = getattr(t, f.name)
            if isinstance(v, torch.Tensor):
                n += v.numel() * v.element_size()
    return n
