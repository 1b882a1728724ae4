"""CPU restatement of the DB1 hot path (TEST INFRASTRUCTURE — never imported by the product path).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this module.
It restates, in plain fp32 torch / numpy on the CPU, what the reference computes on the path named in
BASELINE.json (Shanghai-Digital-Brain-Laboratory/BDM-DB1, paths relative to /root/reference):

    src/model/transformer_xl.py      forward :506-619, attention :112-243, FFN :276-292, embeddings :621-748
    src/model/activations.py         GEGLU :19-32
    src/tokenizer/vision_embedding.py PatchEmbeddings :65-86, VisionEmbedding :117-180
    src/tokenizer/scalar_tokenizer.py discretize :28-45, decode :47-63
    src/data/rl_dataset.py           _get_action_flag_and_position_id :44-71, pad/shift :711-716, :738-746, :865-872

It is written functionally over a flat {state_dict key: tensor} mapping that uses the reference's own parameter
names, so that a reference checkpoint, the reference module and the B200 module can all be compared through it.
Parity pin: tests/test_oracle_golden.py checks every function here against tests/golden/*.npz, which were produced by
tools/make_golden.py importing and running the unmodified reference in the build container. The reference itself ships
no tests or golden vectors, so those fixtures are the pin.
Gradients come from torch autograd applied to this restatement.
"""
import math
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------------------------------
# integer paths (bit-exact contracts)
# --------------------------------------------------------------------------------------------------------------------


def discretize(x, is_action, num_bins=1024, mu=100.0, M=256.0):
    """scalar_tokenizer.py:28-45 — mu-law (observations only), clamp to [-1,1], uniform bins. float32 throughout."""
    x = torch.as_tensor(np.array(x, dtype=np.float32))
    if not is_action:
        x = torch.sign(x) * torch.log(torch.abs(x) * mu + 1.0) / torch.log(torch.tensor(mu * M + 1.0))
        x = torch.clamp(x, -1, 1)
    x = ((x + 1) / 2 * num_bins).int()
    return torch.clamp(x, 0, num_bins - 1).int().numpy()


def decode(tok, is_action, num_bins=1024, mu=100.0, M=256.0):
    """scalar_tokenizer.py:47-63."""
    x = torch.as_tensor(np.array(tok)).float()
    x = torch.clamp(x, 0, num_bins - 1)
    x = (x / num_bins) * 2 - 1
    if not is_action:
        x = torch.sign(x) * ((1 + M * mu) ** torch.abs(x) - 1) / mu
    return x.numpy()


def action_flag_and_position_id(seq_length, obs_len, act_len, prepend_trans_num=0):
    """rl_dataset.py:44-71 for a window that starts on a transition boundary.

    A transition is [obs_len observation tokens, SEP, act_len action tokens]; position ids 1..obs_len+1 on the
    observation + separator, 0 on actions; action flag 1 on action tokens except inside the prompt prefix."""
    step = obs_len + act_len + 1
    flag = np.zeros(seq_length, dtype=np.int64)
    pos = np.zeros(seq_length, dtype=np.int64)
    for s in range(0, seq_length, step):
        n = min(obs_len + 1, seq_length - s)
        pos[s:s + n] = 1 + np.arange(n)
    for s in range(prepend_trans_num * step, seq_length, step):
        flag[s + obs_len + 1:min(seq_length, s + step)] = 1
    return flag, pos


def rl_sequence(obs_tokens, act_tokens, sep_id, seq_len, pad_id=0, prepend_trans_num=0):
    """Token layout of one RL sample: rl_dataset.py:683-697 (join), :711-716/:865-872 (pad or truncate to L+1),
    :738-746 (shift into input / label / loss_mask).

    obs_tokens [T, obs_len] and act_tokens [T, act_len] are already vocabulary ids (-1 = image patch slot).
    Returns dict(tensor_seq [L], label [L], loss_mask [L] float32, position_id [L])."""
    obs_tokens = np.asarray(obs_tokens, dtype=np.int64)
    act_tokens = np.asarray(act_tokens, dtype=np.int64)
    T, obs_len = obs_tokens.shape
    act_len = act_tokens.shape[1]
    sep = np.full((T, 1), sep_id, dtype=np.int64)
    flat = np.concatenate([obs_tokens, sep, act_tokens], axis=1).reshape(-1)
    flag, pos = action_flag_and_position_id(len(flat), obs_len, act_len, prepend_trans_num)
    n = seq_len + 1
    if len(flat) >= n:
        flat, flag, pos = flat[:n], flag[:n], pos[:n]
    else:
        padn = n - len(flat)
        flat = np.concatenate([flat, np.full(padn, pad_id, dtype=np.int64)])
        flag = np.concatenate([flag, np.zeros(padn, dtype=np.int64)])
        pos = np.concatenate([pos, np.zeros(padn, dtype=np.int64)])
    return dict(tensor_seq=flat[:-1].copy(), label=flat[1:].copy(), loss_mask=flag[1:].astype(np.float32),
                position_id=pos[:-1].copy())


def patch_position_indices(h0, w0, vocab=128):
    """vision_embedding.py:130-148, 170-172 (eval branch): midpoint of each patch's interval on a `vocab`-step axis.
    The reference does the division in float32 and truncates."""
    seq = torch.arange(h0 * w0)
    row = torch.div(seq, w0, rounding_mode="trunc")
    col = seq % w0
    col_hi = ((col + 1) / w0 * vocab).to(torch.int32)
    col_lo = (col / w0 * vocab).to(torch.int32)
    row_hi = ((row + 1) / h0 * vocab).to(torch.int32)
    row_lo = (row / h0 * vocab).to(torch.int32)
    return (((row_lo + row_hi) / 2).int().numpy(), ((col_lo + col_hi) / 2).int().numpy(),
            row_lo.numpy(), row_hi.numpy(), col_lo.numpy(), col_hi.numpy())


# --------------------------------------------------------------------------------------------------------------------
# floating-point path
# --------------------------------------------------------------------------------------------------------------------


def default_config(**kw):
    """Hyper-parameters of the released 1.3B run (scripts/evaluate/evaluate_rl_1.2B.sh:16-19, 70-83)."""
    c = dict(n_embed=2048, n_position=1024, n_layer=24, n_head=16, n_inner=8192, pre_lnorm=False, mem_len=1024,
             same_length=True, untie_r=False, text_vocab_size=32000, num_discrete_values=1024,
             num_continuous_bin=1024, overlap_with_text=True, embd_pdrop=0.0, drop=0.0, dropattn=0.0,
             activation_fn="geglu", layer_norm_epsilon=1e-5, share_input_output_embedding=True, use_deepnorm=False,
             fp16=False, vision_patch_size=16, vision_num_input_channels=3, vision_position_vocab_size=128,
             vision_hidden_dropout_prob=0.0)
    c.update(kw)
    return SimpleNamespace(**c)


def tiny_config(**kw):
    """DB1-tiny of BASELINE.json configs[0]: 2 layers, d_model 128, 4 heads, GeGLU 512."""
    base = dict(n_embed=128, n_position=256, n_layer=2, n_head=4, n_inner=512, mem_len=256)
    base.update(kw)
    return default_config(**base)


def total_vocab(cfg):
    """transformer_xl.py:382-390."""
    v = cfg.text_vocab_size + cfg.num_continuous_bin + (0 if cfg.overlap_with_text else cfg.num_discrete_values)
    return v + 1


def positional_rows(klen, d_model, clamp_len, half_phase=False):
    """transformer_xl.py:34-50, 569-574: row c holds the sinusoid of distance min(klen-1-c, clamp_len).
    half_phase=True restates what those lines compute after module.half() (DeepSpeed fp16): pos_seq is built in fp16,
    inv_freq is the fp16-cast buffer, their outer product and the sin / cos results are fp16 (pinned by
    tests/golden/posemb_half.npz, generated from the reference module after .half())."""
    inv_freq = 1 / (10000 ** (torch.arange(0.0, d_model, 2.0) / d_model))
    pos = torch.arange(klen - 1, -1, -1.0)
    if clamp_len > 0:
        pos = pos.clamp(max=clamp_len)
    if half_phase:
        s = torch.outer(pos.half().float(), inv_freq.half().float()).half().float()
        return torch.cat([s.sin().half().float(), s.cos().half().float()], dim=-1)
    s = torch.outer(pos, inv_freq)
    return torch.cat([s.sin(), s.cos()], dim=-1)


def attention_mask_ok(qlen, klen, mem_len, same_length):
    """Boolean [qlen, klen]: True where attention is allowed. transformer_xl.py:551-567 restated as a predicate on
    delta = mlen + i - j: allowed iff delta >= 0 and (not same_length or klen <= mem_len or delta < mem_len)."""
    mlen = klen - qlen
    i = torch.arange(qlen)[:, None]
    j = torch.arange(klen)[None, :]
    delta = mlen + i - j
    ok = delta >= 0
    if same_length and klen > mem_len:
        ok = ok & (delta < mem_len)
    return ok


def rel_attention_core(q, k, v, rk, u, vb, ok, scale):
    """transformer_xl.py:161-225 without the pad/view trick.

    q [B,Q,H,D], k,v [B,K,H,D], rk [K,H,D] (row c = distance K-1-c), u,vb [H,D], ok [Q,K] bool.
    BD[i,j] = (q_i+vb).rk[j + Q-1-i] for j <= i + (K-Q)  (the identity behind _rel_shift, :98-110)."""
    B, Q, H, D = q.shape
    K = k.shape[1]
    ac = torch.einsum("bihd,bjhd->bhij", q + u, k)
    bd_raw = torch.einsum("bihd,chd->bhic", q + vb, rk)
    i = torch.arange(Q)[:, None]
    j = torch.arange(K)[None, :]
    idx = (j + Q - 1 - i).clamp(0, K - 1)
    bd = torch.gather(bd_raw, 3, idx[None, None].expand(B, H, Q, K))
    s = (ac + bd) * scale
    s = s.masked_fill(~ok[None, None], -1e30)
    p = torch.softmax(s, dim=-1)
    return torch.einsum("bhij,bjhd->bihd", p, v), p, s


def decoder_layer(x, pe, sd, prefix, cfg, ok, mem=None):
    """One RelPartialLearnableDecoderLayer (transformer_xl.py:326-353) in post-LN form (pre_lnorm=False)."""
    assert not cfg.pre_lnorm, "the released model is post-LN; pre-LN is not restated"
    H = cfg.n_head
    d = cfg.n_embed
    D = d // H
    B, Q, _ = x.shape
    a = prefix + "dec_attn."
    cat = x if mem is None else torch.cat([mem, x], 1)
    heads = F.linear(cat, sd[a + "qkv_net.weight"])
    rk = F.linear(pe, sd[a + "r_net.weight"]).view(-1, H, D)
    q, k, v = torch.chunk(heads, 3, dim=-1)
    q = q[:, -Q:]
    K = k.shape[1]
    u = sd["r_w_bias"] if not cfg.untie_r else sd[a + "r_w_bias"]
    vb = sd["r_r_bias"] if not cfg.untie_r else sd[a + "r_r_bias"]
    o, _, _ = rel_attention_core(q.reshape(B, Q, H, D), k.reshape(B, K, H, D), v.reshape(B, K, H, D), rk, u, vb, ok,
                                 1.0 / math.sqrt(D))
    attn_out = F.linear(o.reshape(B, Q, d), sd[a + "o_net.weight"])
    alpha = (2 * cfg.n_layer) ** 0.25 if cfg.use_deepnorm else 1.0
    x = F.layer_norm(x * alpha + attn_out, (d,), sd[a + "layer_norm.weight"], sd[a + "layer_norm.bias"],
                     cfg.layer_norm_epsilon)
    f = prefix + "pos_ff."
    hmid = F.linear(x, sd[f + "CoreNet.0.weight"], sd[f + "CoreNet.0.bias"])
    if cfg.activation_fn == "geglu":
        aa, gg = hmid.chunk(2, dim=-1)  # activations.py:26-29: value first, gate second
        act = aa * F.gelu(gg)
    else:
        act = F.gelu(hmid)
    core = F.linear(act, sd[f + "CoreNet.2.weight"], sd[f + "CoreNet.2.bias"])
    return F.layer_norm(x * alpha + core, (d,), sd[f + "layer_norm.weight"], sd[f + "layer_norm.bias"],
                        cfg.layer_norm_epsilon)


def patch_embeddings(pixels, sd, prefix, cfg):
    """vision_embedding.py:65-86: per-patch standardisation, ResNet-v2 block, 16x16/stride-16 projection.
    pixels [N,C,H,W] -> [N, (H/16)*(W/16), d]."""
    ps = cfg.vision_patch_size
    N, C, Hh, Ww = pixels.shape
    h0, w0 = Hh // ps, Ww // ps
    x = pixels.reshape(N, C, h0, ps, w0, ps).permute(0, 2, 4, 1, 3, 5).reshape(N * h0 * w0, C, ps, ps).float()
    mean = x.mean(dim=(-2, -1), keepdim=True)
    std = x.std(dim=(-2, -1), keepdim=True)  # unbiased
    x = (x - mean) / (1e-6 + std)
    x = x / math.sqrt(ps)
    x = F.conv2d(x, sd[prefix + "conv1.weight"], sd[prefix + "conv1.bias"], padding=1)
    res = x
    x = F.group_norm(x, 32, sd[prefix + "residual_path.0.weight"], sd[prefix + "residual_path.0.bias"], 1e-5)
    x = F.conv2d(F.gelu(x), sd[prefix + "residual_path.2.weight"], sd[prefix + "residual_path.2.bias"], padding=1)
    x = F.group_norm(x, 32, sd[prefix + "residual_path.3.weight"], sd[prefix + "residual_path.3.bias"], 1e-5)
    x = F.conv2d(F.gelu(x), sd[prefix + "residual_path.5.weight"], sd[prefix + "residual_path.5.bias"], padding=1)
    x = res + x
    x = F.conv2d(x, sd[prefix + "projection.weight"], sd[prefix + "projection.bias"], stride=ps)
    return x.view(N, h0 * w0, -1)


def vision_embedding(pixels, sd, cfg, prefix="vision_encoder."):
    """vision_embedding.py:117-180, eval branch (deterministic midpoint position indices)."""
    ps = cfg.vision_patch_size
    emb = patch_embeddings(pixels, sd, prefix + "patch_embeddings.", cfg)
    h0, w0 = pixels.shape[2] // ps, pixels.shape[3] // ps
    row, col = patch_position_indices(h0, w0, cfg.vision_position_vocab_size)[:2]
    row = torch.as_tensor(row).long()
    col = torch.as_tensor(col).long()
    return emb + sd[prefix + "row_position_embeddings.weight"][row][None] + \
        sd[prefix + "col_position_embeddings.weight"][col][None]


def embed_task(task, sd, cfg):
    """transformer_xl.py:621-748. `task` is a dict with key 'type' in {'rl','nlp','ic','vqa'} and the reference's
    field names. Returns (embeddings [B,L,d], loss_mask, label) — label with -1 replaced by 0 as at :645."""
    W = sd["word_embedding.weight"]
    t = task["type"]
    if t == "rl":
        tok = torch.as_tensor(task["tensor_seq"]).long()
        B, L = tok.shape
        emb = W[tok.clamp(min=0)] * (tok >= 0)[..., None]
        label = task.get("label")
        if task.get("vision_seq") is not None:
            img = torch.as_tensor(task["vision_seq"])
            vis = vision_embedding(img.reshape(-1, *img.shape[-3:]), sd, cfg).reshape(B, -1, cfg.n_embed)
            # the k-th -1 slot of row b receives patch embedding vis[b, k]  (:639-642; every row has the same count)
            slot = (tok == -1)
            kth = torch.cumsum(slot.long(), dim=1) - 1
            gathered = torch.gather(vis, 1, kth.clamp(min=0)[..., None].expand(B, L, cfg.n_embed))
            emb = torch.where(slot[..., None], gathered, emb)
            if label is not None:
                label = torch.as_tensor(label).long().clone()
                label[label == -1] = 0
        emb = emb + sd["rl_local_timestep_embedding.weight"][torch.as_tensor(task["position_id"]).long()]
        return emb, task.get("loss_mask"), label
    if t == "nlp":
        return W[torch.as_tensor(task["text_seq"]).long()], task.get("loss_mask"), task.get("label")
    if t in ("ic", "vqa"):
        prompt = W[torch.as_tensor(task["prompt_seq"]).long()]
        vis = vision_embedding(torch.as_tensor(task["img_seq"]), sd, cfg)
        text = W[torch.as_tensor(task["text_seq"]).long()]
        return torch.cat([prompt, vis, text], 1), task.get("loss_mask"), task.get("label")
    raise ValueError(t)


def forward(tasks, sd, cfg, compute_loss=True, mems=None, return_hidden=False):
    """TransformerXL.forward (transformer_xl.py:506-619), eval semantics (no dropout). sd: fp32 tensors."""
    assert not (compute_loss and mems is not None)
    embs, masks, labels = [], [], []
    for t in tasks:
        e, m, l = embed_task(t, sd, cfg)
        embs.append(e)
        masks.append(m)
        labels.append(l)
    x = torch.cat(embs, 0)
    Q = x.shape[1]
    mlen = mems[0].shape[1] if mems is not None else 0
    K = Q + mlen
    mem_len = cfg.mem_len if cfg.mem_len is not None else 0
    if cfg.same_length:
        ok = attention_mask_ok(Q, K, mem_len, True)
    else:
        ok = attention_mask_ok(Q, K, mem_len, False)
    pe = positional_rows(K, cfg.n_embed, cfg.n_position, half_phase=bool(getattr(cfg, "pos_phase_half", False)))
    hids = []
    for li in range(cfg.n_layer):
        hids.append(x)
        x = decoder_layer(x, pe, sd, "h.%d." % li, cfg, ok, None if mems is None else mems[li])
    if cfg.share_input_output_embedding:
        logits = F.linear(x, sd["word_embedding.weight"])
    else:
        logits = F.linear(x, sd["lm_head.weight"])
    loss = None
    if compute_loss:
        lab = torch.cat([torch.as_tensor(l).long() for l in labels], 0)
        msk = torch.cat([torch.as_tensor(m).float() for m in masks], 0)
        ce = F.cross_entropy(logits.reshape(-1, logits.shape[-1]).float(), lab.reshape(-1), reduction="none")
        loss = (ce * msk.reshape(-1)).sum() / msk.sum()
    out = (logits, loss)
    if mems is not None:
        end = mlen + max(0, Q)
        beg = max(0, end - mem_len)
        new_mems = [torch.cat([mems[i], hids[i]], 1)[:, beg:end].detach() for i in range(len(hids))]
        out = out + (new_mems,)
    if return_hidden:
        out = out + (x,)
    return out


def state_shapes(cfg):
    """{state_dict key: shape} of the reference TransformerXL (345 keys for the 1.3B config incl. the aliased
    ic_encoder.* entries, transformer_xl.py:394-433, vision_embedding.py:50-113)."""
    d, H = cfg.n_embed, cfg.n_head
    D = d // H
    V = total_vocab(cfg)
    inner = cfg.n_inner if cfg.n_inner is not None else 4 * d
    mid = inner // 2 if cfg.activation_fn == "geglu" else inner
    s = {"word_embedding.weight": (V, d), "pos_emb.inv_freq": (d // 2,)}
    if not cfg.untie_r:
        s["r_w_bias"] = (H, D)
        s["r_r_bias"] = (H, D)
    ps, C = cfg.vision_patch_size, cfg.vision_num_input_channels
    for enc in ("vision_encoder.", "ic_encoder."):
        pe = enc + "patch_embeddings."
        s[pe + "conv1.weight"] = (64, C, 3, 3)
        s[pe + "conv1.bias"] = (64,)
        s[pe + "projection.weight"] = (d, 64, ps, ps)
        s[pe + "projection.bias"] = (d,)
        for n in ("0", "3"):
            s[pe + "residual_path.%s.weight" % n] = (64,)
            s[pe + "residual_path.%s.bias" % n] = (64,)
        for n in ("2", "5"):
            s[pe + "residual_path.%s.weight" % n] = (64, 64, 3, 3)
            s[pe + "residual_path.%s.bias" % n] = (64,)
        s[enc + "row_position_embeddings.weight"] = (cfg.vision_position_vocab_size, d)
        s[enc + "col_position_embeddings.weight"] = (cfg.vision_position_vocab_size, d)
    s["rl_local_timestep_embedding.weight"] = (513, d)
    for li in range(cfg.n_layer):
        a = "h.%d.dec_attn." % li
        s[a + "qkv_net.weight"] = (3 * d, d)
        s[a + "o_net.weight"] = (d, d)
        s[a + "r_net.weight"] = (d, d)
        if cfg.untie_r:
            s[a + "r_r_bias"] = (H, D)
            s[a + "r_w_bias"] = (H, D)
        else:
            s[a + "r_r_bias"] = (H, D)  # shared parameter registered again under every layer
            s[a + "r_w_bias"] = (H, D)
        s[a + "layer_norm.weight"] = (d,)
        s[a + "layer_norm.bias"] = (d,)
        f = "h.%d.pos_ff." % li
        s[f + "CoreNet.0.weight"] = (inner, d)
        s[f + "CoreNet.0.bias"] = (inner,)
        s[f + "CoreNet.2.weight"] = (d, mid)
        s[f + "CoreNet.2.bias"] = (d,)
        s[f + "layer_norm.weight"] = (d,)
        s[f + "layer_norm.bias"] = (d,)
    if not cfg.share_input_output_embedding:
        s["lm_head.weight"] = (V, d)
    return s


def _name_seed(name, seed):
    h = 1469598103934665603
    for ch in name.encode():
        h = ((h ^ ch) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return (h ^ (seed * 0x9E3779B97F4A7C15)) & 0x7FFFFFFFFFFFFFFF


def synth_state_dict(cfg, seed=0, std=0.02):
    """Deterministic random weights reproducible on any machine (CPU generator seeded per parameter name), used to
    load the SAME weights into the reference, this oracle and the B200 module. Scales follow _init_weights
    (transformer_xl.py:456-468) except that norm scales/biases and Linear biases are perturbed so that tests see them."""
    sd = {}
    shapes = state_shapes(cfg)
    for name, shape in shapes.items():
        if name.startswith("ic_encoder."):
            continue
        if name.endswith("r_r_bias") and name.startswith("h.") and not cfg.untie_r:
            continue
        if name.endswith("r_w_bias") and name.startswith("h.") and not cfg.untie_r:
            continue
        g = torch.Generator().manual_seed(_name_seed(name, seed))
        if name == "pos_emb.inv_freq":
            sd[name] = 1 / (10000 ** (torch.arange(0.0, cfg.n_embed, 2.0) / cfg.n_embed))
        elif "layer_norm.weight" in name or ("residual_path" in name and name.endswith(("0.weight", "3.weight"))):
            sd[name] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith(".bias"):
            sd[name] = 0.02 * torch.randn(shape, generator=g)
        elif "conv1.weight" in name:
            sd[name] = torch.randn(shape, generator=g) * (1.0 / math.sqrt(27))
        elif "residual_path" in name:
            sd[name] = torch.randn(shape, generator=g) * (1.0 / math.sqrt(576))
        elif "projection.weight" in name:
            sd[name] = torch.randn(shape, generator=g) * (std / 4)
        else:
            sd[name] = std * torch.randn(shape, generator=g)
    # aliases
    for name in list(sd):
        if name.startswith("vision_encoder."):
            sd["ic_encoder." + name[len("vision_encoder."):]] = sd[name]
    if not cfg.untie_r:
        for li in range(cfg.n_layer):
            sd["h.%d.dec_attn.r_r_bias" % li] = sd["r_r_bias"]
            sd["h.%d.dec_attn.r_w_bias" % li] = sd["r_w_bias"]
    return sd
