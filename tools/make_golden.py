"""Generate tests/golden/*.npz by importing and running the UNMODIFIED reference from /root/reference.

Run in the build container only (the reference does not travel to the GPU box):
    python tools/make_golden.py
The reference ships no tests or golden vectors, so these fixtures pin the CPU oracle (oracle/db1_oracle.py) to the
reference's actual behaviour; tests/test_oracle_golden.py then checks the oracle against them on any machine.
Weights come from oracle.db1_oracle.synth_state_dict (deterministic CPU generator), loaded into the reference module
with load_state_dict(strict=True) — this also proves the key/shape inventory in oracle.state_shapes().
"""
import os
import sys
import types
from argparse import Namespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(1, REF)
for name in ("gym", "d4rl", "tree"):
    sys.modules.setdefault(name, types.ModuleType(name))

from oracle import db1_oracle as orc  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def ref_namespace(cfg):
    return Namespace(**vars(cfg))


def gen_tokenizer():
    from src.tokenizer.scalar_tokenizer import ContinuousScalarTokenizer
    tk = ContinuousScalarTokenizer()
    g = torch.Generator().manual_seed(7)
    # bin edges of the mu-law map and their float neighbours, plus random draws over several scales
    k = torch.arange(0, 1025, dtype=torch.float64)
    y = k / 512 - 1
    edges = (torch.sign(y) * (torch.exp(torch.abs(y) * np.log(25601.0)) - 1) / 100).float()
    neigh = torch.cat([edges, torch.nextafter(edges, torch.tensor(1e9)), torch.nextafter(edges, torch.tensor(-1e9))])
    rnd = torch.cat([torch.randn(4000, generator=g), torch.randn(4000, generator=g) * 30,
                     (torch.rand(2000, generator=g) - 0.5) * 700, torch.randn(2000, generator=g) * 1e-3])
    special = torch.tensor([-300, -1, -.5, -1e-3, 0, 1e-3, .5, 1, 256, 300], dtype=torch.float32)
    obs = torch.cat([special, neigh, rnd]).float()
    act_edges = (torch.arange(0, 1025, dtype=torch.float64) / 512 - 1).float()
    act = torch.cat([torch.tensor([-1.5, -1, -.999, -.5, 0, .001, .5, .998, .999, 1, 2]), act_edges,
                     torch.nextafter(act_edges, torch.tensor(9.0)), torch.nextafter(act_edges, torch.tensor(-9.0)),
                     torch.rand(4000, generator=g) * 2.2 - 1.1]).float()
    toks = torch.arange(0, 1024, dtype=torch.int32)
    np.savez_compressed(os.path.join(OUT, "tokenizer.npz"),
                        obs=obs.numpy(), obs_tok=tk.discretize(obs.numpy(), is_action=False).numpy(),
                        act=act.numpy(), act_tok=tk.discretize(act.numpy(), is_action=True).numpy(),
                        toks=toks.numpy(), dec_obs=tk.decode(toks.clone(), is_action=False).numpy(),
                        dec_act=tk.decode(toks.clone(), is_action=True).numpy())


def gen_rl_layout():
    from src.data import rl_dataset as rd
    cases = [(19, 4, 3, 1), (257, 17, 6, 0), (1025, 17, 6, 2), (100, 25, 1, 0), (28, 24, 3, 0), (64, 1, 1, 3)]
    out = {}
    for n, (seq, ol, al, pre) in enumerate(cases):
        flag, pos = rd._get_action_flag_and_position_id(0, seq - 1, ol, al, pre)
        out["case%d" % n] = np.array([seq, ol, al, pre])
        out["flag%d" % n] = flag
        out["pos%d" % n] = pos
    # pad / truncate rule (:865-872) on a toy sequence
    for n, (cur, tgt) in enumerate([(10, 16), (16, 16), (20, 16)]):
        x = np.arange(cur, dtype=np.int64) + 5
        out["pad_in%d" % n] = x
        out["pad_out%d" % n] = np.asarray(rd._truncate_or_pad_to_match_seq_len(x, tgt))
        out["pad_tgt%d" % n] = np.array([tgt])
    np.savez_compressed(os.path.join(OUT, "rl_layout.npz"), **out)


def gen_patch_positions():
    from src.tokenizer.vision_embedding import VisionEmbedding
    cfg = orc.tiny_config()
    ve = VisionEmbedding(ref_namespace(cfg)).eval()
    out = {}
    grids = [(5, 5), (4, 5), (14, 14), (6, 6), (1, 7), (2, 3)]
    for n, (h0, w0) in enumerate(grids):
        captured = {}

        def hook_row(mod, inp):
            captured["row"] = inp[0].clone()

        def hook_col(mod, inp):
            captured["col"] = inp[0].clone()

        h1 = ve.row_position_embeddings.register_forward_pre_hook(hook_row)
        h2 = ve.col_position_embeddings.register_forward_pre_hook(hook_col)
        with torch.no_grad():
            ve(torch.zeros(1, 3, h0 * 16, w0 * 16))
        h1.remove()
        h2.remove()
        out["grid%d" % n] = np.array([h0, w0])
        out["row%d" % n] = captured["row"].numpy().reshape(-1)
        out["col%d" % n] = captured["col"].numpy().reshape(-1)
    np.savez_compressed(os.path.join(OUT, "patch_positions.npz"), **out)


def make_tasks(cfg, seed, with_images, L):
    """Small mixed batch in plain numpy (dict-of-arrays form understood by the oracle)."""
    g = np.random.default_rng(seed)
    V = orc.total_vocab(cfg)
    sep = V - 1
    tasks = []
    # RL, continuous control: obs 5 + SEP + act 2
    T = L // 8 + 1
    rows = []
    for _ in range(2):
        obs = cfg.text_vocab_size + g.integers(0, 1024, size=(T, 5))
        act = cfg.text_vocab_size + g.integers(0, 1024, size=(T, 2))
        rows.append(orc.rl_sequence(obs, act, sep, L))
    tasks.append(dict(type="rl", tensor_seq=np.stack([r["tensor_seq"] for r in rows]),
                      label=np.stack([r["label"] for r in rows]), loss_mask=np.stack([r["loss_mask"] for r in rows]),
                      position_id=np.stack([r["position_id"] for r in rows]), vision_seq=None))
    # NLP
    txt = g.integers(0, cfg.text_vocab_size, size=(1, L + 1))
    tasks.append(dict(type="nlp", text_seq=txt[:, :-1].copy(), label=txt[:, 1:].copy(),
                      loss_mask=(g.random((1, L)) < 0.9).astype(np.float32)))
    if with_images:
        # RL with image observations: one 32x48 frame (6 patch slots) + 1 discrete action per transition
        npatch = 6
        step = npatch + 2
        T2 = L // step + 1
        obs = -np.ones((T2, npatch), dtype=np.int64)
        act = g.integers(0, 18, size=(T2, 1))
        r = orc.rl_sequence(obs, act, sep, L)
        nslots = int((r["tensor_seq"] == -1).sum())
        nfr = (nslots + npatch - 1) // npatch
        frames = np.round(g.random((1, nfr, 3, 32, 48)) * 255) / 255
        tasks.append(dict(type="rl", tensor_seq=r["tensor_seq"][None], label=r["label"][None],
                          loss_mask=r["loss_mask"][None], position_id=r["position_id"][None],
                          vision_seq=frames.astype(np.float32)))
        # image caption: prompt 4 + 32x32 image (4 patches) + text
        nt = L - 4 - 4
        cap = g.integers(0, cfg.text_vocab_size, size=(1, nt))
        lab = g.integers(0, cfg.text_vocab_size, size=(1, L))
        lm = np.zeros((1, L), dtype=np.float32)
        lm[:, 8:] = 1
        tasks.append(dict(type="ic", prompt_seq=g.integers(0, cfg.text_vocab_size, size=(1, 4)),
                          img_seq=(np.round(g.random((1, 3, 32, 32)) * 255) / 255).astype(np.float32),
                          text_seq=cap, label=lab, loss_mask=lm))
    return tasks


def to_ref_inputs(tasks):
    from src.data.input_specs import ICTaskInput, NLPTaskInput, RLTaskInput
    out = []
    for t in tasks:
        T = lambda k: None if t.get(k) is None else torch.as_tensor(t[k])  # noqa: E731
        if t["type"] == "rl":
            out.append(RLTaskInput(position_id=T("position_id"), attention_mask=None, loss_mask=T("loss_mask"),
                                   label=T("label").clone(), text_seq=None, vision_seq=T("vision_seq"),
                                   tensor_seq=T("tensor_seq")))
        elif t["type"] == "nlp":
            out.append(NLPTaskInput(position_id=None, attention_mask=None, loss_mask=T("loss_mask"), label=T("label"),
                                    text_seq=T("text_seq"), text_len=None))
        else:
            out.append(ICTaskInput(position_id=None, attention_mask=None, loss_mask=T("loss_mask"), label=T("label"),
                                   prompt_seq=T("prompt_seq"), img_seq=T("img_seq"), text_seq=T("text_seq"),
                                   img_id_seq=None))
    return out


GRAD_KEYS = ["r_w_bias", "r_r_bias", "h.0.dec_attn.r_net.weight", "h.1.dec_attn.qkv_net.weight",
             "h.0.pos_ff.CoreNet.0.bias", "h.1.pos_ff.CoreNet.2.weight", "h.0.dec_attn.layer_norm.weight",
             "word_embedding.weight", "rl_local_timestep_embedding.weight",
             "vision_encoder.patch_embeddings.conv1.weight", "vision_encoder.patch_embeddings.residual_path.2.weight",
             "vision_encoder.patch_embeddings.residual_path.3.bias", "vision_encoder.patch_embeddings.projection.bias",
             "vision_encoder.row_position_embeddings.weight"]


def gen_model(name, cfg, seed, with_images, L):
    from src.model import TransformerXL
    torch.manual_seed(0)
    model = TransformerXL(ref_namespace(cfg))
    sd = orc.synth_state_dict(cfg, seed=seed)
    missing = model.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    assert set(model.state_dict().keys()) == set(orc.state_shapes(cfg).keys())
    for k, v in model.state_dict().items():
        assert tuple(v.shape) == tuple(orc.state_shapes(cfg)[k]), k
    model.eval()
    tasks = make_tasks(cfg, seed + 100, with_images, L)
    logits, loss = model(to_ref_inputs(tasks))
    loss.backward()
    grads = dict(model.named_parameters())
    out = {"loss": np.array([loss.item()], dtype=np.float64),
           "logits_sub": logits.detach()[:, ::7, ::13].numpy().copy(),
           "logits_absmax": np.array([logits.abs().max().item()]),
           "logits_sum": np.array([logits.double().sum().item()])}
    for k in GRAD_KEYS:
        gk = grads[k].grad
        if gk is None:
            continue
        flat = gk.reshape(-1)
        stride = max(1, flat.numel() // 4096)
        out["grad:" + k] = flat[::stride].numpy().copy()
        out["gradnorm:" + k] = np.array([gk.double().norm().item()])
    for ti, t in enumerate(tasks):
        for k, v in t.items():
            if k == "type":
                out["task%d:type" % ti] = np.array(t["type"])
            elif v is not None:
                out["task%d:%s" % (ti, k)] = np.asarray(v)
    out["cfg"] = np.array(repr(vars(cfg)))
    out["seed"] = np.array([seed])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "loss", loss.item(), "logits", tuple(logits.shape))


def gen_attn_layer():
    """One decoder layer with memory, an active sliding window (mem_len < klen) and an active distance clamp."""
    from src.model.transformer_xl import RelPartialLearnableDecoderLayer, PositionalEmbedding
    cfg = orc.tiny_config(n_position=40, mem_len=48, text_vocab_size=480)
    sd = orc.synth_state_dict(cfg, seed=3)
    d, H = cfg.n_embed, cfg.n_head
    layer = RelPartialLearnableDecoderLayer(H, d, d // H, cfg.n_inner, 0.0, dropatt=0.0, activation="geglu",
                                            pre_lnorm=False, r_w_bias=torch.nn.Parameter(sd["r_w_bias"].clone()),
                                            r_r_bias=torch.nn.Parameter(sd["r_r_bias"].clone()),
                                            layer_norm_epsilon=cfg.layer_norm_epsilon).eval()
    lsd = {k[len("h.0."):]: v for k, v in sd.items() if k.startswith("h.0.")}
    layer.load_state_dict(lsd, strict=True)
    g = torch.Generator().manual_seed(11)
    out = {}
    for n, (Q, Mlen) in enumerate([(96, 0), (24, 40)]):
        K = Q + Mlen
        x = torch.randn(2, Q, d, generator=g)
        mem = torch.randn(2, Mlen, d, generator=g) if Mlen else None
        all_ones = x.new_ones((Q, K), dtype=torch.uint8)
        mask_len = K - cfg.mem_len
        shift = Q - mask_len if mask_len > 0 else Q
        mask = (torch.triu(all_ones, 1 + Mlen) + torch.tril(all_ones, -shift))[None]
        pos = torch.arange(K - 1, -1, -1.0).clamp_(max=cfg.n_position)
        pe = PositionalEmbedding(d)(pos)
        with torch.no_grad():
            y = layer(x, pe, attention_mask=mask, mems=mem)[0]
        out["x%d" % n] = x.numpy()
        if mem is not None:
            out["mem%d" % n] = mem.numpy()
        out["y%d" % n] = y.numpy()
        out["mask%d" % n] = mask[0].numpy()
        out["pe%d" % n] = pe[0].numpy()
    np.savez_compressed(os.path.join(OUT, "layer.npz"), **out)


def gen_rl_assemble():
    """RLFullDataset.get (rl_dataset.py:614-752) + postprocess_obs_and_act (:393-473) on raw arrays, run UNBOUND on a
    stand-in `self` that carries exactly the attributes those two methods read (the real constructor needs d4rl / gym
    environments that do not exist here). Cases: continuous control inside an episode; a window that runs past the
    episode end (action flags zeroed, padding); image frames + discrete actions with fewer transitions than
    transition_num (frame padding, -1 fill of the padded transitions' observation slots); dict observation."""
    import types as _t
    from src.data import rl_dataset as rd
    from src.tokenizer.scalar_tokenizer import ContinuousScalarTokenizer

    def map_structure(f, *xs):  # dm-tree is absent: the two shapes the reference uses (array or flat dict of arrays)
        if isinstance(xs[0], dict):
            return {k: f(*[x[k] for x in xs]) for k in xs[0]}
        return f(*xs)
    rd.tree.map_structure = map_structure
    rng = np.random.default_rng(5)
    out = {}

    def run(n, obs, act, obs_types, obs_dims, path_length, start, end, L, transition_num, overlap, obs_dim, act_dim):
        fake = _t.SimpleNamespace()
        fake.indices = np.array([[0, start, end]])
        fake.path_lengths = [path_length]
        fake.discretizer = ContinuousScalarTokenizer(1024)
        fake.num_discrete_values = 1024
        fake.use_prompt = False
        fake.vision_patch_size = 16
        fake.transition_num = transition_num
        fake.text_tokenizer = _t.SimpleNamespace(vocab_size=32000)
        fake.overlap_with_text = overlap
        fake.observation_dim = obs_dim
        fake.action_dim = act_dim
        fake.output_sequence_length = L
        fake.env = object()
        fake.obs_type_spec = obs_types
        fake.observation_dims_for_spec = obs_dims
        sl = lambda x: x[start:min(end, path_length)]  # noqa: E731
        fake.get_obs_action_by_path_idx = lambda p, s_, e_: (map_structure(sl, obs), act[start:min(end, path_length)])
        fake.postprocess_obs_and_act = _t.MethodType(rd.RLFullDataset.postprocess_obs_and_act, fake)
        res = rd.RLFullDataset.get(fake, 0)
        out["meta%d" % n] = np.array([path_length, start, end, L, transition_num, int(overlap), obs_dim, act_dim])
        for k in ("tensor_seq", "label", "loss_mask", "position_id"):
            out["%s%d" % (k, n)] = getattr(res, k)[0].numpy()
        if res.vision_seq is not None:
            out["vision_shape%d" % n] = np.array(res.vision_seq.shape)
        return res

    # 0: continuous control, window inside the episode (obs 17 floats, act 6 floats), truncated to L
    o = rng.standard_normal((60, 17)).astype(np.float32) * 3
    a = rng.uniform(-1, 1, (60, 6)).astype(np.float32)
    out["obs0"], out["act0"] = o, a
    run(0, o, a, "float", 17, 60, 5, 5 + 43, 1024, 43, True, 17, 6)
    # 1: window past the episode end (end_ind > path_length): flags zeroed from the episode end on, zero padding
    out["obs1"], out["act1"] = o, a
    run(1, o, a, "float", 17, 60, 40, 40 + 43, 1024, 43, False, 17, 6)
    # 2: frames + discrete actions, fewer transitions than transition_num -> frame padding and -1 fill
    img = rng.random((30, 3, 32, 48)).astype(np.float32)
    ad = rng.integers(0, 18, size=(30,)).astype(np.int64)
    out["img_shape2"], out["act2"] = np.array(img.shape), ad
    run(2, img, ad, "image", 6, 30, 22, 22 + 12, 128, 12, True, 6, 1)
    # 3: dict observation {image, float state}
    st = rng.standard_normal((30, 4)).astype(np.float32)
    out["img_shape3"], out["state3"], out["act3"] = np.array(img.shape), st, ad
    run(3, {"a_img": img, "b_state": st}, ad, {"a_img": "image", "b_state": "float"}, {"a_img": 6, "b_state": 4}, 30, 3, 3 + 9,
        128, 9, False, 10, 1)
    np.savez_compressed(os.path.join(OUT, "rl_assemble.npz"), **out)


def gen_collate():
    """my_collate_fn (data_samplers.py:28-42) on a shuffled list of single-sample task objects of three types."""
    from src.data.data_samplers import my_collate_fn
    from src.data.input_specs import ICTaskInput, NLPTaskInput, RLTaskInput
    g = torch.Generator().manual_seed(8)
    L = 16

    def rl(i, with_img):
        return RLTaskInput(position_id=torch.randint(0, 5, (1, L), generator=g), attention_mask=None,
                           loss_mask=torch.randint(0, 2, (1, L), generator=g).float(), label=torch.randint(0, 99, (1, L), generator=g),
                           text_seq=None, vision_seq=torch.rand(1, 2, 3, 16, 16, generator=g) if with_img else None,
                           tensor_seq=torch.randint(0, 99, (1, L), generator=g))

    def nlp(i):
        return NLPTaskInput(position_id=None, attention_mask=None, loss_mask=torch.ones(1, L), label=torch.randint(0, 99, (1, L), generator=g),
                            text_seq=torch.randint(0, 99, (1, L), generator=g), text_len=None)

    def ic(i):
        return ICTaskInput(position_id=None, attention_mask=None, loss_mask=torch.ones(1, L), label=torch.randint(0, 99, (1, L), generator=g),
                           prompt_seq=torch.randint(0, 99, (1, 3), generator=g), img_seq=torch.rand(1, 3, 16, 16, generator=g),
                           text_seq=torch.randint(0, 99, (1, L - 4), generator=g), img_id_seq=None)
    order = ["nlp", "rl", "ic", "rl", "nlp", "nlp", "ic", "rl"]
    samples = [dict(nlp=nlp, rl=lambda i: rl(i, True), ic=ic)[t](i) for i, t in enumerate(order)]
    out = {"order": np.array(order)}
    from dataclasses import fields
    for i, smp in enumerate(samples):
        for f in fields(smp):
            v = getattr(smp, f.name)
            if isinstance(v, torch.Tensor):
                out["in%d:%s" % (i, f.name)] = v.numpy()
    import copy
    merged = my_collate_fn([copy.deepcopy(x) for x in samples])
    out["out_types"] = np.array([type(m).__name__ for m in merged])
    for gi, m in enumerate(merged):
        for f in fields(m):
            v = getattr(m, f.name)
            if isinstance(v, torch.Tensor):
                out["out%d:%s" % (gi, f.name)] = v.numpy()
    np.savez_compressed(os.path.join(OUT, "collate.npz"), **out)


def gen_posemb_half():
    """PositionalEmbedding as the reference evaluates it after module.half() (DeepSpeed fp16): fp16 pos_seq (:569-571),
    fp16-cast inv_freq buffer (:44). Full table for demb 128; a strided sample of rows / columns for demb 2048."""
    from src.model.transformer_xl import PositionalEmbedding
    out = {}
    for n, (klen, demb, clamp, rs, cs) in enumerate([(300, 128, 256, 1, 1), (1024, 2048, 1024, 37, 29),
                                                      (2048, 2048, 1024, 61, 31)]):
        pe = PositionalEmbedding(demb).half()
        pos = torch.arange(klen - 1, -1, -1.0, dtype=torch.float16)
        pos.clamp_(max=clamp)
        t = pe(pos)[0]
        assert t.dtype == torch.float16
        out["case%d" % n] = np.array([klen, demb, clamp, rs, cs])
        out["rows%d" % n] = t[::rs, ::cs].numpy()
    np.savez_compressed(os.path.join(OUT, "posemb_half.npz"), **out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if "--only-new" in sys.argv:  # fixtures added in round 2 (the others are unchanged)
        gen_posemb_half()
        gen_rl_assemble()
        gen_collate()
        raise SystemExit(0)
    gen_posemb_half()
    gen_rl_assemble()
    gen_collate()
    gen_tokenizer()
    gen_rl_layout()
    gen_patch_positions()
    gen_attn_layer()
    gen_model("tiny_text_rl", orc.tiny_config(text_vocab_size=480), seed=1, with_images=False, L=256)
    gen_model("tiny_mixed_images", orc.tiny_config(text_vocab_size=480), seed=2, with_images=True, L=128)
    gen_model("tiny_window_clamp", orc.tiny_config(text_vocab_size=480, n_position=100, mem_len=160), seed=4,
              with_images=False, L=256)
    print("golden fixtures written to", OUT)
