"""CPU: pin the oracle restatement (oracle/db1_oracle.py) to the reference's own outputs (tests/golden/*.npz, produced
by tools/make_golden.py from the unmodified reference) and check the host-side integer paths of the product."""
import ast
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import db1_oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False)


def test_oracle_discretize_matches_reference_bit_exact():
    g = _load("tokenizer")
    assert np.array_equal(orc.discretize(g["obs"], False), g["obs_tok"])
    assert np.array_equal(orc.discretize(g["act"], True), g["act_tok"])
    # the known answers recorded in SURVEY.md section 8 a12
    assert g["obs_tok"][:10].tolist() == [0, 279, 313, 507, 512, 516, 710, 744, 1023, 1023]
    assert g["act_tok"][:11].tolist() == [0, 0, 0, 256, 512, 512, 768, 1022, 1023, 1023, 1023]
    np.testing.assert_array_equal(orc.decode(g["toks"], False), g["dec_obs"])
    np.testing.assert_array_equal(orc.decode(g["toks"], True), g["dec_act"])


def test_host_library_discretize_bit_exact():
    """The product's C++ tokenizer (libdb1_host.so) against the reference's golden vectors."""
    from src.tokenizer.scalar_tokenizer import ContinuousScalarTokenizer
    g = _load("tokenizer")
    tk = ContinuousScalarTokenizer()
    got = tk.discretize(g["obs"], is_action=False).numpy()
    bad = np.nonzero(got != g["obs_tok"])[0]
    assert bad.size == 0, "mismatching inputs: %s" % g["obs"][bad][:10]
    assert np.array_equal(tk.discretize(g["act"], is_action=True).numpy(), g["act_tok"])
    assert np.array_equal(tk.decode(g["toks"], is_action=True).numpy(), g["dec_act"])
    np.testing.assert_allclose(tk.decode(g["toks"], is_action=False).numpy(), g["dec_obs"], rtol=2e-6, atol=0)


def test_oracle_rl_layout_matches_reference():
    g = _load("rl_layout")
    n = 0
    while "case%d" % n in g:
        seq, ol, al, pre = g["case%d" % n].tolist()
        flag, pos = orc.action_flag_and_position_id(seq, ol, al, pre)
        assert np.array_equal(flag, g["flag%d" % n]) and np.array_equal(pos, g["pos%d" % n])
        n += 1
    assert n >= 6
    # SURVEY.md a13 known answer
    flag, pos = orc.action_flag_and_position_id(19, 4, 3, 1)
    assert "".join(map(str, flag)) == "0000000000000111000" and "".join(map(str, pos)) == "1234500012345000123"
    for k in range(3):
        x, tgt = g["pad_in%d" % k], int(g["pad_tgt%d" % k][0])
        y = x[:tgt] if len(x) >= tgt else np.concatenate([x, np.zeros(tgt - len(x), dtype=x.dtype)])
        assert np.array_equal(y, g["pad_out%d" % k])


def test_host_library_rl_layout_matches_oracle():
    import ctypes as C
    from db1_sm100 import _lib
    rng = np.random.default_rng(0)
    for (T, ol, al, L, pre) in [(3, 4, 3, 18, 1), (43, 17, 6, 1024, 0), (5, 6, 1, 64, 2), (2, 1, 1, 3, 0)]:
        obs = rng.integers(-1, 2000, size=(T, ol)).astype(np.int64)
        act = rng.integers(0, 2000, size=(T, al)).astype(np.int64)
        ref = orc.rl_sequence(obs, act, 33024, L, prepend_trans_num=pre)
        ts = np.empty(L, np.int64); lb = np.empty(L, np.int64); lm = np.empty(L, np.float32); ps = np.empty(L, np.int64)
        p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        rc = _lib.hostlib().db1_rl_layout(p(obs), p(act), T, ol, al, C.c_longlong(33024), L, C.c_longlong(0), pre,
                                          p(ts), p(lb), p(lm), p(ps))
        assert rc == 0
        assert np.array_equal(ts, ref["tensor_seq"]) and np.array_equal(lb, ref["label"])
        assert np.array_equal(lm, ref["loss_mask"]) and np.array_equal(ps, ref["position_id"])


def test_host_library_rl_sample_idx_known_answer():
    """helpers.cpp:82-115 known answer measured on the reference build."""
    import ctypes as C
    from db1_sm100 import _lib
    lens = np.array([5, 3, 1, 7], dtype=np.int32)
    f = _lib.hostlib().db1_build_rl_sample_idx
    f.restype = C.c_longlong
    n = f(lens.ctypes.data_as(C.c_void_p), C.c_longlong(4), 3, None, C.c_longlong(0))
    assert n == 12
    out = np.empty((n, 3), dtype=np.int32)
    assert f(lens.ctypes.data_as(C.c_void_p), C.c_longlong(4), 3, out.ctypes.data_as(C.c_void_p), C.c_longlong(n)) == 12
    assert out[:7].tolist() == [[0, 0, 3], [0, 1, 4], [0, 2, 5], [0, 3, 5], [1, 0, 3], [1, 1, 3], [3, 0, 3]]
    assert out[-1].tolist() == [3, 5, 7]


def test_oracle_patch_positions_match_reference():
    g = _load("patch_positions")
    n = 0
    while "grid%d" % n in g:
        h0, w0 = g["grid%d" % n].tolist()
        row, col = orc.patch_position_indices(h0, w0)[:2]
        assert np.array_equal(row, g["row%d" % n]) and np.array_equal(col, g["col%d" % n])
        n += 1
    assert orc.patch_position_indices(5, 5)[1][:5].tolist() == [12, 38, 63, 89, 115]


def test_oracle_layer_with_memory_window_and_clamp():
    g = _load("layer")
    cfg = orc.tiny_config(n_position=40, mem_len=48, text_vocab_size=480)
    sd = orc.synth_state_dict(cfg, seed=3)
    for n in range(2):
        x = torch.as_tensor(g["x%d" % n])
        mem = torch.as_tensor(g["mem%d" % n]) if ("mem%d" % n) in g else None
        Q = x.shape[1]
        K = Q + (mem.shape[1] if mem is not None else 0)
        ok = orc.attention_mask_ok(Q, K, cfg.mem_len, True)
        assert np.array_equal((~ok).numpy().astype(np.uint8), (g["mask%d" % n] > 0).astype(np.uint8))
        pe = orc.positional_rows(K, cfg.n_embed, cfg.n_position)
        np.testing.assert_allclose(pe.numpy(), g["pe%d" % n], atol=1e-6)
        y = orc.decoder_layer(x, pe, sd, "h.0.", cfg, ok, mem)
        np.testing.assert_allclose(y.numpy(), g["y%d" % n], atol=2e-5, rtol=1e-5)


def _tasks_from_golden(g):
    tasks = []
    ti = 0
    while "task%d:type" % ti in g:
        t = {"type": str(g["task%d:type" % ti])}
        for k in g.files:
            pre = "task%d:" % ti
            if k.startswith(pre) and k != pre + "type":
                t[k[len(pre):]] = g[k]
        t.setdefault("vision_seq", None)
        tasks.append(t)
        ti += 1
    return tasks


@pytest.mark.parametrize("name", ["tiny_text_rl", "tiny_mixed_images", "tiny_window_clamp"])
def test_oracle_model_forward_backward_matches_reference(name):
    g = _load(name)
    cfg = SimpleNamespace(**ast.literal_eval(str(g["cfg"])))
    sd = {k: v.clone().requires_grad_(v.is_floating_point() and k != "pos_emb.inv_freq")
          for k, v in orc.synth_state_dict(cfg, seed=int(g["seed"][0])).items()}
    # aliases must stay aliases for the gradient to accumulate the way the reference's shared parameters do
    for k in list(sd):
        if k.startswith("ic_encoder."):
            sd[k] = sd["vision_encoder." + k[len("ic_encoder."):]]
        if k.startswith("h.") and k.endswith(("r_r_bias", "r_w_bias")):
            sd[k] = sd[k.split(".")[-1]]
    logits, loss = orc.forward(_tasks_from_golden(g), sd, cfg)
    assert abs(loss.item() - g["loss"][0]) < 2e-6 * abs(g["loss"][0])
    np.testing.assert_allclose(logits.detach()[:, ::7, ::13].numpy(), g["logits_sub"], atol=3e-6, rtol=1e-5)
    assert abs(logits.double().sum().item() - g["logits_sum"][0]) < 1e-3 * max(1.0, abs(g["logits_sum"][0]))
    loss.backward()
    checked = 0
    for k in g.files:
        if not k.startswith("grad:"):
            continue
        p = sd[k[5:]]
        assert p.grad is not None, k
        flat = p.grad.reshape(-1)
        stride = max(1, flat.numel() // 4096)
        ref = g[k]
        np.testing.assert_allclose(flat[::stride].numpy(), ref, atol=1e-6 + 2e-4 * np.abs(ref).max(), rtol=0)
        assert abs(p.grad.double().norm().item() - g["gradnorm:" + k[5:]][0]) < 1e-4 * g["gradnorm:" + k[5:]][0] + 1e-9
        checked += 1
    assert checked >= 8


def _rl_idx(lens, T):
    import ctypes as C
    from db1_sm100 import _lib
    f = _lib.hostlib().db1_build_rl_sample_idx
    f.restype = C.c_longlong
    n = f(lens.ctypes.data_as(C.c_void_p), C.c_longlong(len(lens)), int(T), None, C.c_longlong(0))
    out = np.empty((n, 3), dtype=np.int32)
    assert f(lens.ctypes.data_as(C.c_void_p), C.c_longlong(len(lens)), int(T), out.ctypes.data_as(C.c_void_p), C.c_longlong(n)) == n
    return out


def _gpt_idx(sizes, doc_idx, seq, ep, tpe):
    import ctypes as C
    from db1_sm100 import _lib
    f = _lib.hostlib().db1_build_sample_idx
    f.restype = C.c_longlong
    args = (sizes.ctypes.data_as(C.c_void_p), doc_idx.ctypes.data_as(C.c_void_p), C.c_longlong(len(doc_idx)), int(seq),
            int(ep), C.c_longlong(int(tpe)))
    n = f(*args, None, C.c_longlong(0))
    out = np.empty((n, 2), dtype=np.int32)
    assert f(*args, out.ctypes.data_as(C.c_void_p), C.c_longlong(n)) == n
    return out


def test_host_index_builders_match_reference_golden():
    """libdb1_host.so vs vectors produced by the reference's compiled helpers.cpp (tools/make_golden_index.py): bit-exact."""
    g = _load("index_builders")
    i = 0
    while "rl%d:lens" % i in g:
        assert np.array_equal(_rl_idx(g["rl%d:lens" % i], g["rl%d:T" % i][0]), g["rl%d:idx" % i]), i
        i += 1
    assert i >= 5
    i = 0
    while "gpt%d:sizes" % i in g:
        seq, ep, tpe = g["gpt%d:args" % i]
        assert np.array_equal(_gpt_idx(g["gpt%d:sizes" % i], g["gpt%d:doc_idx" % i], seq, ep, tpe), g["gpt%d:idx" % i]), i
        i += 1
    assert i >= 5


def test_host_index_builders_match_live_reference_when_built():
    """Randomised comparison against oracle/_ref (the reference's helpers.cpp compiled in the build container)."""
    import contextlib
    import glob
    import io
    import sys
    ref_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
    if not glob.glob(os.path.join(ref_dir, "helpers*.so")):
        pytest.skip("oracle/_ref not built (make -C oracle; needs /root/reference)")
    sys.path.insert(0, ref_dir)
    import helpers
    rng = np.random.default_rng(5)
    for _ in range(20):
        lens = rng.integers(1, 60, size=int(rng.integers(1, 40))).astype(np.int32)
        T = int(rng.integers(1, 50))
        assert np.array_equal(_rl_idx(lens, T), np.asarray(helpers.build_rl_sample_idx(lens, T)))
        ndoc, ep, seq = int(rng.integers(1, 50)), int(rng.integers(1, 4)), int(rng.integers(2, 300))
        sizes = rng.integers(1, 400, size=ndoc).astype(np.int32)
        doc_idx = np.concatenate([rng.permutation(ndoc) for _ in range(ep)]).astype(np.int32)
        if ep * int(sizes.sum()) - 1 < seq:
            continue
        with contextlib.redirect_stdout(io.StringIO()):
            ref = np.asarray(helpers.build_sample_idx(sizes, doc_idx, seq, ep, int(sizes.sum())))
        assert np.array_equal(_gpt_idx(sizes, doc_idx, seq, ep, int(sizes.sum())), ref)


def test_positional_rows_half_phase_match_reference_after_half():
    """oracle.positional_rows(half_phase=True) == the reference's PositionalEmbedding after module.half() (bit-exact:
    fp16 positions, fp16-cast inv_freq, fp16 product, fp16 sin / cos); the fp32 phases differ from it by percents on
    far rows, which is why the B200 module defaults to the half-phase mode when its buffers are fp16."""
    g = _load("posemb_half")
    n = 0
    worst = 0.0
    while "case%d" % n in g:
        klen, demb, clamp, rs, cs = [int(x) for x in g["case%d" % n]]
        got = orc.positional_rows(klen, demb, clamp, half_phase=True)[::rs, ::cs]
        ref = torch.from_numpy(g["rows%d" % n].astype(np.float32))
        assert torch.equal(got, ref), "case %d" % n
        f32 = orc.positional_rows(klen, demb, clamp)[::rs, ::cs]
        worst = max(worst, ((f32 - ref).norm() / ref.norm()).item())
        n += 1
    assert n == 3 and worst > 1e-2  # the two modes really are different tables


def test_host_rl_assemble_matches_reference_get_bit_exact():
    """libdb1_host.so:db1_rl_assemble vs RLFullDataset.get of the unmodified reference (tests/golden/rl_assemble.npz, made by
    tools/make_golden.py:gen_rl_assemble): continuous control, a window past the episode end, frames + discrete actions
    with frame padding (-1 fill), dict observation {image, float}; both vocabulary-overlap settings."""
    from db1_sm100 import host
    g = _load("rl_assemble")

    def check(n, res):
        for k in ("tensor_seq", "label", "position_id"):
            assert np.array_equal(getattr(res, k)[0].numpy(), g["%s%d" % (k, n)]), (n, k)
        assert np.array_equal(res.loss_mask[0].numpy(), g["loss_mask%d" % n]), n
    for n in (0, 1):
        pl, start, end, L, tn, ov, od, ad = [int(x) for x in g["meta%d" % n]]
        e = min(end, pl)
        check(n, host.rl_assemble(g["obs%d" % n][start:e], g["act%d" % n][start:e], L, overlap_with_text=bool(ov)))
    pl, start, end, L, tn, ov, od, ad = [int(x) for x in g["meta2"]]
    e = min(end, pl)
    frames = np.zeros(tuple(g["img_shape2"]), np.float32)
    res = host.rl_assemble(None, g["act2"][start:e], L, frames=frames[start:e], transition_num=tn, overlap_with_text=bool(ov))
    check(2, res)
    assert tuple(res.vision_seq.shape) == tuple(g["vision_shape2"])
    pl, start, end, L, tn, ov, od, ad = [int(x) for x in g["meta3"]]
    e = min(end, pl)
    frames = np.zeros(tuple(g["img_shape3"]), np.float32)
    res = host.rl_assemble(g["state3"][start:e], g["act3"][start:e], L, frames=frames[start:e], transition_num=tn,
                           overlap_with_text=bool(ov))
    check(3, res)
    assert tuple(res.vision_seq.shape) == tuple(g["vision_shape3"])


def test_host_collate_matches_reference_my_collate_fn():
    """db1_collate_plan + db1_concat_rows vs my_collate_fn of the unmodified reference (tests/golden/collate.npz): grouping by
    task type in order of first appearance, fields concatenated on dim 0, None fields stay None."""
    from db1_sm100 import host
    from src.data.input_specs import ICTaskInput, NLPTaskInput, RLTaskInput
    g = _load("collate")
    cls = {"nlp": NLPTaskInput, "rl": RLTaskInput, "ic": ICTaskInput}
    from dataclasses import fields
    samples = []
    for i, t in enumerate(g["order"]):
        kw = {f.name: (torch.from_numpy(g["in%d:%s" % (i, f.name)]) if "in%d:%s" % (i, f.name) in g else None)
              for f in fields(cls[str(t)])}
        samples.append(cls[str(t)](**kw))
    merged = host.collate(samples)
    assert [type(m).__name__ for m in merged] == [str(x) for x in g["out_types"]]
    for gi, m in enumerate(merged):
        for f in fields(m):
            v = getattr(m, f.name)
            key = "out%d:%s" % (gi, f.name)
            if key in g:
                assert isinstance(v, torch.Tensor) and v.dtype == torch.from_numpy(g[key]).dtype
                assert np.array_equal(v.numpy(), g[key]), key
            else:
                assert v is None, key
