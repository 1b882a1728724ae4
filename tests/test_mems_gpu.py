"""GPU parity of memory-augmented inference (`forward(..., compute_loss=False, mems=...)`, transformer_xl.py:124-133,
:470-504, :615-619; the path src/evaluation/evaluate_rl.py:157-266 drives) against the oracle: logits of the new
tokens within 2e-3 of max (fp16 vs fp32), new memories = the reference's concatenate-and-slice of layer inputs."""
import numpy as np
import pytest
import torch

from tests import util
from tests.test_model_gpu import _build

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,mlen,qlen", [(1, 256, 1), (2, 256, 5), (1, 100, 37), (2, 256, 128), (1, 0, 16)])
def test_forward_with_mems_matches_oracle(cuda, B, mlen, qlen):
    from oracle import db1_oracle as orc
    cfg = orc.tiny_config(text_vocab_size=480)  # mem_len = n_position = 256, same_length
    model, sd = _build(cfg, 11, cuda)
    model.eval()
    g = torch.Generator().manual_seed(3)
    tok = torch.randint(0, 480, (B, qlen), generator=g)
    task = dict(type="nlp", text_seq=tok.numpy(), label=np.zeros((B, qlen), np.int64), loss_mask=np.ones((B, qlen), np.float32))
    mems32 = [torch.randn(B, mlen, cfg.n_embed, generator=g).half().float() for _ in range(cfg.n_layer)]
    sdo = {k: v.clone() for k, v in sd.items()}
    for k in list(sdo):
        if k.startswith("h.") and k.endswith(("r_r_bias", "r_w_bias")):
            sdo[k] = sdo[k.split(".")[-1]]
    with torch.no_grad():
        ologits, oloss, omems = orc.forward([task], sdo, cfg, compute_loss=False, mems=mems32)
        out = model(util.to_model_inputs([task], cuda), compute_loss=False, mems=[m.half().to(cuda) for m in mems32])
    assert len(out) == 3 and out[1] is None
    logits, _, new_mems = out
    assert logits.shape == (B, qlen, ologits.shape[-1])
    assert util.rel_err(logits, ologits) <= 2e-3
    assert len(new_mems) == cfg.n_layer
    for nm, om in zip(new_mems, omems):
        assert nm.shape == om.shape
        assert util.rel_err(nm, om) <= 2e-3


def test_init_mem_and_stepwise_decode_equals_one_shot(cuda):
    """Feeding a sequence token by token through the memory path gives the logits of the one-shot causal forward
    (same_length window not reached: 24 tokens << mem_len), starting from zero-length memories."""
    from oracle import db1_oracle as orc
    cfg = orc.tiny_config(text_vocab_size=480)
    model, _sd = _build(cfg, 12, cuda)
    model.eval()
    L = 24
    tok = torch.randint(0, 480, (1, L), generator=torch.Generator().manual_seed(5))
    full = dict(type="nlp", text_seq=tok.numpy(), label=np.zeros((1, L), np.int64), loss_mask=np.ones((1, L), np.float32))
    with torch.no_grad():
        ref_logits, _ = model(util.to_model_inputs([full], cuda), compute_loss=False)
        zero = model.init_mem(1)
        assert len(zero) == cfg.n_layer and zero[0].shape == (1, cfg.mem_len, cfg.n_embed)
        mems = [m[:, :0] for m in zero]  # empty memories: pure incremental decoding
        outs = []
        for t in range(0, L, 8):
            piece = dict(type="nlp", text_seq=tok[:, t:t + 8].numpy(), label=np.zeros((1, 8), np.int64),
                         loss_mask=np.ones((1, 8), np.float32))
            lg, _, mems = model(util.to_model_inputs([piece], cuda), compute_loss=False, mems=mems)
            outs.append(lg)
    step_logits = torch.cat(outs, 1)
    assert util.rel_err(step_logits, ref_logits) <= 3e-3
    assert mems[0].shape == (1, L, cfg.n_embed)


@pytest.mark.parametrize("B", [1, 2])
def test_kv_cached_decode_matches_recompute_path_and_oracle(cuda, B):
    """init_mem(kv_cache=True) + forward(..., mems=<KVMemory>) - the decode loop of evaluate_rl.py:157-266
    (a transition's observation tokens, then one action token at a time) on cached keys / values - gives, step by step,
    the logits of the reference-format memory path (full re-projection of cat(mem, w)) and of the oracle, and ends with
    the same memories. Starts from init_mem's zero memories (which ARE attended), runs long enough for the ring buffer to
    wrap; the masked arg-max (evaluate_rl.py:96-138, :196-199) is checked on the same logits."""
    from oracle import db1_oracle as orc
    from db1_sm100 import ops
    cfg = orc.tiny_config(text_vocab_size=480, mem_len=64, n_position=64)
    model, sd = _build(cfg, 13, cuda)
    model.eval()
    sdo = {k: v.clone() for k, v in sd.items()}
    for k in list(sdo):
        if k.startswith("h.") and k.endswith(("r_r_bias", "r_w_bias")):
            sdo[k] = sdo[k.split(".")[-1]]
    g = torch.Generator().manual_seed(7)
    V = orc.total_vocab(cfg)
    kv = model.init_mem(B, kv_cache=True)
    mems = model.init_mem(B)
    omems = [m.float().cpu() for m in mems]
    qlens = [9, 1, 1, 1, 12, 1, 1, 30, 1, 17, 1, 1]  # 76 rows > mem_len = 64: the ring wraps
    for step, q in enumerate(qlens):
        tok = torch.randint(0, 480, (B, q), generator=g)
        task = dict(type="nlp", text_seq=tok.numpy(), label=np.zeros((B, q), np.int64), loss_mask=np.ones((B, q), np.float32))
        with torch.no_grad():
            lg_kv, none, kv2 = model(util.to_model_inputs([task], cuda), compute_loss=False, mems=kv)
            lg_re, _, mems = model(util.to_model_inputs([task], cuda), compute_loss=False, mems=mems)
            ol, _, omems = orc.forward([task], sdo, cfg, compute_loss=False, mems=omems)
        assert kv2 is kv and none is None
        e_kv, e_re = util.rel_err(lg_kv, ol), util.rel_err(lg_re, ol)
        print("step %d q %d: cached vs oracle %.2e, recompute path vs oracle %.2e" % (step, q, e_kv, e_re))
        assert e_kv <= 3e-3, step
        assert util.rel_err(lg_kv, lg_re) <= 2e-3, step
        # continuous-action head: arg-max over [text_vocab, V - 1) (everything else -1e10, :107-112), last position
        last = lg_kv[:, -1, :]
        buf = last.contiguous()
        pred = ops.masked_argmax(buf, cfg.text_vocab_size, V - 1)
        ref = last.float().clone()
        ref[:, :cfg.text_vocab_size] -= 1e10
        ref[:, -1] -= 1e10
        assert torch.equal(pred, ref.argmax(-1))
        # discrete head with an environment action mask (:113-123)
        n_act = 18
        amask = (torch.rand(n_act, generator=g) < 0.5).float()
        amask[3] = 1.0
        pen = ((amask - 1).abs() * 1e10).to(cuda)
        pred_d = ops.masked_argmax(buf, 0, n_act, add_mask=pen)
        ref_d = last.float().clone()
        ref_d[:, n_act:] -= 1e10
        ref_d[:, :n_act] -= pen
        assert torch.equal(pred_d, ref_d.argmax(-1))
    for a, b_, c in zip(kv.to_mems(), mems, omems):
        assert a.shape == b_.shape
        assert util.rel_err(a, b_) <= 2e-3
        assert util.rel_err(a, c) <= 3e-3


def test_decode_graph_replays_the_cached_step(cuda):
    """DecodeGraph (one cached decode step captured in a CUDA graph; the ring head lives in device memory) gives, replay
    after replay, what the same steps give when launched one kernel at a time, including across the ring wrap, and leaves
    the same memories behind."""
    from oracle import db1_oracle as orc
    from db1_sm100.functions import DecodeGraph
    from src.data.input_specs import RLTaskInput
    cfg = orc.tiny_config(text_vocab_size=480, mem_len=64, n_position=64)
    model, _sd = _build(cfg, 21, cuda)
    model.eval()
    B = 2
    g = torch.Generator().manual_seed(3)
    kv_a = model.init_mem(B, kv_cache=True)
    kv_b = model.init_mem(B, kv_cache=True)
    # a few eager steps first, so that the graph is captured on a memory whose head is not at slot 0
    for _ in range(5):
        tok = torch.randint(0, 480, (B, 1), generator=g).to(cuda)
        for kv in (kv_a, kv_b):
            inp = [RLTaskInput(position_id=torch.zeros(B, 1, dtype=torch.int64, device=cuda), attention_mask=None,
                               loss_mask=None, label=None, text_seq=None, vision_seq=None, tensor_seq=tok)]
            with torch.no_grad():
                model(inp, compute_loss=False, mems=kv)
    graph = DecodeGraph(model, kv_a, 1)
    assert kv_a.head == kv_b.head == 5
    for a, b_ in zip(kv_a.to_mems(), kv_b.to_mems()):
        assert torch.equal(a, b_)  # capture and warm-up left no trace
    for step in range(70):  # 5 + 70 > mem_len: the ring wraps under the graph
        tok = torch.randint(0, 480, (B, 1), generator=g).to(cuda)
        pos = torch.full((B, 1), step % 7, dtype=torch.int64, device=cuda)
        lg_g = graph.step(tok, pos).clone()
        inp = [RLTaskInput(position_id=pos, attention_mask=None, loss_mask=None, label=None, text_seq=None,
                           vision_seq=None, tensor_seq=tok)]
        with torch.no_grad():
            lg_e, _, _ = model(inp, compute_loss=False, mems=kv_b)
        assert util.rel_err(lg_g, lg_e) <= 1e-6, step
    assert kv_a.head == kv_b.head and int(kv_a.head_dev.item()) == kv_a.head
    for a, b_ in zip(kv_a.to_mems(), kv_b.to_mems()):
        assert torch.equal(a, b_)
