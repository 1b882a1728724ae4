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
