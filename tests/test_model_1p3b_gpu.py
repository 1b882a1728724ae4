"""GPU parity at the configurations bench.py measures (BASELINE.json configs 2-4): DB1-1.3B width, L = 1024.

  * full depth (24 layers), one 1024-token RL sequence, forward AND backward against the fp32 CPU oracle;
  * B = 4 (the benchmarked micro-batch): every row of the B=4 forward equals the B=1 forward of that sequence;
  * 1.3B width, 2 layers (bounds the CPU oracle's time), config C3 (80x80 frames through the ResNet patch embedder with the
    d = 2048 projection, K = 16 384) and config C4 (RL + text + image-caption list), forward + backward against the oracle.
Tolerances: one layer / one op reproduces the oracle to fp16 rounding (asserted in test_model_gpu.py); chained through 24
post-LN layers the bound is the reference's OWN fp16-vs-fp32 drift, read from tests/golden/ref_fp16_drift.json (written by
tools/ref_fp16_drift.py from the unmodified reference) - not a literal in this file.
"""
import json
import os

import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu

LOSS_SCALE = 4096.0


def _drift():
    with open(os.path.join(util.GOLD, "ref_fp16_drift.json")) as f:
        return json.load(f)


def _oracle_sd(sd):
    sdo = {k: v.clone().requires_grad_(v.is_floating_point() and k != "pos_emb.inv_freq") for k, v in sd.items()}
    for k in list(sdo):
        if k.startswith("ic_encoder."):
            sdo[k] = sdo["vision_encoder." + k[len("ic_encoder."):]]
        if k.startswith("h.") and k.endswith(("r_r_bias", "r_w_bias")):
            sdo[k] = sdo[k.split(".")[-1]]
    return sdo


def _task_dict(t):
    """db1_sm100.synth task object -> the oracle's dict form."""
    name = type(t).__name__
    f = lambda x: None if x is None else x.cpu().float().numpy() if x.is_floating_point() else x.cpu().numpy()  # noqa: E731
    if name == "RLTaskInput":
        return dict(type="rl", tensor_seq=f(t.tensor_seq), label=f(t.label), loss_mask=f(t.loss_mask),
                    position_id=f(t.position_id), vision_seq=f(t.vision_seq))
    if name == "NLPTaskInput":
        return dict(type="nlp", text_seq=f(t.text_seq), label=f(t.label), loss_mask=f(t.loss_mask))
    return dict(type="ic", prompt_seq=f(t.prompt_seq), img_seq=f(t.img_seq), text_seq=f(t.text_seq), label=f(t.label),
                loss_mask=f(t.loss_mask))


def _model(cfg, seed, cuda):
    from oracle import db1_oracle as orc
    from src.model import TransformerXL
    sd = orc.synth_state_dict(cfg, seed=seed)
    sd = {k: (v.half().float() if v.is_floating_point() and k != "pos_emb.inv_freq" else v) for k, v in sd.items()}
    model = TransformerXL(cfg)
    model.load_state_dict(sd, strict=True)
    model = model.half().to(cuda).eval()
    model.pos_emb.phase_dtype = torch.float32  # fp32 oracle on the other side
    return model, sd


def _rms_rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt()).item()


def test_1p3b_full_depth_L1024_forward_backward_vs_oracle(cuda):
    """24 layers, d 2048, L = 1024 (the benchmarked sequence length), eval mode, loss scale 4096.
    Logits / loss: within the reference's own fp16 drift. Gradients (unscaled) of the first / a middle / the last layer's
    projections, the shared biases, the tied embedding and LayerNorm affine parameters: rel-L2 against fp32 autograd of
    the oracle, bounded by 3x the logits' rms drift bound (gradients cross the same 24 layers twice)."""
    from db1_sm100 import synth
    from oracle import db1_oracle as orc
    drift = _drift()
    cfg = orc.default_config()
    model, sd = _model(cfg, 21, cuda)
    L = 1024
    rl4 = synth.rl_continuous_batch(cfg, 4, L, seed=177)
    rows = [type(rl4)(position_id=rl4.position_id[b:b + 1], attention_mask=None, loss_mask=rl4.loss_mask[b:b + 1],
                      label=rl4.label[b:b + 1], text_seq=None, vision_seq=None, tensor_seq=rl4.tensor_seq[b:b + 1])
            for b in range(4)]
    # ---- B = 4 forward: row b == the B = 1 forward of sequence b (same kernels, other tile schedule)
    with torch.no_grad():
        l4, _ = model([synth.to_device(rl4, cuda)])
        for b in range(4):
            l1, _ = model([synth.to_device(rows[b], cuda)])
            assert util.rel_err(l4[b], l1[0]) <= 1e-3, b
    # ---- B = 1 forward + backward against the oracle
    logits, loss = model([synth.to_device(rows[0], cuda)])
    (loss * LOSS_SCALE).backward()
    torch.cuda.synchronize()
    sdo = _oracle_sd(sd)
    ologits, oloss = orc.forward([_task_dict(rows[0])], sdo, cfg)
    oloss.backward()
    err, rms = util.rel_err(logits, ologits), _rms_rel(logits, ologits.detach())
    print("1.3B L=1024 logits: max-norm rel %.3e rms rel %.3e (reference fp16 drift %.3e / %.3e); loss %.6f vs %.6f"
          % (err, rms, drift["max_norm_rel"], drift["rms_rel"], loss.item(), oloss.item()))
    assert err <= drift["max_norm_rel"] and rms <= drift["rms_rel"]
    assert abs(loss.item() - oloss.item()) <= 2e-3 * abs(oloss.item())
    names = ["r_w_bias", "r_r_bias", "word_embedding.weight", "rl_local_timestep_embedding.weight"]
    for li in (0, 11, 23):
        for sfx in ("dec_attn.qkv_net.weight", "dec_attn.r_net.weight", "dec_attn.o_net.weight", "pos_ff.CoreNet.0.weight",
                    "pos_ff.CoreNet.2.weight", "dec_attn.layer_norm.weight", "dec_attn.layer_norm.bias",
                    "pos_ff.layer_norm.weight", "pos_ff.layer_norm.bias", "pos_ff.CoreNet.0.bias", "pos_ff.CoreNet.2.bias"):
            names.append("h.%d.%s" % (li, sfx))
    params = dict(model.named_parameters())
    worst = {n: util.rel_l2(params[n].grad.float() / LOSS_SCALE, sdo[n].grad) for n in names}
    print("1.3B L=1024 gradient rel-L2 (worst 8):", sorted(worst.items(), key=lambda kv: -kv[1])[:8])
    bound = 3 * drift["rms_rel"]
    assert max(worst.values()) <= bound, {k: v for k, v in worst.items() if v > bound}


@pytest.mark.parametrize("workload", ["atari_C3", "mixed_C4"])
def test_1p3b_width_image_and_mixed_batches_vs_oracle(cuda, workload):
    """DB1-1.3B width (d 2048, 16 heads, GeGLU 8192), 2 layers, L = 1024:
    C3 = one sequence of 80x80 frames (25 patch slots + SEP + 1 discrete action per transition, 38 frames = 950 patches)
         through the ResNet patch embedder and its 16x16 -> 2048 projection (K = 16 384);
    C4 = the per-rank list of BASELINE config 4: RLTaskInput + NLPTaskInput + ICTaskInput (224x224 image, 196 patches).
    Forward and backward against the oracle at the DB1-tiny tolerances (2 layers: no deep-chain amplification)."""
    from db1_sm100 import synth
    from oracle import db1_oracle as orc
    cfg = orc.default_config(n_layer=2)
    model, sd = _model(cfg, 23, cuda)
    L = 1024
    if workload == "atari_C3":
        batch = [synth.rl_atari_batch(cfg, 1, L, seed=31)]
    else:
        batch = [synth.rl_continuous_batch(cfg, 1, L, seed=41), synth.nlp_batch(cfg, 1, L, seed=42),
                 synth.ic_batch(cfg, 1, L, seed=43)]
    tasks = [_task_dict(t) for t in batch]  # before the model overwrites image-slot labels (-1 -> 0, :645)
    logits, loss = model([synth.to_device(t, cuda) for t in batch])
    (loss * LOSS_SCALE).backward()
    torch.cuda.synchronize()
    sdo = _oracle_sd(sd)
    ologits, oloss = orc.forward(tasks, sdo, cfg)
    oloss.backward()
    err = util.rel_err(logits, ologits)
    print("%s: logits max-norm rel %.3e, loss %.6f vs %.6f" % (workload, err, loss.item(), oloss.item()))
    assert err <= 3e-3
    assert abs(loss.item() - oloss.item()) <= 1e-3 * abs(oloss.item())
    worst = {}
    for k, p in model.named_parameters():
        og = sdo[k].grad
        if og is None or og.abs().max() == 0:
            continue
        assert p.grad is not None, k
        worst[k] = util.rel_l2(p.grad.float() / LOSS_SCALE, og)
    print("%s worst gradient rel-L2:" % workload, sorted(worst.items(), key=lambda kv: -kv[1])[:6])
    assert max(worst.values()) <= 2e-2, {k: v for k, v in worst.items() if v > 2e-2}
    if workload == "atari_C3":
        assert any("patch_embeddings.projection" in k or "patch_embedding.projection" in k for k in worst), sorted(worst)[:40]
