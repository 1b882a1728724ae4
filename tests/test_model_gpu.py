"""GPU parity of the full DB1 module (src.model.TransformerXL on the sm_100a kernels) against the CPU oracle and the
reference-generated golden fixtures. Tolerances: fp16 storage / fp32 accumulation vs an fp32 oracle —
logits max|d|/max|ref| <= 2e-3 (tiny, 2 layers), loss rel 1e-3, gradients rel-L2 <= 2e-2."""
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu


def _build(cfg, seed, cuda):
    from oracle import db1_oracle as orc
    from src.model import TransformerXL
    sd = orc.synth_state_dict(cfg, seed=seed)
    model = TransformerXL(cfg)
    res = model.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert set(model.state_dict().keys()) == set(orc.state_shapes(cfg).keys())
    model = model.half().to(cuda)
    # these tests compare with the fp32 oracle and the fp32 reference goldens: exact fp32 sinusoid phases (the default after
    # .half() is the reference's fp16 phase arithmetic, covered by test_half_phase_default_matches_oracle_half_phase)
    model.pos_emb.phase_dtype = torch.float32
    return model, sd


@pytest.mark.parametrize("name", ["tiny_text_rl", "tiny_window_clamp", "tiny_mixed_images"])
def test_tiny_forward_backward_matches_oracle_and_golden(cuda, name):
    from oracle import db1_oracle as orc
    g = util.load_golden(name)
    cfg = util.golden_cfg(g)
    model, sd = _build(cfg, int(g["seed"][0]), cuda)
    model.eval()
    tasks = util.tasks_from_golden(g)
    logits, loss = model(util.to_model_inputs(tasks, cuda))
    # fp16 gradients need loss scaling, exactly as the reference's DeepSpeed fp16 run (initial scale 2^12,
    # scripts/evaluate/evaluate_rl_1.2B.sh:36-39); gradients are compared after unscaling
    LOSS_SCALE = 4096.0
    (loss * LOSS_SCALE).backward()
    torch.cuda.synchronize()
    # reference golden (fp32 reference run in the build container)
    assert abs(loss.item() - g["loss"][0]) < 1e-3 * abs(g["loss"][0])
    sub = logits[:, ::7, ::13].float().cpu()
    ref_sub = torch.as_tensor(g["logits_sub"])
    assert (sub - ref_sub).abs().max().item() <= 2e-3 * float(g["logits_absmax"][0])
    # full tensors against the oracle
    sdo = {k: v.clone().requires_grad_(v.is_floating_point() and k != "pos_emb.inv_freq") for k, v in sd.items()}
    for k in list(sdo):
        if k.startswith("ic_encoder."):
            sdo[k] = sdo["vision_encoder." + k[len("ic_encoder."):]]
        if k.startswith("h.") and k.endswith(("r_r_bias", "r_w_bias")):
            sdo[k] = sdo[k.split(".")[-1]]
    ologits, oloss = orc.forward(tasks, sdo, cfg)
    oloss.backward()
    assert util.rel_err(logits, ologits) <= 2e-3
    worst = {}
    for k, p in model.named_parameters():
        og = sdo[k].grad
        if og is None or og.abs().max() == 0:
            continue
        assert p.grad is not None, k
        worst[k] = util.rel_l2(p.grad.float() / LOSS_SCALE, og)
    bad = {k: v for k, v in worst.items() if v > 2e-2}
    print("worst gradient rel-L2:", sorted(worst.items(), key=lambda kv: -kv[1])[:6])
    assert not bad, bad
    assert len(worst) >= 20


@pytest.mark.parametrize("ga", [1, 2])
def test_engine_direct_gradient_writes_match_autograd(cuda, ga):
    """DB1Engine lets the wgrad kernels write into its flat buckets (gradient sink). The result must equal the gradients
    autograd accumulates on its own, step after step (no stale data from the previous window, shared parameters
    accumulated once per use, accumulation over `ga` micro-steps)."""
    from db1_sm100 import functions as F_
    from db1_sm100.engine import DB1Engine
    g = util.load_golden("tiny_mixed_images")
    cfg = util.golden_cfg(g)
    tasks = util.tasks_from_golden(g)
    model, _sd = _build(cfg, 5, cuda)
    model.eval()
    SCALE = 1024.0

    def plain_grads(task_sets):
        for p in model.parameters():
            p.grad = None
        F_.set_grad_sink(None)
        for ts in task_sets:
            _, loss = model(util.to_model_inputs(ts, cuda))
            (loss * (SCALE / len(task_sets))).backward()
        return {n: (p.grad.clone() if p.grad is not None else None) for n, p in model.named_parameters()}

    sets_a = [tasks] * ga
    sets_b = [tasks[:2]] * ga  # text + RL only: the vision encoder gets no gradient in this window
    ref_a, ref_b = plain_grads(sets_a), plain_grads(sets_b)
    for p in model.parameters():
        p.grad = None
    try:
        eng = DB1Engine(model, gradient_accumulation_steps=ga, loss_scale=SCALE)
        for ref, sets in ((ref_a, sets_a), (ref_b, sets_b), (ref_a, sets_a)):
            for ts in sets:
                _, loss = eng(util.to_model_inputs(ts, cuda))
                eng.backward(loss)
            torch.cuda.synchronize()
            for n, p in model.named_parameters():
                r = ref[n]
                if r is None:
                    assert p.grad.abs().max().item() == 0, n
                else:
                    assert util.rel_l2(p.grad, r) <= 2e-3 or (p.grad.float() - r.float()).abs().max().item() < 1e-3, n
        assert len(eng._sink_seen) >= 20
    finally:
        F_.set_grad_sink(None)


def test_full_depth_1p3b_logits_match_oracle(cuda):
    """DB1-1.3B (24 layers, d 2048, 16 heads of 128, GeGLU 8192, vocab 33 025), one 256-token RL sequence, fp32 CPU
    oracle on the same (fp16-representable) weights.

    Two statements, both measured on B200 (tools/debug_depth.py prints the per-layer table):
      * per layer / per op with the SAME input: error = fp16 rounding of the stored output (rms 5.5e-4, max ~1e-3 of the
        layer output; tied head 3e-4 rms) - asserted here for the first, a middle and the last layer and the head;
      * chained through all 24 layers the fp16 storage rounding is amplified by the random-init post-LN network itself
        (~1.13x per layer): 2.0e-2 max-norm / 1.7e-2 rms on the logits. The REFERENCE's own fp16 mode (module.half(), what
        its DeepSpeed fp16 run executes) drifts MORE from its fp32 mode on the same weights: 3.8e-2 / 2.6e-2
        (tools/ref_fp16_drift.py, build container). BASELINE.json's 1e-3 is therefore a per-layer figure, not reachable
        end to end by any fp16-storage implementation including the reference; the bound asserted for the chain is the
        reference's own drift."""
    from db1_sm100 import functions as F_
    from db1_sm100 import synth
    from oracle import db1_oracle as orc
    from src.model import TransformerXL
    cfg = orc.default_config()
    sd = orc.synth_state_dict(cfg, seed=21)
    sd = {k: (v.half().float() if v.is_floating_point() and k != "pos_emb.inv_freq" else v) for k, v in sd.items()}
    model = TransformerXL(cfg)
    model.load_state_dict(sd, strict=True)
    model = model.half().to(cuda).eval()
    model.pos_emb.phase_dtype = torch.float32  # fp32 oracle on the other side
    L = 256
    rl = synth.rl_continuous_batch(cfg, 1, L, seed=77)
    task = dict(type="rl", tensor_seq=rl.tensor_seq.numpy(), label=rl.label.numpy(), loss_mask=rl.loss_mask.numpy(),
                position_id=rl.position_id.numpy(), vision_seq=None)
    sdo = dict(sd)
    for k in list(sdo):
        if k.startswith("h.") and k.endswith(("r_r_bias", "r_w_bias")):
            sdo[k] = sdo[k.split(".")[-1]]

    def rms_rel(a, b):
        a = a.float().cpu()
        return ((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt()).item()

    with torch.no_grad():
        logits, loss = model([synth.to_device(rl, cuda)])
        x, _m, _l = orc.embed_task(task, sdo, cfg)
        ok = orc.attention_mask_ok(L, L, cfg.mem_len, cfg.same_length)
        pe = orc.positional_rows(L, cfg.n_embed, cfg.n_position)
        pos_rows = model.pos_emb.rows(L, model.clamp_len, 0.0)
        for li, block in enumerate(model.h):
            x_prev = x
            x = orc.decoder_layer(x, pe, sdo, "h.%d." % li, cfg, ok, None)
            if li in (0, 11, 23):  # same input -> only this layer's own rounding
                h_same = block(x_prev.half().to(cuda), pos_rows, window=1 << 30)[0]
                assert util.rel_err(h_same, x) <= 2e-3 and rms_rel(h_same, x) <= 8e-4, li
        ologits = torch.nn.functional.linear(x, sdo["word_embedding.weight"])
        lg_same = F_.head_logits(x.half().to(cuda), model.word_embedding.weight)
        assert util.rel_err(lg_same, ologits) <= 1e-3 and rms_rel(lg_same, ologits) <= 5e-4
        _ol, oloss = orc.forward([task], sdo, cfg)
    err, rms = util.rel_err(logits, ologits), rms_rel(logits, ologits)
    print("1.3B logits, chained: max-norm rel err %.3e, rms rel err %.3e, loss %.6f vs %.6f" % (err, rms, loss.item(), oloss.item()))
    import json
    import os
    with open(os.path.join(util.GOLD, "ref_fp16_drift.json")) as f:
        drift = json.load(f)  # the reference's own fp16-vs-fp32 drift on these weights (tools/ref_fp16_drift.py --write)
    assert err <= drift["max_norm_rel"] and rms <= drift["rms_rel"]
    assert abs(loss.item() - oloss.item()) <= 2e-3 * abs(oloss.item())


def test_fused_adam_step_matches_torch_adamw_and_skips_on_overflow(cuda):
    """DB1Engine(fused_adam=...): db1_grad_sumsq + db1_clip_coef + db1_adam_step on the flat buckets vs torch.optim.AdamW on
    fp32 copies fed the same unscaled, clipped gradients; overflow (inf in a gradient) skips the step and halves the scale."""
    from db1_sm100 import functions as F_
    from db1_sm100.engine import DB1Engine
    g = util.load_golden("tiny_text_rl")
    cfg = util.golden_cfg(g)
    tasks = util.tasks_from_golden(g)
    model, _sd = _build(cfg, 6, cuda)
    model.train()
    hp = dict(lr=3e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1)
    CLIP, SCALE = 0.5, 1024.0
    try:
        eng = DB1Engine(model, loss_scale=SCALE, clip_grad=CLIP, fused_adam=hp, loss_scale_window=0)
        ref_params = [p.detach().float().clone().requires_grad_(True) for p in model.parameters()]
        ref_opt = torch.optim.AdamW(ref_params, lr=hp["lr"], betas=hp["betas"], eps=hp["eps"], weight_decay=hp["weight_decay"])
        for step in range(3):
            _, loss = eng(util.to_model_inputs(tasks, cuda))
            eng.backward(loss)
            grads = [p.grad.float() / SCALE for p in model.parameters()]
            norm = torch.sqrt(sum((x * x).sum() for x in grads))
            coef = min(1.0, CLIP / (norm.item() + 1e-6))
            assert eng.step() is True
            assert abs(eng.grad_norm().item() - norm.item()) <= 1e-3 * norm.item()
            for rp, gr in zip(ref_params, grads):
                rp.grad = gr * coef
            ref_opt.step()
            flat_master = {id(p): None for p in model.parameters()}
            for b in eng.buckets:
                off = 0
                for p in b.params:
                    flat_master[id(p)] = b.master[off:off + p.numel()].view_as(p)
                    off += (p.numel() + 7) // 8 * 8
            worst = max(util.rel_l2(flat_master[id(p)], rp) for p, rp in zip(model.parameters(), ref_params))
            assert worst <= 2e-5, (step, worst)
            for p in model.parameters():  # fp16 parameters = rounded masters
                assert torch.equal(p.data, flat_master[id(p)].half())
        # overflow: poison one gradient
        _, loss = eng(util.to_model_inputs(tasks, cuda))
        eng.backward(loss)
        eng.buckets[1].flat[3] = float("inf")
        before = eng.buckets[1].master.clone()
        assert eng.step() is False and eng.loss_scale == SCALE / 2
        assert torch.equal(eng.buckets[1].master, before)
    finally:
        F_.set_grad_sink(None)


def test_reference_error_behaviour_is_kept(cuda):
    """AssertionError for compute_loss with mems (:515-517), ValueError when same_length with mem_len == 0 masks every key
    (:205-206), in-place label[label == -1] = 0 on RL batches with images (:645), Db1Error (no fallback) off the GPU."""
    from db1_sm100._lib import Db1Error
    from oracle import db1_oracle as orc
    from src.model import TransformerXL
    g = util.load_golden("tiny_mixed_images")
    cfg = util.golden_cfg(g)
    tasks = util.tasks_from_golden(g)
    model, sd = _build(cfg, 5, cuda)
    model.eval()
    inputs = util.to_model_inputs(tasks, cuda)
    with pytest.raises(AssertionError):
        model(inputs, compute_loss=True, mems=model.init_mem(1))
    rl_img = [t for t in inputs if getattr(t, "vision_seq", None) is not None][0]
    assert (rl_img.label == -1).any()
    with torch.no_grad():
        model(inputs)
    assert not (rl_img.label == -1).any()
    cfg0 = orc.tiny_config(text_vocab_size=480, mem_len=0)
    m0 = TransformerXL(cfg0).half().to(cuda).eval()
    with pytest.raises(ValueError):
        m0(util.to_model_inputs(tasks[1:2], cuda))
    cpu_model = TransformerXL(cfg)
    with pytest.raises(Db1Error):
        cpu_model(util.to_model_inputs(tasks[1:2], torch.device("cpu"), half=False))


def test_half_phase_default_matches_oracle_half_phase(cuda):
    """After module.half() the sinusoid phases follow the reference's fp16 arithmetic (transformer_xl.py:44, :569-571):
    db1_posemb_half_phase rows are bit-identical to the oracle's half-phase rows (pinned to the reference after .half()
    by tests/golden/posemb_half.npz), differ visibly from the fp32-phase rows, and the model's logits match the oracle
    run with the same rows to the usual tolerance."""
    import copy
    from oracle import db1_oracle as orc
    from db1_sm100 import synth
    cfg = orc.tiny_config(text_vocab_size=480)
    model, sd = _build(cfg, 6, cuda)
    model.eval()
    model.pos_emb.phase_dtype = None  # the default
    assert model.pos_emb.inv_freq.dtype == torch.float16
    for klen, clamp in ((256, 256), (300, 100)):
        rows = model.pos_emb.rows(klen, clamp, 0.0).float().cpu()
        assert torch.equal(rows, orc.positional_rows(klen, cfg.n_embed, clamp, half_phase=True))
        assert (rows - orc.positional_rows(klen, cfg.n_embed, clamp)).abs().max().item() > 1e-2
    big = orc.positional_rows(1024, 2048, 1024, half_phase=True)
    m2 = type(model.pos_emb)(2048).half().to(cuda)
    assert torch.equal(m2.rows(1024, 1024, 0.0).float().cpu(), big)
    nlp = synth.nlp_batch(cfg, 2, 256, seed=12)
    task = dict(type="nlp", text_seq=nlp.text_seq.numpy(), label=nlp.label.numpy(), loss_mask=nlp.loss_mask.numpy())
    with torch.no_grad():
        logits, loss = model([synth.to_device(nlp, cuda)])
    sdo = {k: v.clone() for k, v in sd.items()}
    for k in list(sdo):
        if k.startswith("h.") and k.endswith(("r_r_bias", "r_w_bias")):
            sdo[k] = sdo[k.split(".")[-1]]
    cfg_h = copy.copy(cfg)
    cfg_h.pos_phase_half = True
    with torch.no_grad():
        ol_h, oloss_h = orc.forward([task], sdo, cfg_h)
    assert util.rel_err(logits, ol_h) <= 2e-3
    assert abs(loss.item() - oloss_h.item()) <= 1e-3 * abs(oloss_h.item())


@pytest.mark.gpu
def test_deepspeed_layout_checkpoint_into_db1_module(cuda, tmp_path):
    """f4: a checkpoint in DeepSpeed 0.6.7's layout (tests/ds_ckpt.py) loaded through DB1Engine.load_checkpoint into the
    B200 module with the fused optimizer: logits equal those of a module built directly from the same weights; the fp32
    masters are rebuilt from the loaded weights (foreign FP16_Optimizer state is not restored)."""
    import warnings
    from oracle import db1_oracle as orc
    from db1_sm100 import synth
    from db1_sm100.engine import DB1Engine
    from src.model import TransformerXL
    from tests.ds_ckpt import write_deepspeed_style_checkpoint
    cfg = orc.tiny_config(text_vocab_size=480)
    ref_model, sd = _build(cfg, 31, cuda)
    ref_model.eval()
    ck = {k: (v.half() if v.is_floating_point() else v) for k, v in ref_model.state_dict().items()}  # fp16 weights, as saved
    write_deepspeed_style_checkpoint(str(tmp_path), {k: v.cpu() for k, v in ck.items()})
    torch.manual_seed(123)
    fresh = TransformerXL(cfg).half().to(cuda).eval()  # different random init
    fresh.pos_emb.phase_dtype = torch.float32
    eng = DB1Engine(fresh, loss_scale=4096.0, fused_adam=dict(lr=1e-4))
    nlp = synth.nlp_batch(cfg, 2, 256, seed=3)
    with torch.no_grad():
        before, _ = eng([synth.to_device(nlp, cuda)])
        want, _ = ref_model([synth.to_device(nlp, cuda)])
    assert not torch.equal(before, want)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        path, client = eng.load_checkpoint(str(tmp_path))
    assert w and client["iteration"] == 1234 and eng.global_steps == 1234 and eng.loss_scale == 32768.0
    with torch.no_grad():
        after, _ = eng([synth.to_device(nlp, cuda)])
    assert torch.equal(after, want)
    for b in eng.buckets:
        assert torch.equal(b.master, b.pflat.float()) and b.m.abs().max().item() == 0


def test_vqa_input_takes_the_image_caption_path(cuda):
    """VQATaskInput (reference _forward_vqa, transformer_xl.py:705-748) embeds prompt | image patches | text exactly like
    ICTaskInput (:675-703): the same tensors wrapped in either dataclass give bit-identical logits and loss and the same gradients."""
    from src.data.input_specs import ICTaskInput, VQATaskInput
    g = util.load_golden("tiny_mixed_images")
    cfg = util.golden_cfg(g)
    tasks = util.tasks_from_golden(g)
    inputs = util.to_model_inputs(tasks, cuda)
    ics = [t for t in inputs if isinstance(t, ICTaskInput)]
    assert ics, "the mixed golden holds an image-caption segment"
    ic = ics[0]
    vqa = VQATaskInput(position_id=ic.position_id, attention_mask=ic.attention_mask, loss_mask=ic.loss_mask, label=ic.label,
                       prompt_seq=ic.prompt_seq, img_seq=ic.img_seq, text_seq=ic.text_seq, img_id_seq=ic.img_id_seq,
                       ques_id_seq=None, ques_len=torch.full((ic.text_seq.shape[0],), 3, device=cuda))
    outs = []
    for task in (ic, vqa):
        model, _sd = _build(cfg, int(g["seed"][0]), cuda)
        model.eval()
        logits, loss = model([task])
        (loss * 4096.0).backward()
        torch.cuda.synchronize()
        outs.append((logits.clone(), loss.clone(), {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert outs[0][2].keys() == outs[1][2].keys() and len(outs[0][2]) > 20
    for k in outs[0][2]:
        # forward is deterministic; several gradients are accumulated with atomics (embedding scatter, fp32 column sums),
        # whose order varies between two runs of the SAME input - hence a tolerance instead of equality
        assert util.rel_l2(outs[0][2][k].float(), outs[1][2][k].float()) < 2e-3, k
