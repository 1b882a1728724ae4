"""Per-layer error growth of the sm_100a path vs the fp32 oracle at DB1-1.3B width (development tool)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bdm-db1_b200"))
sys.path.insert(0, ROOT)
from oracle import db1_oracle as orc  # noqa: E402
from db1_sm100 import synth  # noqa: E402
from src.model import TransformerXL  # noqa: E402

nl = int(sys.argv[1]) if len(sys.argv) > 1 else 24
cuda = torch.device("cuda")
cfg = orc.default_config(n_layer=nl)
sd = orc.synth_state_dict(cfg, seed=21)
sd = {k: (v.half().float() if v.is_floating_point() and k != "pos_emb.inv_freq" else v) for k, v in sd.items()}
model = TransformerXL(cfg)
model.load_state_dict(sd, strict=True)
model = model.half().to(cuda).eval()
L = 256
rl = synth.rl_continuous_batch(cfg, 1, L, seed=77)
task = dict(type="rl", tensor_seq=rl.tensor_seq.numpy(), label=rl.label.numpy(), loss_mask=rl.loss_mask.numpy(),
            position_id=rl.position_id.numpy(), vision_seq=None)
sdo = dict(sd)
for k in list(sdo):
    if k.startswith("h.") and k.endswith(("r_r_bias", "r_w_bias")):
        sdo[k] = sdo[k.split(".")[-1]]


def rel(a, b):
    a = a.float().cpu()
    return ((a - b).abs().max() / b.abs().max()).item(), ((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt()).item()


with torch.no_grad():
    x, _m, _l = orc.embed_task(task, sdo, cfg)
    ok = orc.attention_mask_ok(L, L, cfg.mem_len, cfg.same_length)
    pe = orc.positional_rows(L, cfg.n_embed, cfg.n_position)
    dev_in = synth.to_device(rl, cuda)
    h, _, _, _ = model._forward_rl(dev_in, 0.0)
    print("embed", rel(h, x))
    pos_rows = model.pos_emb.rows(L, model.clamp_len, 0.0)
    print("posemb", rel(pos_rows, pe.reshape(L, -1)))
    hx = h
    for li, block in enumerate(model.h):
        x_prev = x
        x = orc.decoder_layer(x, pe, sdo, "h.%d." % li, cfg, ok, None)
        hx = block(hx, pos_rows, window=1 << 30)[0]
        # same-input error: feed the ORACLE's previous activation (rounded to fp16) through our layer
        h_same = block(x_prev.half().to(cuda), pos_rows, window=1 << 30)[0]
        print("layer %2d  chained max/rms %.2e %.2e | same-input max/rms %.2e %.2e | |x| max %.1f rms %.2f" %
              ((li,) + rel(hx, x) + rel(h_same, x) + (x.abs().max().item(), x.pow(2).mean().sqrt().item())))
    logits = torch.nn.functional.linear(x, sdo["word_embedding.weight"])
    from db1_sm100 import functions as F_
    lg = F_.head_logits(hx, model.word_embedding.weight)
    print("logits chained", rel(lg, logits), "same-input", rel(F_.head_logits(x.half().to(cuda), model.word_embedding.weight), logits))
