"""How far does the REFERENCE's own fp16 mode (module.half(), what DeepSpeed fp16 runs) drift from its fp32 mode at
DB1-1.3B depth/width with random-init weights? Build container only (imports /root/reference). CPU, takes minutes.
    python tools/ref_fp16_drift.py [n_layer] [L]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import db1_oracle as orc  # noqa: E402

sys.path.insert(0, "/root/reference")
import types  # noqa: E402
for m in ("gym", "d4rl", "tree"):
    sys.modules.setdefault(m, types.ModuleType(m))
from src.model import TransformerXL  # noqa: E402  (the reference's)
from src.data.input_specs import NLPTaskInput  # noqa: E402

_pos = [a for a in sys.argv[1:] if not a.startswith("--")]
nl = int(_pos[0]) if len(_pos) > 0 else 24
L = int(_pos[1]) if len(_pos) > 1 else 128
cfg = orc.default_config(n_layer=nl)
ns = types.SimpleNamespace(**vars(cfg))
torch.manual_seed(0)
model = TransformerXL(ns).eval()
sd = orc.synth_state_dict(cfg, seed=21)
sd = {k: (v.half().float() if v.is_floating_point() and k != "pos_emb.inv_freq" else v) for k, v in sd.items()}
model.load_state_dict(sd, strict=False)
tok = torch.randint(0, 32000, (1, L), generator=torch.Generator().manual_seed(3))
mk = lambda: NLPTaskInput(position_id=None, attention_mask=None, loss_mask=torch.ones(1, L), label=tok.clone(),  # noqa: E731
                          text_seq=tok.clone(), text_len=None)
with torch.no_grad():
    t0 = time.time()
    l32, _ = model([mk()])
    print("fp32 forward %.1f s" % (time.time() - t0), flush=True)
    model.half()
    t0 = time.time()
    l16, _ = model([mk()])
    print("fp16 forward %.1f s" % (time.time() - t0), flush=True)
d = (l16.float() - l32)
mx, rms = (d.abs().max() / l32.abs().max()).item(), (d.pow(2).mean().sqrt() / l32.pow(2).mean().sqrt()).item()
print("reference fp16 vs reference fp32, %d layers, L=%d: max-norm rel %.3e, rms rel %.3e" % (nl, L, mx, rms))
if "--write" in sys.argv:
    import json
    out = os.path.join(ROOT, "tests", "golden", "ref_fp16_drift.json")
    with open(out, "w") as f:
        json.dump({"what": "logits of the UNMODIFIED reference after module.half() vs the same module in fp32 (CPU), random-init "
                           "weights from oracle.synth_state_dict(seed=21) rounded to fp16, one NLP sequence",
                   "n_layer": nl, "seq_len": L, "max_norm_rel": mx, "rms_rel": rms,
                   "generated_by": "python tools/ref_fp16_drift.py %d %d --write" % (nl, L)}, f, indent=1)
    print("written", out)
