import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "bdm-db1_b200"))
from oracle import db1_oracle as orc
from tests import util
from tests.test_model_gpu import _build
cuda = torch.device("cuda")
cfg = orc.tiny_config(text_vocab_size=480, mem_len=64, n_position=64)
model, sd = _build(cfg, 13, cuda); model.eval()
sdo = {k: v.clone() for k, v in sd.items()}
for k in list(sdo):
    if k.startswith("h.") and k.endswith(("r_r_bias", "r_w_bias")): sdo[k] = sdo[k.split(".")[-1]]
B = 2
g = torch.Generator().manual_seed(7)
mems = model.init_mem(B); omems = [m.float().cpu() for m in mems]
for step, q in enumerate([9, 1, 1]):
    tok = torch.randint(0, 480, (B, q), generator=g)
    task = dict(type="nlp", text_seq=tok.numpy(), label=np.zeros((B, q), np.int64), loss_mask=np.ones((B, q), np.float32))
    with torch.no_grad():
        lg_v, _, m_v = model(util.to_model_inputs([task], cuda), compute_loss=False, mems=mems)
        lg_c, _, m_c = model(util.to_model_inputs([task], cuda), compute_loss=False, mems=[m.clone().contiguous() for m in mems])
        ol, _, omems = orc.forward([task], sdo, cfg, compute_loss=False, mems=omems)
    print(step, q, "views vs oracle %.2e | contiguous copies vs oracle %.2e | mems contiguous: %s" % (util.rel_err(lg_v, ol), util.rel_err(lg_c, ol), mems[0].is_contiguous()))
    print("   new mems vs oracle:", ["%.1e" % util.rel_err(a, b) for a, b in zip(m_v, omems)])
    mems = m_v
