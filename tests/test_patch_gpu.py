"""GPU parity of the image patch embedder (db1_sm100.vision: standardise + conv3x3, GroupNorm + GELU + im2col, tcgen05
GEMMs for the 64->64 convs and the 16x16 projection) against the oracle's restatement of
src/tokenizer/vision_embedding.py:65-86 and its fp32 autograd gradients. Tolerance: fp16 storage of every intermediate
(as the reference's half-precision module) vs an fp32 oracle: output 3e-3 of max, gradients rel-L2 2e-2."""
from types import SimpleNamespace

import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu


def _params(d, C, seed):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s, sc=1.0: torch.randn(*s, generator=g) * sc  # noqa: E731
    pre = "pe."
    return {
        pre + "conv1.weight": r(64, C, 3, 3, sc=0.3), pre + "conv1.bias": r(64, sc=0.1),
        pre + "residual_path.0.weight": 1 + r(64, sc=0.1), pre + "residual_path.0.bias": r(64, sc=0.1),
        pre + "residual_path.2.weight": r(64, 64, 3, 3, sc=0.06), pre + "residual_path.2.bias": r(64, sc=0.1),
        pre + "residual_path.3.weight": 1 + r(64, sc=0.1), pre + "residual_path.3.bias": r(64, sc=0.1),
        pre + "residual_path.5.weight": r(64, 64, 3, 3, sc=0.06), pre + "residual_path.5.bias": r(64, sc=0.1),
        pre + "projection.weight": r(d, 64, 16, 16, sc=0.02), pre + "projection.bias": r(d, sc=0.1),
    }


@pytest.mark.parametrize("N,H,W,d", [(2, 32, 48, 128), (1, 16, 16, 256), (5, 80, 80, 128)])
def test_patch_embed_fwd_bwd_matches_oracle(cuda, N, H, W, d):
    from db1_sm100 import vision
    from oracle import db1_oracle as orc
    C = 3
    sd = {k: v.half().float() for k, v in _params(d, C, 3).items()}  # both sides start from the same fp16 values
    g = torch.Generator().manual_seed(4)
    pixels = torch.rand(N, C, H, W, generator=g).half()
    P = N * (H // 16) * (W // 16)
    pos = (torch.randn(P, d, generator=g) * 0.1).half()
    dout = torch.randn(N, P // N, d, generator=g).half()
    cfg = SimpleNamespace(vision_patch_size=16)
    ref_sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    pos_ref = pos.float().requires_grad_(True)
    ref = orc.patch_embeddings(pixels.float(), ref_sd, "pe.", cfg) + pos_ref.view(N, P // N, d)
    (ref * dout.float()).sum().backward()

    dev = {k: v.half().to(cuda).requires_grad_(True) for k, v in sd.items()}
    pos_d = pos.to(cuda).requires_grad_(True)
    names = ["conv1.weight", "conv1.bias", "residual_path.0.weight", "residual_path.0.bias", "residual_path.2.weight",
             "residual_path.2.bias", "residual_path.3.weight", "residual_path.3.bias", "residual_path.5.weight",
             "residual_path.5.bias", "projection.weight", "projection.bias"]
    out = vision.PatchEmbedFn.apply(pixels.to(cuda), pos_d, *[dev["pe." + n] for n in names], 1e-5, 1e-5)
    assert out.shape == (N, P // N, d)
    assert util.rel_err(out, ref) <= 3e-3
    out.backward(dout.to(cuda))
    torch.cuda.synchronize()
    assert util.rel_l2(pos_d.grad, pos_ref.grad) <= 1e-3
    worst = {n: util.rel_l2(dev["pe." + n].grad, ref_sd["pe." + n].grad) for n in names}
    print(worst)
    assert max(worst.values()) <= 2e-2, worst


def test_patch_embed_fp32_low_contrast_pixels(cuda):
    """fp32 frames are standardised from their unrounded values, as the reference does (vision_embedding.py:73-79:
    standardise in the input dtype, cast afterwards). On low-contrast patches (Atari backgrounds) a prior fp16 cast of
    the pixels is amplified by 1/(1e-6 + std) to percent-level errors; the fp32 path stays at kernel rounding."""
    from db1_sm100 import vision
    from oracle import db1_oracle as orc
    N, C, H, W, d = 2, 3, 32, 32, 128
    sd = {k: v.half().float() for k, v in _params(d, C, 3).items()}
    g = torch.Generator().manual_seed(9)
    pixels = 0.62 + 2e-3 * torch.rand(N, C, H, W, generator=g)  # contrast of a few fp16 ulps around 0.62
    cfg = SimpleNamespace(vision_patch_size=16)
    with torch.no_grad():
        ref = orc.patch_embeddings(pixels, sd, "pe.", cfg)
        ref16 = orc.patch_embeddings(pixels.half().float(), sd, "pe.", cfg)
    dev = {k: v.half().to(cuda) for k, v in sd.items()}
    names = ["conv1.weight", "conv1.bias", "residual_path.0.weight", "residual_path.0.bias", "residual_path.2.weight",
             "residual_path.2.bias", "residual_path.3.weight", "residual_path.3.bias", "residual_path.5.weight",
             "residual_path.5.bias", "projection.weight", "projection.bias"]
    with torch.no_grad():
        out = vision.PatchEmbedFn.apply(pixels.to(cuda), None, *[dev["pe." + n] for n in names], 1e-5, 1e-5)
    err32 = util.rel_err(out, ref)
    assert err32 <= 3e-3, err32
    assert util.rel_err(ref16, ref) > 10 * err32  # what casting the pixels first would have cost


def test_transpose_and_dropout_kernels(cuda):
    from db1_sm100 import ops
    x = torch.randn(7, 64, 9).half().to(cuda)
    y = ops.transpose(x, torch.empty(7, 9, 64, dtype=torch.half, device=cuda), 7, 64, 9)
    assert torch.equal(y, x.transpose(1, 2).contiguous())
    a = torch.randn(64, 256).half().to(cuda)
    o1 = ops.dropout(a, torch.empty_like(a), 0.25, 99)
    o2 = ops.dropout(a, torch.empty_like(a), 0.25, 99)
    assert torch.equal(o1, o2)
    keep = o1 != 0
    assert abs(keep.float().mean().item() - 0.75) < 0.02
    assert util.rel_err(o1[keep], (a.float() / 0.75)[keep]) < 1e-3
