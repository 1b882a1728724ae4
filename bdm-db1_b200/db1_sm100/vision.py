"""PatchEmbeddings.forward/backward (src/tokenizer/vision_embedding.py:65-86) as a fixed sequence of C-ABI launches.

    pixels [N,C,H,W] --db1_patch_conv1_fwd--> y1 [P,256,64]            (patch split, standardise, conv3x3 C->64)
    y1 --db1_gn_gelu_im2col--> col2 [P*256,576] --db1_gemm_f16(W2)--> c2       (GroupNorm, GELU, conv3x3 64->64)
    c2 --db1_gn_gelu_im2col--> col3 --db1_gemm_f16(W3, resid = y1)--> out      (second pair + the residual add)
    out.view(P,16384) --db1_gemm_f16(Wp, bias, resid = row+col position rows)--> [P, d]   (16x16/stride-16 projection)

Activations are channels-last per patch, so conv weights are used as [co][tap*64+ci] and the projection weight as
[d][pixel*64+c]; db1_transpose_f16 produces those views of the reference-shaped parameters (and maps the gradients back).
"""
import torch

from . import ops
from .functions import _f32zeros, _to_half


def _splits(P, cap=128):
    """Largest divisor of P that is <= cap: number of K-slices of the weight-gradient GEMMs (K = P*256 pixels)."""
    for s in range(min(P, cap), 0, -1):
        if P % s == 0:
            return s
    return 1


def _conv_wgrad(dY, col, P):
    """dW' [64, 576] = dY^T [64, M] . col [M, 576], M = P*256, as S parallel K-slices + one column-sum (fp32)."""
    dev = dY.device
    M = P * 256
    S = _splits(P)
    Ks = M // S
    part = torch.empty(S, 64, 576, dtype=torch.float16, device=dev)
    ops.gemm(dY, col, part, 64, 576, Ks, lda=64, ldb=576, ldc=576, a_mn=True, b_mn=True, Z1=S,
             a_z=(Ks * 64, 0), b_z=(Ks * 576, 0), c_z=(64 * 576, 0))
    acc = _f32zeros(64 * 576, dev)
    ops.colsum(part.view(S, 64 * 576), acc, S, 64 * 576)
    return acc


class PatchEmbedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pixels, pos_sum, W1, b1, g1, be1, W2, b2, g2, be2, W3, b3, Wp, bp, eps1, eps2):
        N, Cc, Hh, Ww = pixels.shape
        dev = pixels.device
        f16 = torch.float16
        d = Wp.shape[0]
        P = N * (Hh // 16) * (Ww // 16)
        M = P * 256
        pixels = pixels.contiguous()
        xs = torch.empty(P, Cc, 256, dtype=f16, device=dev)
        y1 = torch.empty(P, 256, 64, dtype=f16, device=dev)
        ops.patch_conv1_fwd(pixels, W1, b1, xs, y1)
        W2t = ops.transpose(W2, torch.empty(64, 9, 64, dtype=f16, device=dev), 64, 64, 9)
        W3t = ops.transpose(W3, torch.empty(64, 9, 64, dtype=f16, device=dev), 64, 64, 9)
        Wpt = ops.transpose(Wp, torch.empty(d, 256, 64, dtype=f16, device=dev), d, 64, 256)
        col2 = torch.empty(M, 576, dtype=f16, device=dev)
        st1 = torch.empty(P, 32, 2, dtype=torch.float32, device=dev)
        ops.gn_gelu_im2col(y1, g1, be1, col2, st1, P, eps1)
        c2 = torch.empty(P, 256, 64, dtype=f16, device=dev)
        ops.gemm(col2, W2t, c2, M, 64, 576, lda=576, ldb=576, ldc=64, bias=b2)
        col3 = torch.empty(M, 576, dtype=f16, device=dev)
        st2 = torch.empty(P, 32, 2, dtype=torch.float32, device=dev)
        ops.gn_gelu_im2col(c2, g2, be2, col3, st2, P, eps2)
        out = torch.empty(P, 256, 64, dtype=f16, device=dev)
        ops.gemm(col3, W3t, out, M, 64, 576, lda=576, ldb=576, ldc=64, bias=b3, resid=y1, ldr=64)
        emb = torch.empty(P, d, dtype=f16, device=dev)
        ops.gemm(out, Wpt, emb, P, d, 16384, lda=16384, ldb=16384, ldc=d, bias=bp, resid=pos_sum,
                 ldr=d if pos_sum is not None else 0)
        ctx.save_for_backward(xs, y1, c2, out, col2, col3, st1, st2, g1, be1, g2, be2, W2t, W3t, Wpt)
        ctx.cfg = (N, Cc, P, d, pos_sum is not None)
        return emb.view(N, P // N, d)

    @staticmethod
    def backward(ctx, demb):
        xs, y1, c2, out, col2, col3, st1, st2, g1, be1, g2, be2, W2t, W3t, Wpt = ctx.saved_tensors
        N, Cc, P, d, has_pos = ctx.cfg
        dev = demb.device
        f16 = torch.float16
        M = P * 256
        dE = demb.reshape(P, d)
        if not dE.is_contiguous():
            dE = dE.contiguous()
        small = _f32zeros(d + 64 * 5 + 64 * 2 + 64 * Cc * 9, dev)  # [dbp | db3 db2 db1 dg2 dbe2 | dg1 dbe1 | dW1]
        dbp = small[0:d]
        db3, db2, db1v, dg2, dbe2, dg1, dbe1 = (small[d + 64 * i:d + 64 * (i + 1)] for i in range(7))
        dW1 = small[d + 64 * 7:]
        # projection
        ops.colsum(dE, dbp, P, d)
        dWpt = torch.empty(d, 16384, dtype=f16, device=dev)
        ops.gemm(dE, out, dWpt, d, 16384, P, lda=d, ldb=16384, ldc=16384, a_mn=True, b_mn=True)
        dWp = ops.transpose(dWpt, torch.empty(d, 64, 16, 16, dtype=f16, device=dev), d, 256, 64)
        dR = torch.empty(P, 256, 64, dtype=f16, device=dev)  # = d(out), also the residual branch's gradient of y1
        ops.gemm(dE, Wpt, dR, P, 16384, d, lda=d, ldb=16384, ldc=16384, b_mn=True)
        # conv3 (64->64) and the second GroupNorm/GELU
        ops.colsum(dR.view(M, 64), db3, M, 64)
        dW3t = _conv_wgrad(dR, col3, P)
        dcol = torch.empty(M, 576, dtype=f16, device=dev)
        ops.gemm(dR, W3t, dcol, M, 576, 64, lda=64, ldb=576, ldc=576, b_mn=True)
        dc2 = torch.empty(P, 256, 64, dtype=f16, device=dev)
        ops.col2im_gn_gelu_bwd(dcol, c2, st2, g2, be2, None, dc2, dg2, dbe2, P)
        # conv2 and the first GroupNorm/GELU (+ the residual branch)
        ops.colsum(dc2.view(M, 64), db2, M, 64)
        dW2t = _conv_wgrad(dc2, col2, P)
        ops.gemm(dc2, W2t, dcol, M, 576, 64, lda=64, ldb=576, ldc=576, b_mn=True)
        dy1 = torch.empty(P, 256, 64, dtype=f16, device=dev)
        ops.col2im_gn_gelu_bwd(dcol, y1, st1, g1, be1, dR, dy1, dg1, dbe1, P)
        # conv1
        ops.colsum(dy1.view(M, 64), db1v, M, 64)
        ops.patch_conv1_bwd(xs, dy1, dW1, P, Cc)
        sh = _to_half(small, (small.numel(),))
        both = _to_half(torch.cat([dW3t, dW2t]), (2 * 64 * 576,))
        dW3 = ops.transpose(both[:64 * 576], torch.empty(64, 64, 3, 3, dtype=f16, device=dev), 64, 9, 64)
        dW2 = ops.transpose(both[64 * 576:], torch.empty(64, 64, 3, 3, dtype=f16, device=dev), 64, 9, 64)
        o = d
        return (None, dE.view(P, d) if has_pos else None, sh[o + 64 * 7:].view(64, Cc, 3, 3), sh[o + 128:o + 192],
                sh[o + 320:o + 384], sh[o + 384:o + 448], dW2, sh[o + 64:o + 128], sh[o + 192:o + 256],
                sh[o + 256:o + 320], dW3, sh[o:o + 64], dWp, sh[0:d], None, None)


def patch_embed(module, pixel_values, pos_sum):
    """PatchEmbeddings.forward on the sm_100a kernels. pos_sum: [N*n_patch, d] row+column position rows or None."""
    from ._lib import Db1Error
    if not pixel_values.is_cuda:
        raise Db1Error("PatchEmbeddings runs on CUDA fp16 tensors only (no CPU / PyTorch fallback)")
    if pixel_values.dtype not in (torch.float16, torch.float32):
        pixel_values = pixel_values.to(torch.float32)  # uint8 / fp64 frames: standardise from fp32, cast afterwards (:73-79)
    if module.patch_size != 16:
        raise NotImplementedError("sm_100a patch embedder is built for 16x16 patches (the released configuration)")
    rp = module.residual_path
    return PatchEmbedFn.apply(pixel_values, pos_sum, module.conv1.weight, module.conv1.bias, rp[0].weight, rp[0].bias,
                              rp[2].weight, rp[2].bias, rp[3].weight, rp[3].bias, rp[5].weight, rp[5].bias,
                              module.projection.weight, module.projection.bias, rp[0].eps, rp[3].eps)


class DropoutFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p, seed):
        ctx.cfg = (p, seed)
        x = x.contiguous()
        return ops.dropout(x, torch.empty_like(x), p, seed)

    @staticmethod
    def backward(ctx, g):
        p, seed = ctx.cfg
        g = g.contiguous()
        return ops.dropout(g, torch.empty_like(g), p, seed), None, None
