"""ResNet-v2-block image patch embedder with learned row/column position embeddings
(module surface of src/tokenizer/vision_embedding.py:36-180; computed by db1_sm100 kernels)."""
import collections.abc

import torch
import torch.nn as nn

from db1_sm100 import functions as F_

CLASS_TOKEN_LENGTH = 0


def to_2tuple(x):
    if isinstance(x, collections.abc.Iterable):
        return x
    return (x, x)


class PatchEmbeddings(nn.Module):
    """Per-patch standardisation -> conv3x3(C->64) -> [GroupNorm(32) -> GELU -> conv3x3(64->64)] x2 -> + residual
    -> conv(patch x patch, stride patch) (reference :36-86). Parameters keep the reference's names and shapes."""

    def __init__(self, patch_size=16, num_channels=3, embed_dim=768, data_type=torch.half):
        super().__init__()
        self.patch_size = patch_size
        self.embed_dim = embed_dim
        self.data_type = data_type
        ps = to_2tuple(patch_size)
        self.conv1 = nn.Conv2d(num_channels, 64, kernel_size=3, stride=1, padding=1, dtype=data_type)
        self.projection = nn.Conv2d(64, embed_dim, kernel_size=ps, stride=ps, dtype=data_type)
        self.residual_path = nn.Sequential(
            nn.GroupNorm(num_groups=32, num_channels=64, dtype=data_type),
            nn.GELU(),
            nn.Conv2d(64, 64, kernel_size=3, stride=1, padding=1, dtype=data_type),
            nn.GroupNorm(num_groups=32, num_channels=64, dtype=data_type),
            nn.GELU(),
            nn.Conv2d(64, 64, kernel_size=3, stride=1, padding=1, dtype=data_type),
        )

    def forward(self, pixel_values, pos_sum=None):
        """pixel_values [N,C,H,W] -> [N, (H/p)*(W/p), embed_dim]; pos_sum (optional [N*n_patch, embed_dim]) is added in
        the projection GEMM's epilogue."""
        return F_.patch_embed(self, pixel_values, pos_sum)


class VisionEmbedding(nn.Module):
    def __init__(self, config):
        super().__init__()
        data_type = torch.half if config.fp16 else torch.float32
        self.data_type = data_type
        self.patch_embeddings = PatchEmbeddings(patch_size=config.vision_patch_size,
                                                num_channels=config.vision_num_input_channels,
                                                embed_dim=config.n_embed, data_type=data_type)
        self.row_position_embeddings = nn.Embedding(config.vision_position_vocab_size, config.n_embed, dtype=data_type)
        self.col_position_embeddings = nn.Embedding(config.vision_position_vocab_size, config.n_embed, dtype=data_type)
        self.dropout = nn.Dropout(config.vision_hidden_dropout_prob)
        self.config = config

    def position_indices(self, height, width, batch_size, device):
        """Row / column table indices per (sample, patch) (reference :130-172): the patch's interval on a
        `vision_position_vocab_size`-step axis; its midpoint in eval mode, a uniform draw from it in training.
        Integer index bookkeeping only; float32 division then truncation exactly as the reference."""
        ps = self.config.vision_patch_size
        vocab = self.config.vision_position_vocab_size
        h0, w0 = height // ps, width // ps
        seq = torch.arange(h0 * w0, device=device)
        row = torch.div(seq, w0, rounding_mode="trunc")
        col = seq % w0
        col_hi = ((col + 1) / w0 * vocab).to(torch.int32)
        col_lo = (col / w0 * vocab).to(torch.int32)
        row_hi = ((row + 1) / h0 * vocab).to(torch.int32)
        row_lo = (row / h0 * vocab).to(torch.int32)
        if self.training:
            # uniform integer in [lo, hi) per (sample, patch), drawn on the device without host syncs
            ur = torch.rand(batch_size, h0 * w0, device=device)
            uc = torch.rand(batch_size, h0 * w0, device=device)
            r = (row_lo[None] + (ur * (row_hi - row_lo)[None]).floor().to(torch.int32))
            c = (col_lo[None] + (uc * (col_hi - col_lo)[None]).floor().to(torch.int32))
            r = torch.minimum(r, (row_hi - 1)[None].clamp(min=0))
            c = torch.minimum(c, (col_hi - 1)[None].clamp(min=0))
        else:
            r = ((row_lo + row_hi) / 2).int()[None].expand(batch_size, -1)
            c = ((col_lo + col_hi) / 2).int()[None].expand(batch_size, -1)
        return r.long().contiguous(), c.long().contiguous()

    def forward(self, pixel_values):
        n, _c, height, width = pixel_values.shape
        r, c = self.position_indices(height, width, n, pixel_values.device)
        # row_emb[r] + col_emb[c] through the embedding-assembly kernel (W = row table, T = column table)
        pos_sum = F_.EmbedFn.apply(r, c, self.row_position_embeddings.weight, self.col_position_embeddings.weight,
                                   None, 0.0)
        return self.patch_embeddings(pixel_values, pos_sum.reshape(-1, pos_sum.size(-1)))
