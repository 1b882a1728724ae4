"""Activation modules referenced by name from the config (API of src/model/activations.py, src/model/utils.py:22-28).
They are parameter-free placeholders inside `PositionwiseFF.CoreNet`: the fused GEMM epilogue computes the activation."""
import torch.nn as nn


class GEGLU(nn.Module):
    """value * gelu_erf(gate) with [value | gate] = chunk(x, 2, -1) (src/model/activations.py:26-29)."""

    def forward(self, x):
        raise RuntimeError("GEGLU is evaluated inside the fused sm_100a GEMM epilogue; call PositionwiseFF.forward")
