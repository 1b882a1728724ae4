"""Process-group registry with the reference's `src.mpu` API (src/mpu/__init__.py:14-32, initialize.py, utils.py).

DB1 trains data-parallel only (README.md:129, TP = PP = 1 in scripts/evaluate/evaluate_rl_1.2B.sh:14-15), so the
data-parallel group is the only one with more than one rank; the tensor / pipeline / embedding getters are kept because
the engine protocol (DeepSpeed's `mpu=` argument) calls them.
"""
import torch

from .initialize import *  # noqa: F401,F403
from .utils import VocabUtility, divide, ensure_divisibility, split_tensor_along_last_dim  # noqa: F401


def print_rank_0(message):
    """Print once per job (src/mpu/__init__.py:19-25)."""
    if not torch.distributed.is_initialized() or torch.distributed.get_rank() == 0:
        print(message, flush=True)


def print_with_rank(message):
    """Print prefixed with the global rank (src/mpu/__init__.py:27-32)."""
    if torch.distributed.is_initialized():
        print("rank: %d" % torch.distributed.get_rank(), message, flush=True)
    else:
        print(message, flush=True)
