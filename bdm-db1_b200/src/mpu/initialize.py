"""Parallel-state registry (API of src/mpu/initialize.py:49-398).

Rank layout is Megatron's: with world = dp * pp * tp, tensor-parallel ranks are adjacent, then data-parallel, then
pipeline stages. `group_layout` is a pure function so the layout can be tested without a process group.
"""
import torch

from .utils import ensure_divisibility

__all__ = [
    "is_unitialized", "initialize_model_parallel", "model_parallel_is_initialized", "get_model_parallel_group",
    "get_tensor_model_parallel_group", "get_pipeline_model_parallel_group", "get_data_parallel_group",
    "get_embedding_group", "set_tensor_model_parallel_world_size", "set_pipeline_model_parallel_world_size",
    "get_tensor_model_parallel_world_size", "get_model_parallel_world_size", "get_pipeline_model_parallel_world_size",
    "set_tensor_model_parallel_rank", "set_pipeline_model_parallel_rank", "get_tensor_model_parallel_rank",
    "get_model_parallel_rank", "get_pipeline_model_parallel_rank", "is_pipeline_first_stage", "is_pipeline_last_stage",
    "get_virtual_pipeline_model_parallel_rank", "set_virtual_pipeline_model_parallel_rank",
    "get_virtual_pipeline_model_parallel_world_size", "get_tensor_model_parallel_src_rank",
    "get_pipeline_model_parallel_first_rank", "get_pipeline_model_parallel_last_rank",
    "get_pipeline_model_parallel_next_rank", "get_pipeline_model_parallel_prev_rank", "get_data_parallel_world_size",
    "get_data_parallel_rank", "destroy_model_parallel", "group_layout",
]


class _State:
    def __init__(self):
        self.reset()

    def reset(self):
        self.groups = {"dp": None, "mp": None, "tp": None, "pp": None, "emb": None}
        self.pp_ranks = None
        self.vpp_rank = None
        self.vpp_world = None
        self.forced = {}  # explicit overrides set through the set_* functions


_S = _State()


def group_layout(world_size, tp=1, pp=1):
    """Rank lists of every group kind: {'dp': [[...], ...], 'mp': ..., 'tp': ..., 'pp': ..., 'emb': ...}."""
    tp = min(tp, world_size)
    pp = min(pp, world_size)
    ensure_divisibility(world_size, tp * pp)
    dp = world_size // (tp * pp)
    per_stage = world_size // pp
    out = {"dp": [], "mp": [], "tp": [], "pp": [], "emb": []}
    for stage in range(pp):
        lo, hi = stage * per_stage, (stage + 1) * per_stage
        for t in range(tp):
            out["dp"].append(list(range(lo + t, hi, tp)))
    for i in range(dp):
        out["mp"].append([g[i] for g in out["dp"]])
    for i in range(world_size // tp):
        out["tp"].append(list(range(i * tp, (i + 1) * tp)))
    for i in range(per_stage):
        ranks = list(range(i, world_size, per_stage))
        out["pp"].append(ranks)
        out["emb"].append([ranks[0], ranks[-1]] if len(ranks) > 1 else ranks)
    return out


def is_unitialized():
    return _S.groups["dp"] is None


def initialize_model_parallel(tensor_model_parallel_size_=1, pipeline_model_parallel_size_=1,
                              virtual_pipeline_model_parallel_size_=None):
    assert torch.distributed.is_initialized()
    assert _S.groups["dp"] is None, "data parallel group is already initialized"
    world = torch.distributed.get_world_size()
    rank = torch.distributed.get_rank()
    if virtual_pipeline_model_parallel_size_ is not None:
        _S.vpp_rank = 0
        _S.vpp_world = virtual_pipeline_model_parallel_size_
    layout = group_layout(world, tensor_model_parallel_size_, pipeline_model_parallel_size_)
    # every rank must create every group, in the same order
    for kind in ("dp", "mp", "tp", "pp", "emb"):
        for ranks in layout[kind]:
            g = torch.distributed.new_group(ranks)
            if rank in ranks:
                _S.groups[kind] = g
                if kind == "pp":
                    _S.pp_ranks = ranks


def model_parallel_is_initialized():
    return not (_S.groups["tp"] is None or _S.groups["pp"] is None or _S.groups["dp"] is None)


def _group(kind, what):
    g = _S.groups[kind]
    assert g is not None, "%s is not initialized" % what
    return g


def get_model_parallel_group():
    return _group("mp", "model parallel group")


def get_tensor_model_parallel_group():
    return _group("tp", "intra_layer_model parallel group")


def get_pipeline_model_parallel_group():
    return _group("pp", "pipeline_model parallel group")


def get_data_parallel_group():
    return _group("dp", "data parallel group")


def get_embedding_group():
    return _group("emb", "embedding group")


def set_tensor_model_parallel_world_size(world_size):
    _S.forced["tp_world"] = world_size


def set_pipeline_model_parallel_world_size(world_size):
    _S.forced["pp_world"] = world_size


def get_tensor_model_parallel_world_size():
    if _S.forced.get("tp_world") is not None:
        return _S.forced["tp_world"]
    return torch.distributed.get_world_size(group=get_tensor_model_parallel_group())


def get_model_parallel_world_size():
    assert get_pipeline_model_parallel_world_size() == 1, "legacy get_model_parallel_world_size is only supported if PP is disabled"
    return get_tensor_model_parallel_world_size()


def get_pipeline_model_parallel_world_size():
    if _S.forced.get("pp_world") is not None:
        return _S.forced["pp_world"]
    return torch.distributed.get_world_size(group=get_pipeline_model_parallel_group())


def set_tensor_model_parallel_rank(rank):
    _S.forced["tp_rank"] = rank


def set_pipeline_model_parallel_rank(rank):
    _S.forced["pp_rank"] = rank


def get_tensor_model_parallel_rank():
    if _S.forced.get("tp_rank") is not None:
        return _S.forced["tp_rank"]
    return torch.distributed.get_rank(group=get_tensor_model_parallel_group())


def get_model_parallel_rank():
    assert get_pipeline_model_parallel_world_size() == 1, "legacy get_model_parallel_rank is only supported if PP is disabled"
    return get_tensor_model_parallel_rank()


def get_pipeline_model_parallel_rank():
    if _S.forced.get("pp_rank") is not None:
        return _S.forced["pp_rank"]
    return torch.distributed.get_rank(group=get_pipeline_model_parallel_group())


def is_pipeline_first_stage(ignore_virtual=False):
    if not ignore_virtual and get_virtual_pipeline_model_parallel_world_size() is not None \
            and get_virtual_pipeline_model_parallel_rank() != 0:
        return False
    return get_pipeline_model_parallel_rank() == 0


def is_pipeline_last_stage(ignore_virtual=False):
    if not ignore_virtual:
        vw = get_virtual_pipeline_model_parallel_world_size()
        if vw is not None and get_virtual_pipeline_model_parallel_rank() != vw - 1:
            return False
    return get_pipeline_model_parallel_rank() == get_pipeline_model_parallel_world_size() - 1


def get_virtual_pipeline_model_parallel_rank():
    return _S.vpp_rank


def set_virtual_pipeline_model_parallel_rank(rank):
    _S.vpp_rank = rank


def get_virtual_pipeline_model_parallel_world_size():
    return _S.vpp_world


def get_tensor_model_parallel_src_rank():
    tp = get_tensor_model_parallel_world_size()
    return (torch.distributed.get_rank() // tp) * tp


def get_pipeline_model_parallel_first_rank():
    assert _S.pp_ranks is not None, "Pipeline parallel group is not initialized"
    return _S.pp_ranks[0]


def get_pipeline_model_parallel_last_rank():
    assert _S.pp_ranks is not None, "Pipeline parallel group is not initialized"
    return _S.pp_ranks[get_pipeline_model_parallel_world_size() - 1]


def get_pipeline_model_parallel_next_rank():
    assert _S.pp_ranks is not None, "Pipeline parallel group is not initialized"
    return _S.pp_ranks[(get_pipeline_model_parallel_rank() + 1) % get_pipeline_model_parallel_world_size()]


def get_pipeline_model_parallel_prev_rank():
    assert _S.pp_ranks is not None, "Pipeline parallel group is not initialized"
    return _S.pp_ranks[(get_pipeline_model_parallel_rank() - 1) % get_pipeline_model_parallel_world_size()]


def get_data_parallel_world_size():
    return torch.distributed.get_world_size(group=get_data_parallel_group())


def get_data_parallel_rank():
    return torch.distributed.get_rank(group=get_data_parallel_group())


def destroy_model_parallel():
    _S.reset()
