"""a1: the task-input dataclasses behave like the reference's (src/data/input_specs.py:24-112) - field sets,
.to / .apply / .append / merge_into_one (incl. its quirks: None fields stay None, merge_into_one wraps the first element's
tensors in lists in place and appends the others' non-None fields). Checked directly against the unmodified reference
module under oracle/_ref/reference when it is installed (build container), and against stated expectations otherwise."""
import dataclasses
import importlib.util
import os

import pytest
import torch

from src.data import input_specs as ours

REF_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "reference", "src",
                        "data", "input_specs.py")


def _ref():
    if not os.path.exists(REF_PATH):
        return None
    spec = importlib.util.spec_from_file_location("ref_input_specs", REF_PATH)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _mk(mod, seed, rows, with_vision):
    g = torch.Generator().manual_seed(seed)
    return mod.RLTaskInput(
        position_id=torch.randint(0, 9, (rows, 6), generator=g), attention_mask=None,
        loss_mask=torch.rand(rows, 6, generator=g), label=torch.randint(0, 50, (rows, 6), generator=g),
        text_seq=None, vision_seq=torch.rand(rows, 2, 3, 16, 16, generator=g) if with_vision else None,
        tensor_seq=torch.randint(0, 50, (rows, 6), generator=g))


def _same(a, b):
    fa, fb = dataclasses.asdict(a), dataclasses.asdict(b)
    assert list(fa) == list(fb)
    for k in fa:
        va, vb = fa[k], fb[k]
        if va is None or vb is None:
            assert va is None and vb is None, k
        elif isinstance(va, list):
            assert isinstance(vb, list) and len(va) == len(vb), k
            for x, y in zip(va, vb):
                assert torch.equal(x, y), k
        else:
            assert va.dtype == vb.dtype and torch.equal(va, vb), k


@pytest.mark.parametrize("cls", ["RLTaskInput", "NLPTaskInput", "ICTaskInput", "VQATaskInput"])
def test_field_sets_match_the_reference(cls):
    expect = {"RLTaskInput": ["position_id", "attention_mask", "loss_mask", "label", "text_seq", "vision_seq", "tensor_seq"],
              "NLPTaskInput": ["position_id", "attention_mask", "loss_mask", "label", "text_seq", "text_len"],
              "ICTaskInput": ["position_id", "attention_mask", "loss_mask", "label", "prompt_seq", "img_seq", "text_seq",
                              "img_id_seq"],
              "VQATaskInput": ["position_id", "attention_mask", "loss_mask", "label", "prompt_seq", "img_seq", "text_seq",
                               "img_id_seq", "ques_id_seq", "ques_len"]}[cls]
    assert [f.name for f in dataclasses.fields(getattr(ours, cls))] == expect
    ref = _ref()
    if ref is not None:
        assert [f.name for f in dataclasses.fields(getattr(ref, cls))] == expect


@pytest.mark.parametrize("with_vision", [False, True])
def test_append_apply_to_and_merge_into_one(with_vision):
    ref = _ref()
    mods = [ours] + ([ref] if ref is not None else [])
    results = []
    for mod in mods:
        a, b, c = _mk(mod, 1, 2, with_vision), _mk(mod, 2, 3, with_vision), _mk(mod, 3, 1, with_vision)
        a.append(b)  # concatenation along dim 0 of every non-None field; None stays None
        assert a.tensor_seq.shape[0] == 5 and a.attention_mask is None and a.text_seq is None
        assert (a.vision_seq is None) == (not with_vision)
        a.apply(lambda t: t * 2 if t.is_floating_point() else t + 1)
        a.to(dtype=torch.float64)
        assert a.label.dtype == torch.float64
        merged = mod.GatoInputBase.merge_into_one([a, c])
        assert merged is a and isinstance(merged.tensor_seq, list) and len(merged.tensor_seq) == 2
        assert merged.attention_mask is None
        assert abs(c.get_datasize() - sum(v.element_size() * v.nelement() for v in
                                         (c.position_id, c.loss_mask, c.label)) / 1024 ** 3) < 1e-12
        results.append(merged)
    if len(results) == 2:
        _same(results[0], results[1])
