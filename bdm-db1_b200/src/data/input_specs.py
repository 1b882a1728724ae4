"""The model's input contract: one object per task type, tensor fields with a leading batch dimension
(API of src/data/input_specs.py:24-112; `my_collate_fn`, data_samplers.py:28-42, builds exactly these)."""
from dataclasses import dataclass, fields
from typing import List, Optional, Union

import torch


@dataclass
class GatoInputBase:
    position_id: Optional[torch.Tensor]
    attention_mask: Optional[torch.Tensor]
    loss_mask: Optional[torch.Tensor]
    label: Optional[torch.Tensor]

    def _items(self):
        return [(f.name, getattr(self, f.name)) for f in fields(self)]

    def get_datasize(self):
        """Size in GiB of the four base fields."""
        n = 0
        for e in (self.position_id, self.attention_mask, self.loss_mask, self.label):
            if e is not None:
                n += e.element_size() * e.nelement()
        return n / (1024 ** 3)

    def to(self, **kwargs):
        for k, v in self._items():
            if v is not None:
                setattr(self, k, v.to(**kwargs))

    def apply(self, fn, *args, **kwargs):
        for k, v in self._items():
            if v is not None:
                setattr(self, k, fn(v, *args, **kwargs))

    def append(self, other):
        assert type(self).__name__ == type(other).__name__
        for k, v in self._items():
            if v is not None:
                setattr(self, k, torch.cat([v, getattr(other, k)], dim=0))

    @staticmethod
    def merge_into_one(data2merge: List["GatoInputBase"]):
        """Turn every non-None field of the first element into a list collecting that field over all elements."""
        head = data2merge[0]
        head.apply(lambda x: [x])
        for other in data2merge[1:]:
            for k, v in other._items():
                if v is not None:
                    getattr(head, k).append(v)
        return head


@dataclass
class RLTaskInput(GatoInputBase):
    text_seq: Union[List, torch.Tensor, None]
    vision_seq: Union[List, torch.Tensor, None]
    tensor_seq: Union[List, torch.Tensor, None]


@dataclass
class NLPTaskInput(GatoInputBase):
    text_seq: Union[List, torch.Tensor, None]
    text_len: Union[List, torch.Tensor, None]


@dataclass
class ICTaskInput(GatoInputBase):
    """prompt `Caption the image:` + image patches + caption tokens."""
    prompt_seq: Union[List, torch.Tensor, None]
    img_seq: Union[List, torch.Tensor, None]
    text_seq: Union[List, torch.Tensor, None]
    img_id_seq: Union[List, torch.Tensor, None]


@dataclass
class VQATaskInput(GatoInputBase):
    """prompt + image patches + `Question: ... Answer: ...` tokens."""
    prompt_seq: Union[List, torch.Tensor, None]
    img_seq: Union[List, torch.Tensor, None]
    text_seq: Union[List, torch.Tensor, None]
    img_id_seq: Union[List, torch.Tensor, None]
    ques_id_seq: Union[List, torch.Tensor, None]
    ques_len: Union[List, torch.Tensor, None]
