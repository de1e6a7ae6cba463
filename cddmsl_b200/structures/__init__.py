"""Minimal containers so the host-side mirror runs stand-alone.  In an integration the reference's own
`detectron2.structures.{Boxes,Instances}` are used instead (anything with `.tensor` / attribute access
works); these are not part of the hot path (out of scope, kept as-is upstream)."""
from __future__ import annotations

from typing import Any, Dict, List, Tuple, Union

import torch


class Boxes:
    """xyxy boxes, Nx4 (detectron2/structures/boxes.py:130-260 subset)."""

    def __init__(self, tensor: torch.Tensor):
        if not isinstance(tensor, torch.Tensor):
            tensor = torch.as_tensor(tensor, dtype=torch.float32)
        tensor = tensor.to(torch.float32)
        if tensor.numel() == 0:
            tensor = tensor.reshape((-1, 4))
        assert tensor.dim() == 2 and tensor.size(-1) == 4, tensor.size()
        self.tensor = tensor

    def clip(self, box_size: Tuple[int, int]) -> None:
        h, w = box_size
        x1 = self.tensor[:, 0].clamp(min=0, max=w)
        y1 = self.tensor[:, 1].clamp(min=0, max=h)
        x2 = self.tensor[:, 2].clamp(min=0, max=w)
        y2 = self.tensor[:, 3].clamp(min=0, max=h)
        self.tensor = torch.stack((x1, y1, x2, y2), dim=-1)

    def nonempty(self, threshold: float = 0.0) -> torch.Tensor:
        b = self.tensor
        return ((b[:, 2] - b[:, 0]) > threshold) & ((b[:, 3] - b[:, 1]) > threshold)

    def area(self) -> torch.Tensor:
        b = self.tensor
        return (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])

    def __getitem__(self, item) -> "Boxes":
        if isinstance(item, int):
            return Boxes(self.tensor[item].view(1, -1))
        return Boxes(self.tensor[item])

    def __len__(self) -> int:
        return self.tensor.shape[0]

    def to(self, *a, **k) -> "Boxes":
        return Boxes(self.tensor.to(*a, **k))

    @property
    def device(self):
        return self.tensor.device

    @staticmethod
    def cat(boxes_list: List["Boxes"]) -> "Boxes":
        if len(boxes_list) == 0:
            return Boxes(torch.empty(0, 4))
        return Boxes(torch.cat([b.tensor for b in boxes_list], dim=0))


class Instances:
    """Per-image bag of equally long fields (detectron2/structures/instances.py subset)."""

    def __init__(self, image_size: Tuple[int, int], **kwargs: Any):
        object.__setattr__(self, "_image_size", image_size)
        object.__setattr__(self, "_fields", {})
        for k, v in kwargs.items():
            self.set(k, v)

    @property
    def image_size(self) -> Tuple[int, int]:
        return self._image_size

    def __setattr__(self, name: str, val: Any) -> None:
        if name.startswith("_"):
            object.__setattr__(self, name, val)
        else:
            self.set(name, val)

    def __getattr__(self, name: str) -> Any:
        fields = object.__getattribute__(self, "_fields")
        if name not in fields:
            raise AttributeError(f"Cannot find field '{name}' in the given Instances!")
        return fields[name]

    def set(self, name: str, value: Any) -> None:
        if len(self._fields):
            assert len(self) == len(value), f"Adding a field of length {len(value)} to Instances of length {len(self)}"
        self._fields[name] = value

    def has(self, name: str) -> bool:
        return name in self._fields

    def get(self, name: str) -> Any:
        return self._fields[name]

    def get_fields(self) -> Dict[str, Any]:
        return self._fields

    def to(self, *a, **k) -> "Instances":
        ret = Instances(self._image_size)
        for kk, v in self._fields.items():
            ret.set(kk, v.to(*a, **k) if hasattr(v, "to") else v)
        return ret

    def __getitem__(self, item: Union[int, slice, torch.Tensor]) -> "Instances":
        if isinstance(item, int):
            item = slice(item, None if item == -1 else item + 1)
        ret = Instances(self._image_size)
        for k, v in self._fields.items():
            ret.set(k, v[item])
        return ret

    def __len__(self) -> int:
        for v in self._fields.values():
            return len(v)
        raise NotImplementedError("Empty Instances does not support __len__!")

    @staticmethod
    def cat(instance_lists: List["Instances"]) -> "Instances":
        """detectron2/structures/instances.py:151-186"""
        assert all(isinstance(i, Instances) for i in instance_lists) and len(instance_lists) > 0
        if len(instance_lists) == 1:
            return instance_lists[0]
        image_size = instance_lists[0].image_size
        for i in instance_lists[1:]:
            assert i.image_size == image_size
        ret = Instances(image_size)
        for k in instance_lists[0]._fields.keys():
            values = [i.get(k) for i in instance_lists]
            v0 = values[0]
            if isinstance(v0, torch.Tensor):
                values = torch.cat(values, dim=0)
            elif isinstance(v0, list):
                values = [x for v in values for x in v]
            elif hasattr(type(v0), "cat"):
                values = type(v0).cat(values)
            else:
                raise ValueError(f"Unsupported type {type(v0)} for concatenation")
            ret.set(k, values)
        return ret
