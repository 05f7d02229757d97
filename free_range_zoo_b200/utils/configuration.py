"""Base class of the openness configuration structures.

Mirrors the contract of the reference base (free_range_zoo/utils/configuration.py:10-44): every configuration
validates itself on construction, validates nested configurations recursively, and ``to(device)`` moves every tensor
field.  The structures stay plain dataclasses because they are the user-facing way to describe an environment; the
engine flattens them ONCE at construction into the POD ``Frz*Params`` structs of ``include/frz.h``.
"""
from __future__ import annotations

import dataclasses
from typing import Any

import torch


class Configuration:
    """Mixin for dataclass configurations (validate-on-init, recursive validate, tensor relocation)."""

    def __post_init__(self) -> None:
        self.validate()

    def validate(self) -> bool:
        for value in vars(self).values():
            if isinstance(value, Configuration):
                value.validate()
        return True

    def to(self, device: Any = 'cpu') -> 'Configuration':
        for name, value in list(vars(self).items()):
            if isinstance(value, (torch.Tensor, Configuration)):
                setattr(self, name, value.to(device))
        return self

    def fields(self):
        return [f.name for f in dataclasses.fields(self)]


def require(condition: bool, message: str) -> None:
    """Raise the reference's error type (ValueError) for an inconsistent configuration."""
    if not condition:
        raise ValueError(message)


def in_unit_interval(value: float) -> bool:
    return 0.0 <= float(value) <= 1.0
