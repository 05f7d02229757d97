"""Observation / mapping containers.

``ObservationDict`` stands in for the ``tensordict.TensorDict`` the reference returns from ``observe``
(free_range_zoo/envs/wildfire/env/wildfire.py:709-717): a mapping with ``batch_size`` and ``device``.  Values may be
registered as thunks; they are evaluated on first access and cached, so views that need a gather or a jagged
re-packing (``others``, ``tasks``) cost nothing on the step path unless a caller actually reads them.
"""
from __future__ import annotations

from collections.abc import MutableMapping
from typing import Callable, Dict, Iterator

import torch


class LazyDict(MutableMapping):
    """dict whose values can be zero-argument callables, resolved (once) on access."""

    def __init__(self, source: Dict | None = None):
        self._store: Dict = dict(source or {})

    def __getitem__(self, key):
        value = self._store[key]
        if callable(value) and not isinstance(value, torch.Tensor):
            value = value()
            self._store[key] = value
        return value

    def __setitem__(self, key, value) -> None:
        self._store[key] = value

    def __delitem__(self, key) -> None:
        del self._store[key]

    def __iter__(self) -> Iterator:
        return iter(self._store)

    def __len__(self) -> int:
        return len(self._store)

    def __repr__(self) -> str:
        return f'{type(self).__name__}(keys={list(self._store)})'


class ObservationDict(LazyDict):
    """Mapping with the two TensorDict attributes callers rely on: ``batch_size`` and ``device``."""

    def __init__(self, source: Dict | None = None, batch_size=None, device=None):
        super().__init__(source)
        self.batch_size = torch.Size(batch_size if batch_size is not None else [])
        self.device = device

    def to(self, device) -> 'ObservationDict':
        return ObservationDict({k: v.to(device) for k, v in self.items()}, self.batch_size, device)

    def clone(self) -> 'ObservationDict':
        return ObservationDict({k: v.clone() for k, v in self.items()}, self.batch_size, self.device)


def jagged_from_padded(padded: torch.Tensor, counts: torch.Tensor, dtype: torch.dtype | None = None) -> torch.Tensor:
    """Dense ``[B, T, ...]`` + per-row lengths ``[B]`` -> ``torch.nested`` jagged tensor ``[B, j, ...]``.

    The reference builds the same object with ``as_nested_tensor(x.split(counts.tolist()))`` -- one Python tensor
    per environment (wildfire.py:630-654,693-697).  Here it is one boolean compaction plus one cumsum; the only
    host synchronisation is the data-dependent size of the compacted values.
    """
    steps = torch.arange(padded.shape[1], device=padded.device)
    keep = steps.unsqueeze(0) < counts.unsqueeze(1)
    values = padded[keep]
    if dtype is not None:
        values = values.to(dtype)
    return _jagged(values, counts)


def _jagged(values: torch.Tensor, counts: torch.Tensor) -> torch.Tensor:
    offsets = torch.zeros(counts.shape[0] + 1, dtype=torch.int64, device=values.device)
    offsets[1:] = torch.cumsum(counts, dim=0)
    longest = int(counts.max().item()) if counts.numel() else 0
    return torch.nested.nested_tensor_from_jagged(values, offsets=offsets, min_seqlen=0, max_seqlen=max(longest, 1))


def jagged_indices_from_mask(mask: torch.Tensor) -> torch.Tensor:
    """Boolean ``[B, T]`` -> jagged int64 ``[B, j]`` holding the positions of the set entries of each row."""
    return _jagged(mask.nonzero(as_tuple=False)[:, 1], mask.sum(dim=1))


def jagged_rows_from_mask(values: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """Dense ``[B, K, C]`` + boolean ``[B, K]`` -> jagged ``[B, j, C]`` keeping the selected rows of each batch entry
    in order (the per-agent task observations of rideshare, rideshare.py:436-455)."""
    return _jagged(values[mask], mask.sum(dim=1))
