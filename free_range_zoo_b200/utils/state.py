"""Environment state container.

Same surface as the reference base class (free_range_zoo/utils/state.py:12-128): a dataclass of batched tensors with
``to / clone / save_initial / restore_initial / save_checkpoint / restore_from_checkpoint / load_state / stack / cat /
unwrap / to_dataframe`` plus ``__len__ / __getitem__ / __hash__``.  In this
engine the fields are *views of the buffers the kernels update in place*, so ``env.state()`` is always current and
assigning into a field (``state.fires[3] = ...``) edits the live device state.
"""
from __future__ import annotations

import copy
import dataclasses
from typing import List, Optional

import torch


@dataclasses.dataclass(eq=False)  # identity equality, content hash (below) -- like the reference's plain class
class State:
    """Base class for the per-domain states; subclasses list their tensors as dataclass fields."""

    def __post_init__(self) -> None:
        self.metadata = {}
        self.initial_state = None
        self.checkpoint = None

    # -- field helpers
    def _tensor_fields(self):
        shared = self.metadata.get('shared', ())
        return [f.name for f in dataclasses.fields(self) if f.name not in shared]

    def __len__(self) -> int:
        return getattr(self, self._tensor_fields()[0]).shape[0]

    def to(self, device='cpu') -> 'State':
        for f in dataclasses.fields(self):
            value = getattr(self, f.name)
            if isinstance(value, torch.Tensor):
                setattr(self, f.name, value.to(device))
        return self

    def clone(self) -> 'State':
        copied = {}
        for f in dataclasses.fields(self):
            value = getattr(self, f.name)
            copied[f.name] = value.clone() if isinstance(value, torch.Tensor) else copy.deepcopy(value)
        other = type(self)(**copied)
        other.metadata = copy.deepcopy(self.metadata)
        return other

    # -- snapshots (reference state.py:36-92; the reference's restore_initial tests a non-existent attribute
    #    `self.initial` (:47) -- implemented here to its evident intent)
    def save_initial(self) -> None:
        self.initial_state = self.clone()

    def save_checkpoint(self) -> None:
        self.checkpoint = self.clone()

    def _restore(self, source: Optional['State'], batch_indices, what: str) -> None:
        if source is None:
            raise ValueError(f'{what} is not saved')
        for name in self._tensor_fields():
            current, saved = getattr(self, name), getattr(source, name)
            if batch_indices is None:
                current.copy_(saved)
            else:
                current[batch_indices] = saved[batch_indices]

    def restore_initial(self, batch_indices: Optional[torch.Tensor] = None) -> None:
        self._restore(self.initial_state, batch_indices, 'Initial state')

    def restore_from_checkpoint(self, batch_indices: Optional[torch.Tensor] = None) -> None:
        self._restore(self.checkpoint, batch_indices, 'Checkpoint')

    def load_state(self, state: 'State', batch_indices: Optional[torch.Tensor] = None) -> None:
        for name in self._tensor_fields():
            current, incoming = getattr(self, name), getattr(state, name)
            if batch_indices is None:
                current.copy_(incoming)
            else:
                current[batch_indices] = incoming.to(current.device)

    # -- batching helpers (reference state.py:130-180)
    @staticmethod
    def _combine(states: List['State'], op, *args, **kwargs) -> 'State':
        first = states[0]
        shared = first.metadata.get('shared', ())
        merged = {}
        for f in dataclasses.fields(first):
            if f.name in shared:
                merged[f.name] = getattr(first, f.name)
            else:
                merged[f.name] = op([getattr(s, f.name) for s in states], *args, **kwargs)
        return type(first)(**merged)

    @staticmethod
    def stack(states: List['State'], *args, **kwargs) -> 'State':
        return State._combine(states, torch.stack, *args, **kwargs)

    @staticmethod
    def cat(states: List['State'], *args, **kwargs) -> 'State':
        return State._combine(states, torch.cat, *args, **kwargs)

    # -- per-environment access (reference state.py:180-238)
    def unwrap(self) -> List['State']:
        """One single-environment state per batch entry (``self[i]`` for every i)."""
        return [self[index] for index in range(len(self))]

    def __getitem__(self, indices) -> 'State':
        shared = self.metadata.get('shared', ())
        return type(self)(**{f.name: getattr(self, f.name) if f.name in shared else getattr(self, f.name)[indices]
                             for f in dataclasses.fields(self)})

    def to_dataframe(self):
        """One row per environment, every field as the string of its nested list; shared fields repeated on every
        row (reference state.py:180-191 -- the CSV logging format)."""
        import pandas as pd
        shared = self.metadata.get('shared', ())
        data = {name: [str(row.tolist()) for row in getattr(self, name)] for name in self._tensor_fields()}
        frame = pd.DataFrame(data)
        for name in shared:
            frame[name] = str(getattr(self, name).tolist())
        return frame

    def __hash__(self) -> int:
        """Hash of the field contents (the reference hashes every tensor with xxhash, utils/caching.py:17-32; here the
        raw bytes are hashed -- equal states hash equal, which is all the contract asks for)."""
        return hash(tuple(getattr(self, f.name).detach().cpu().contiguous().numpy().tobytes()
                          for f in dataclasses.fields(self)))
