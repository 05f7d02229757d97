"""Action / observation space objects.

The reference represents spaces with the external Rust extension ``free_range_rust`` (pyproject.toml:15;
``Space.Discrete / OneOf / Vector / Box / Tuple / Dict``) and REBUILDS a ``Space.Vector`` of ``parallel_envs``
``OneOf`` objects from ``counts.tolist()`` for every agent on every step
(envs/wildfire/env/spaces/actions.py:10-41, envs/rideshare/env/spaces/actions.py:10-50,
envs/cybersecurity/env/spaces/actions.py:11-99).  ``free_range_rust`` performs no transition arithmetic -- only
construction, equality and sampling -- and is not installable here, so this module provides:

* ``Space``: a small pure-Python structural equivalent (same constructors, ``.spaces/.n/.start/.low/.high``,
  ``==``, ``sample()``, ``sample_nested()``) whose expected constructions match the reference's space unit tests
  (tests/free_range_zoo/envs/*/env/spaces/test_action_space.py).
* ``BatchedActionSpace``: what ``env.action_space(agent)`` returns.  It keeps the per-environment choice table on the
  device and only materialises Python ``Space`` objects if a caller asks for ``.spaces``; ``sample_nested()`` draws
  for the whole batch with three tensor ops.  The zero-round-trip path for rollouts is ``env.sample_actions()``.
"""
from __future__ import annotations

import functools
import random
from typing import Dict as TDict, List, Sequence

import torch


class Space:
    """Structural space description: kind + payload, comparable and hashable."""

    __slots__ = ('kind', 'payload')

    def __init__(self, kind: str, payload):
        self.kind = kind
        self.payload = payload

    # -- constructors (same call signatures as free_range_rust.Space)
    @staticmethod
    def Discrete(n: int, start: int = 0) -> 'Space':
        return Space('Discrete', (int(n), int(start)))

    @staticmethod
    def Box(low: Sequence, high: Sequence) -> 'Space':
        return Space('Box', (tuple(low), tuple(high)))

    @staticmethod
    def OneOf(spaces: Sequence['Space']) -> 'Space':
        return Space('OneOf', tuple(spaces))

    @staticmethod
    def Tuple(spaces: Sequence['Space']) -> 'Space':
        return Space('Tuple', tuple(spaces))

    @staticmethod
    def Vector(spaces: Sequence['Space']) -> 'Space':
        return Space('Vector', tuple(spaces))

    @staticmethod
    def Dict(spaces: TDict[str, 'Space']) -> 'Space':
        return Space('Dict', tuple(sorted(spaces.items())))

    # -- accessors
    @property
    def spaces(self):
        if self.kind == 'Dict':
            return dict(self.payload)
        if self.kind in ('OneOf', 'Tuple', 'Vector'):
            return list(self.payload)
        raise AttributeError(f'{self.kind} space has no sub-spaces')

    @property
    def n(self) -> int:
        return self.payload[0]

    @property
    def start(self) -> int:
        return self.payload[1]

    @property
    def low(self):
        return list(self.payload[0])

    @property
    def high(self):
        return list(self.payload[1])

    def __len__(self) -> int:
        return len(self.payload)

    def __eq__(self, other) -> bool:
        return isinstance(other, Space) and self.kind == other.kind and self.payload == other.payload

    def __hash__(self) -> int:
        return hash((self.kind, self.payload))

    def __repr__(self) -> str:
        if self.kind == 'Discrete':
            return f'Discrete({self.n}, start={self.start})'
        return f'{self.kind}({list(self.payload)!r})'

    # -- sampling
    def sample(self):
        if self.kind == 'Discrete':
            return self.start + random.randrange(self.n)
        if self.kind == 'Box':
            return [random.uniform(lo, hi) for lo, hi in zip(*self.payload)]
        if self.kind == 'OneOf':
            index = random.randrange(len(self.payload))
            return index, self.payload[index].sample()
        if self.kind == 'Dict':
            return {key: space.sample() for key, space in self.payload}
        return [space.sample() for space in self.payload]

    def sample_nested(self):
        """Flat ``[choice index, value]`` for ``OneOf``; element-wise for containers (wrappers/space_validator.py)."""
        if self.kind == 'Discrete':
            return [self.sample()]
        if self.kind == 'OneOf':
            index = random.randrange(len(self.payload))
            return [index, *self.payload[index].sample_nested()]
        if self.kind == 'Dict':
            return {key: space.sample_nested() for key, space in self.payload}
        if self.kind == 'Box':
            return self.sample()
        return [space.sample_nested() for space in self.payload]


@functools.lru_cache(maxsize=256)
def one_of_discrete(starts: tuple) -> Space:
    """``OneOf([Discrete(1, start=s) for s in starts])`` -- the shape of every per-environment action space."""
    return Space.OneOf([Space.Discrete(1, start=s) for s in starts])


class BatchedActionSpace:
    """``Space.Vector`` of per-environment ``OneOf([Discrete(1, start=...), ...])`` kept as device tensors.

    starts: int32 [B, C] -- ``start`` of the c-th choice of each environment (padded); counts: int32 [B] -- number of
    valid choices (>= 1: the noop choice is always there).
    """

    kind = 'Vector'

    def __init__(self, starts: torch.Tensor, counts: torch.Tensor):
        self.starts = starts
        self.counts = counts

    def __len__(self) -> int:
        return self.counts.shape[0]

    @functools.cached_property
    def spaces(self) -> List[Space]:
        starts, counts = self.starts.tolist(), self.counts.tolist()
        return [one_of_discrete(tuple(row[:n])) for row, n in zip(starts, counts)]

    def __eq__(self, other) -> bool:
        if isinstance(other, BatchedActionSpace):
            return self.spaces == other.spaces
        return isinstance(other, Space) and other.kind == 'Vector' and self.spaces == other.spaces

    def __hash__(self) -> int:
        return hash(tuple(self.spaces))

    def sample_tensor(self, generator: torch.Generator | None = None) -> torch.Tensor:
        """int32 [B, 2] = (choice index, action id), uniform over each environment's choices; stays on the device."""
        draw = torch.rand(self.counts.shape[0], device=self.counts.device, generator=generator)
        index = torch.minimum((draw * self.counts).to(torch.int64), self.counts.to(torch.int64) - 1)
        value = self.starts.gather(1, index.unsqueeze(1)).squeeze(1)
        return torch.stack([index.to(torch.int32), value.to(torch.int32)], dim=1)

    def sample_nested(self) -> List[List[int]]:
        return self.sample_tensor().tolist()

    def sample(self) -> List[List[int]]:
        return self.sample_nested()
