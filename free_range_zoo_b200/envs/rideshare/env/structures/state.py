"""Rideshare state (reference envs/rideshare/env/structures/state.py:11-22).

The reference stores ``passengers`` as ONE flat table ``[N_total, 11]`` sorted by environment.  The engine keeps a
fixed-capacity table per environment, ``passenger_table [B, K, 11]`` with ``passenger_count [B]`` valid rows each;
``passengers`` rebuilds the reference's flat view on demand (one boolean compaction).
"""
from dataclasses import dataclass

import torch

from free_range_zoo_b200.utils.state import State


@dataclass(eq=False)
class RideshareState(State):
    """
    agents:          int32 [B, A, 2] (y, x)
    passenger_table: int32 [B, K, 11] (batch, y, x, dest_y, dest_x, fare, state, association, entered_step,
                     accepted_step, picked_step); rows >= passenger_count[b] are undefined
    passenger_count: int32 [B]
    """
    agents: torch.Tensor
    passenger_table: torch.Tensor
    passenger_count: torch.Tensor

    @property
    def passengers(self) -> torch.Tensor:
        """The reference's flat ``[N_total, 11]`` table (rows of env 0, then env 1, ...)."""
        K = self.passenger_table.shape[1]
        keep = torch.arange(K, device=self.passenger_table.device).unsqueeze(0) < self.passenger_count.unsqueeze(1)
        rows = self.passenger_table[keep].clone()
        rows[:, 0] = keep.nonzero(as_tuple=False)[:, 0].to(rows.dtype)
        return rows

    def __getitem__(self, indices) -> 'RideshareState':
        return RideshareState(agents=self.agents[indices], passenger_table=self.passenger_table[indices],
                              passenger_count=self.passenger_count[indices])
