"""Wildfire state: same fields / dtypes / shapes as the reference (envs/wildfire/env/structures/state.py:11-33)."""
from dataclasses import dataclass

import torch

from free_range_zoo_b200.utils.state import State


@dataclass(eq=False)
class WildfireState(State):
    """
    fires:        int32 [B, H, W]  sign = lit (>0) / unlit or out (<0) / no fire possible (0); |value| = power needed
    intensity:    int32 [B, H, W]
    fuel:         int32 [B, H, W]
    agents:       int32 [A, 2]     (y, x), shared by every environment, static
    suppressants: float32 [B, A]
    capacity:     float32 [B, A]
    equipment:    int32 [B, A]
    """
    fires: torch.Tensor
    intensity: torch.Tensor
    fuel: torch.Tensor
    agents: torch.Tensor
    suppressants: torch.Tensor
    capacity: torch.Tensor
    equipment: torch.Tensor

    def __post_init__(self):
        super().__post_init__()
        self.metadata = {'shared': ('agents', )}

    def __getitem__(self, indices) -> 'WildfireState':
        return WildfireState(fires=self.fires[indices], intensity=self.intensity[indices], fuel=self.fuel[indices],
                             agents=self.agents, suppressants=self.suppressants[indices],
                             capacity=self.capacity[indices], equipment=self.equipment[indices])
