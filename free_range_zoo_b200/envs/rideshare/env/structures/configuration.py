"""Configuration structures of the rideshare domain.

Same field names and meaning as the reference (free_range_zoo/envs/rideshare/env/structures/configuration.py:14-174).
The reference's configuration also acts as a factory for its four ``nn.Module`` transitions (:127-158); here the
whole step is one fused kernel, so the configuration only carries data.
"""
from __future__ import annotations

import functools
from dataclasses import dataclass

import torch

from free_range_zoo_b200.utils.configuration import Configuration, require


@dataclass
class RewardConfiguration(Configuration):
    """Reward terms (reference configuration.py:14-60). ``pick_cost`` and ``use_pooling_rewards`` are accepted for
    compatibility but, exactly as in the reference step (rideshare.py:309-363), never enter the reward."""
    pick_cost: float
    move_cost: float
    drop_cost: float
    noop_cost: float
    accept_cost: float
    pool_limit_cost: float
    use_pooling_rewards: bool
    use_variable_move_cost: bool
    use_waiting_costs: bool
    wait_limit: torch.Tensor  # i32 [3] limits for unaccepted / accepted / riding passengers
    long_wait_time: int
    general_wait_cost: float
    long_wait_cost: float

    def validate(self) -> bool:
        require(len(self.wait_limit) == 3, 'Wait limit should have three elements.')
        require(bool(self.wait_limit.min() > 0), 'Wait limit elements should all be greater than 0.')
        require(self.long_wait_time > 0, 'Long wait time should be greater than 0.')
        return True


@dataclass
class PassengerConfiguration(Configuration):
    """Task openness: the passenger entry schedule (reference configuration.py:63-80).

    ``schedule`` is i32 [tasks, 7] = (timestep, batch | -1 wildcard, y, x, y_dest, x_dest, fare).
    """
    schedule: torch.Tensor

    def validate(self) -> bool:
        require(self.schedule.dim() == 2, 'Schedule should be a 2D tensor')
        require(self.schedule.shape[-1] == 7, 'Schedule should have 7 elements in the last dimesion.')
        return True


@dataclass
class AgentConfiguration(Configuration):
    """Driver settings (reference configuration.py:83-110)."""
    start_positions: torch.Tensor  # i32 [A, 2]
    pool_limit: int
    use_diagonal_travel: bool
    use_fast_travel: bool

    @functools.cached_property
    def num_agents(self) -> int:
        return self.start_positions.shape[0]

    def validate(self) -> bool:
        require(self.pool_limit > 0, 'Pool limit must be greater than 0')
        return True


@dataclass
class RideshareConfiguration(Configuration):
    """Top-level rideshare configuration (reference configuration.py:113-174)."""
    grid_height: int
    grid_width: int
    agent_config: AgentConfiguration
    passenger_config: PassengerConfiguration
    reward_config: RewardConfiguration

    @functools.cached_property
    def max_fare(self) -> int:
        return int(self.passenger_config.schedule[:, 6].max().item())

    def validate(self) -> bool:
        super().validate()
        require(self.grid_width >= 1, 'grid_width should be greater than 0')
        require(self.grid_height >= 1, 'grid_height should be greater than 0')
        return True
