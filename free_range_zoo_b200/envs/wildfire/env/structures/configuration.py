"""Openness configuration structures of the wildfire domain.

Field names, order, defaults and derived properties follow the reference structures
(free_range_zoo/envs/wildfire/env/structures/configuration.py:15-390) so an existing configuration script keeps
working after switching the import root to ``free_range_zoo_b200``.
"""
from __future__ import annotations

import functools
import math

import numpy as np
from dataclasses import dataclass
from typing import List

import torch

from free_range_zoo_b200.utils.configuration import Configuration, in_unit_interval, require


@dataclass
class RewardConfiguration(Configuration):
    """Reward terms (reference configuration.py:15-56)."""
    fire_rewards: torch.Tensor  # f32 [H, W] reward for extinguishing the fire of a cell
    bad_attack_penalty: float
    burnout_penalty: float
    burnout_penalty_scaled: bool = False
    termination_reward: float = 0.0
    termination_kappa: float = 0.0
    localize_putouts: bool = False

    def validate(self) -> bool:
        require(self.fire_rewards.dim() == 2, 'fire_rewards should be a 2D tensor')
        require(not (self.burnout_penalty != 0 and self.burnout_penalty_scaled),
                'burnout_penalty and burnout_penalty_scaled are mutually exclusive')
        return True


@dataclass
class FireConfiguration(Configuration):
    """Task (fire) openness (reference configuration.py:59-161)."""
    fire_types: torch.Tensor  # i32 [H, W] suppressant power needed per cell, 0 = no fire possible
    num_fire_states: int
    lit: torch.Tensor  # bool [H, W]
    intensity_increase_probability: float
    intensity_decrease_probability: float
    extra_power_decrease_bonus: float
    burnout_probability: float
    base_spread_rate: float
    max_spread_rate: float
    random_ignition_probability: float
    cell_size: float
    wind_direction: float
    ignition_temp: torch.Tensor  # i32 [H, W]
    initial_fuel: int

    @functools.cached_property
    def burned_out(self) -> int:
        return self.num_fire_states - 1

    @functools.cached_property
    def almost_burned_out(self) -> int:
        return self.num_fire_states - 2

    @functools.cached_property
    def max_fire_type(self) -> int:
        return int(self.fire_types.max().item())

    @functools.cached_property
    def realistic_spread_rates(self) -> List[float]:
        """Wind-adjusted spread rate towards N, E, S, W (Eck et al. 2020; reference configuration.py:120-135)."""
        per_cell = self.base_spread_rate / self.cell_size
        headroom = 1 - self.base_spread_rate / self.max_spread_rate
        return [
            per_cell / (1 - np.cos(quarter * 0.5 * np.pi - self.wind_direction) * headroom)
            for quarter in range(4)
        ]

    def validate(self) -> bool:
        require(self.fire_types.dim() == 2, 'fires should be a 2D tensor')
        require(self.num_fire_states >= 4, 'num_fire_states should be greater than 4')
        require(self.lit.dim() == 2, 'lit should be a 2D tensor')
        for name in ('intensity_increase_probability', 'intensity_decrease_probability', 'burnout_probability',
                     'random_ignition_probability'):
            require(in_unit_interval(getattr(self, name)), f'{name} should be between 0 and 1')
        require(0.0 <= self.wind_direction <= 2 * math.pi, 'Wind direction must be between 0 and 2 * pi')
        require(self.lit.shape == self.fire_types.shape == self.ignition_temp.shape,
                'lit, fire_types, and ignition_temp must have the same shape')
        return True


@dataclass
class AgentConfiguration(Configuration):
    """Agent openness: suppressant, capacity and equipment dynamics (reference configuration.py:164-268)."""
    agents: torch.Tensor  # i32 [A, 2] (y, x), static
    fire_reduction_power: torch.Tensor  # [A]
    attack_range: torch.Tensor  # [A]
    suppressant_states: int
    initial_suppressant: int
    suppressant_decrease_probability: float
    suppressant_refill_probability: float
    initial_equipment_state: int
    equipment_states: torch.Tensor  # f32 [E, 3] modifiers (capacity, power, range)
    repair_probability: float
    degrade_probability: float
    critical_error_probability: float
    initial_capacity: int
    tank_switch_probability: float
    possible_capacities: torch.Tensor  # f32 [C]
    capacity_probabilities: torch.Tensor  # f32 [C]

    @functools.cached_property
    def num_agents(self) -> int:
        return self.agents.shape[0]

    @functools.cached_property
    def max_fire_reduction_power(self) -> float:
        return self.fire_reduction_power.max().item()

    @functools.cached_property
    def num_equipment_states(self) -> int:
        return self.equipment_states.shape[0]

    def validate(self) -> bool:
        require(self.agents.dim() == 2, 'agents should be a 2D tensor')
        require(self.fire_reduction_power.dim() == 1, 'fire_reduction_power should be a 1D tensor')
        require(self.attack_range.dim() == 1, 'attack_range should be a 1D tensor')
        require(self.agents.shape[0] == self.fire_reduction_power.shape[0],
                'agents, fire_reduction_power, and attack_range should have the same length')
        require(self.suppressant_states >= 2, 'suppressant_states should be greater than 2')
        require(self.initial_suppressant <= self.suppressant_states,
                'init_suppressant should be less than suppressant_states')
        for name in ('suppressant_decrease_probability', 'suppressant_refill_probability', 'repair_probability',
                     'degrade_probability', 'critical_error_probability', 'tank_switch_probability'):
            require(in_unit_interval(getattr(self, name)), f'{name} should be between 0 and 1')
        require(self.equipment_states.dim() == 2, 'equipment_states should be a 2D tensor')
        require(self.equipment_states.shape[1] == 3,
                'equipment_states should have 3 modifers: suppressant maximum, power, range')
        require(self.initial_equipment_state <= self.equipment_states.shape[0],
                'initial_equipment_state should be less than the number of equipment states')
        require(self.degrade_probability + self.critical_error_probability <= 1,
                'degrade_probability + critical_error_probability should be less than or equal to 1')
        require(self.possible_capacities.dim() == 1, 'possible_suppressant_maximums should be a 1D tensor')
        require(self.capacity_probabilities.dim() == 1, 'suppressant_maximum_probabilities should be a 1D tensor')
        require(self.possible_capacities.shape[0] == self.capacity_probabilities.shape[0],
                'possible_suppressant_maximums and suppressant_maximum_probabilities should have the same length')
        require(self.possible_capacities.min() >= 1, 'possible_suppressant_maximums should be greater than 1')
        require(self.capacity_probabilities.sum().item() == 1, 'suppressant_maximum_probabilities should sum to 1')
        return True


@dataclass
class StochasticConfiguration(Configuration):
    """Switches for every stochastic element (agent / task / frame openness; reference configuration.py:271-322)."""
    special_burnout_probability: bool
    suppressant_refill: bool
    suppressant_decrease: bool
    tank_switch: bool
    critical_error: bool
    degrade: bool
    repair: bool
    fire_increase: bool
    fire_decrease: bool
    fire_spread: bool
    realistic_fire_spread: bool
    random_fire_ignition: bool
    fire_fuel: bool

    def validate(self) -> bool:
        require(self.fire_spread or not self.realistic_fire_spread, 'Cannot use realistic fire spread without fire spread')
        require(self.degrade or not self.critical_error, 'Cannot have critical errors without equipment degradation')
        return True


@dataclass
class WildfireConfiguration(Configuration):
    """Top-level wildfire configuration (reference configuration.py:325-390)."""
    grid_width: int
    grid_height: int
    fire_config: FireConfiguration
    agent_config: AgentConfiguration
    reward_config: RewardConfiguration
    stochastic_config: StochasticConfiguration

    @functools.cached_property
    def fire_spread_weights(self) -> torch.Tensor:
        """3x3 cross-correlation filter [[0,N,0],[W,0,E],[0,S,0]] as f32 [1,1,3,3] (reference :347-363)."""
        weights = torch.zeros((1, 1, 3, 3), dtype=torch.float32)
        if not self.stochastic_config.fire_spread:
            return weights
        if self.stochastic_config.realistic_fire_spread:
            north, east, south, west = self.fire_config.realistic_spread_rates
        else:
            north = east = south = west = self.fire_config.base_spread_rate
        weights[0, 0, 0, 1], weights[0, 0, 1, 2], weights[0, 0, 2, 1], weights[0, 0, 1, 0] = north, east, south, west
        return weights

    @functools.cached_property
    def fire_random_spread_weight(self) -> float:
        """Bias added to every unlit cell's ignition probability (reference :365-371)."""
        return self.fire_config.random_ignition_probability if self.stochastic_config.random_fire_ignition else 0.0

    def validate(self) -> bool:
        super().validate()
        require(self.grid_width >= 1, 'grid_width should be greater than 0')
        require(self.grid_height >= 1, 'grid_height should be greater than 0')
        require(self.fire_config.lit.shape == self.reward_config.fire_rewards.shape,
                'lit and fire_rewards should have the same shape')
        return True
