"""Configuration structures of the cybersecurity domain.

Same field names and derived properties as the reference
(free_range_zoo/envs/cybersecurity/env/structures/configuration.py:16-272); the transition factories of the
reference (:32-50) are dropped because the fused step kernel replaces the ``nn.Module`` transitions.
"""
from __future__ import annotations

import functools
from dataclasses import dataclass
from typing import Tuple

import torch

from free_range_zoo_b200.utils.configuration import Configuration, require


def _probabilities(tensor: torch.Tensor) -> bool:
    return bool(tensor.min() >= 0) and bool(tensor.max() <= 1)


@dataclass
class AttackerConfiguration(Configuration):
    """Attacker presence openness (reference configuration.py:100-140)."""
    initial_presence: torch.Tensor  # bool [Att]
    threat: torch.Tensor  # f32 [Att]
    persist_probs: torch.Tensor  # f32 [Att]
    return_probs: torch.Tensor  # f32 [Att]

    @functools.cached_property
    def num_attackers(self) -> int:
        return self.threat.size(0)

    @functools.cached_property
    def highest_threat(self) -> float:
        return self.threat.max().item()

    def validate(self) -> bool:
        require(_probabilities(self.persist_probs), 'Persist probabilities must be between 0 and 1.')
        require(_probabilities(self.return_probs), 'Return probabilities must be between 0 and 1.')
        require(self.threat.size(0) == self.persist_probs.size(0) == self.return_probs.size(0),
                'The size of threats must match the size of persist and return probabilities.')
        require(self.threat.size(0) == self.initial_presence.size(0),
                'The size of threats must match the size of initial presence values.')
        return True


@dataclass
class DefenderConfiguration(Configuration):
    """Defender presence openness and start locations (reference configuration.py:143-185)."""
    initial_location: torch.Tensor  # i32 [D], -1 = home node
    initial_presence: torch.Tensor  # bool [D]
    mitigation: torch.Tensor  # f32 [D]
    persist_probs: torch.Tensor  # f32 [D]
    return_probs: torch.Tensor  # f32 [D]

    @functools.cached_property
    def num_defenders(self) -> int:
        return self.mitigation.size(0)

    @functools.cached_property
    def highest_mitigation(self) -> float:
        return self.mitigation.max().item()

    def validate(self) -> bool:
        require(_probabilities(self.persist_probs), 'Persist probabilities must be between 0 and 1.')
        require(_probabilities(self.return_probs), 'Return probabilities must be between 0 and 1.')
        require(self.mitigation.size(0) == self.persist_probs.size(0) == self.return_probs.size(0),
                'The size of mitigations must match the size of persist and return probabilities.')
        require(self.mitigation.size(0) == self.initial_location.size(0) == self.initial_presence.size(0),
                'The size of mitigations must match the size of initial location and presence values.')
        return True


@dataclass
class NetworkConfiguration(Configuration):
    """Subnetwork graph and state ladder (reference configuration.py:188-237); home node is -1."""
    patched_states: int
    vulnerable_states: int
    exploited_states: int
    temperature: float
    initial_state: torch.Tensor  # i32 [N]
    adj_matrix: torch.Tensor  # bool [N, N]

    @functools.cached_property
    def criticality(self) -> torch.Tensor:
        """Out-degree of each node (reference :215-217)."""
        return self.adj_matrix.sum(dim=1)

    @functools.cached_property
    def num_nodes(self) -> int:
        return self.adj_matrix.size(0)

    @functools.cached_property
    def num_states(self) -> int:
        return self.patched_states + self.vulnerable_states + self.exploited_states

    def validate(self) -> bool:
        require(self.initial_state.size(0) == self.adj_matrix.size(0),
                'The size of initial state must match the number of nodes.')
        require(self.adj_matrix.size(0) == self.adj_matrix.size(1), 'The adjacency matrix must be square.')
        return True


@dataclass
class StochasticConfiguration(Configuration):
    """Whether subnetwork states move stochastically (reference configuration.py:240-253)."""
    network_state: bool


@dataclass
class RewardConfiguration(Configuration):
    """Reward terms (reference configuration.py:256-272)."""
    bad_action_penalty: float
    patch_reward: float
    network_state_rewards: torch.Tensor  # f32 [num_states]


@dataclass
class CybersecurityConfiguration(Configuration):
    """Top-level cybersecurity configuration (reference configuration.py:16-97)."""
    attacker_config: AttackerConfiguration
    defender_config: DefenderConfiguration
    network_config: NetworkConfiguration
    reward_config: RewardConfiguration
    stochastic_config: StochasticConfiguration

    @functools.cached_property
    def attacker_observation_bounds(self) -> Tuple[float, int]:
        return (self.attacker_config.highest_threat, 1)

    @functools.cached_property
    def defender_observation_bounds(self) -> Tuple[float, int, int]:
        return (self.defender_config.highest_mitigation, 1, self.network_config.num_nodes - 1)

    @functools.cached_property
    def network_observation_bounds(self) -> Tuple[int]:
        return (self.network_config.num_states, )

    @functools.cached_property
    def num_agents(self) -> int:
        return self.attacker_config.num_attackers + self.defender_config.num_defenders

    @functools.cached_property
    def persist_probs(self) -> torch.Tensor:
        return torch.cat([self.attacker_config.persist_probs, self.defender_config.persist_probs])

    @functools.cached_property
    def return_probs(self) -> torch.Tensor:
        return torch.cat([self.attacker_config.return_probs, self.defender_config.return_probs])

    @functools.cached_property
    def initial_presence(self) -> torch.Tensor:
        return torch.cat([self.attacker_config.initial_presence, self.defender_config.initial_presence])

    def validate(self) -> bool:
        super().validate()
        require(self.reward_config.network_state_rewards.size(0) == self.network_config.num_states,
                'The number of network state rewards must match the number of network states.')
        return True
