"""Cybersecurity state: same fields / dtypes as the reference (envs/cybersecurity/env/structures/state.py:13-27)."""
from dataclasses import dataclass

import torch

from free_range_zoo_b200.utils.state import State


@dataclass(eq=False)
class CybersecurityState(State):
    """
    network_state: int32 [B, N]       exploitation state of each subnetwork (higher = worse)
    location:      int32 [B, D]       node of each defender, -1 = home node
    presence:      bool  [B, Att+D]   presence of every agent, attackers first
    """
    network_state: torch.Tensor
    location: torch.Tensor
    presence: torch.Tensor

    def __getitem__(self, indices) -> 'CybersecurityState':
        return CybersecurityState(network_state=self.network_state[indices], location=self.location[indices],
                                  presence=self.presence[indices])
