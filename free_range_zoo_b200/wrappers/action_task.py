"""``action_mapping_wrapper_v0``: every observation is returned together with the agent's action -> task mapping.

Same contract as the reference wrapper (free_range_zoo/wrappers/action_task.py:10-59, applied through supersuit's
``shared_wrapper``): wherever the environment hands out an agent's observation (``reset``, ``step``, ``observe``,
``last``) the caller receives the tuple ``(observation, {'agent_action_mapping': mapping})`` where ``mapping`` is the
jagged ``[B, n_tasks(agent)]`` tensor of environment-local task indices.  The baseline policies index
``observation['tasks']`` through it (e.g. envs/wildfire/baselines/strongest.py:28-62).  Here the mapping is built
lazily on the device from the step kernel's task masks (one cumulative sum, no host round trip).
"""
from __future__ import annotations

from typing import Any, Dict, Tuple


class ActionTaskMappingWrapper:
    """Wraps a Parallel or AEC environment of this package; unknown attributes are forwarded."""

    def __init__(self, env):
        self.env = env

    def __getattr__(self, name: str):
        if name == 'env':
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def _with_mapping(self, agent: str, observation) -> Tuple[Any, Dict[str, Any]]:
        return observation, {'agent_action_mapping': self.env.unwrapped.agent_action_mapping[agent]}

    def _map(self, observations: Dict[str, Any]) -> Dict[str, Any]:
        return {agent: self._with_mapping(agent, observation) for agent, observation in observations.items()}

    # -- Parallel API
    def reset(self, *args, **kwargs):
        result = self.env.reset(*args, **kwargs)
        if result is None:  # AEC reset returns nothing
            return None
        observations, infos = result
        return self._map(observations), infos

    def step(self, actions):
        result = self.env.step(actions)
        if result is None:  # AEC step returns nothing
            return None
        observations, rewards, terminations, truncations, infos = result
        return self._map(observations), rewards, terminations, truncations, infos

    # -- AEC API
    def observe(self, agent: str = None):
        if agent is None:  # Parallel observe(): every agent
            return self._map(self.env.observe())
        return self._with_mapping(agent, self.env.observe(agent))

    def last(self, observe: bool = True):
        observation, reward, termination, truncation, info = self.env.last(observe)
        if observe:
            observation = self._with_mapping(self.env.agent_selection, observation)
        return observation, reward, termination, truncation, info


def action_mapping_wrapper_v0(env, **kwargs) -> ActionTaskMappingWrapper:
    """Apply the action -> task mapping wrapper (reference wrappers/action_task.py:49-59)."""
    return ActionTaskMappingWrapper(env)
