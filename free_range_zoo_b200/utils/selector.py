"""Round-robin agent selector with the pettingzoo ``agent_selector`` contract used by the AEC step cycle
(reference utils/env.py:155-156,220,242: ``reset() / next() / is_last() / is_first()``)."""
from __future__ import annotations

from typing import Sequence


class AgentSelector:

    def __init__(self, agent_order: Sequence[str]):
        self.reinit(agent_order)

    def reinit(self, agent_order: Sequence[str]) -> None:
        self.agent_order = list(agent_order)
        self._cursor = 0
        self.selected_agent = None

    def reset(self) -> str:
        self.reinit(self.agent_order)
        return self.next()

    def next(self) -> str:
        self.selected_agent = self.agent_order[self._cursor]
        self._cursor = (self._cursor + 1) % len(self.agent_order)
        return self.selected_agent

    def is_last(self) -> bool:
        return self.selected_agent == self.agent_order[-1]

    def is_first(self) -> bool:
        return self.selected_agent == self.agent_order[0]
