"""Drives the UNMODIFIED reference environments (imported through ``oracle/ref_shim.py``) for timing -- TEST / BASELINE
INFRASTRUCTURE used by ``bench.py --impl reference`` and the ``cpu_baseline`` leg; never imported by the product.

The rollout is the one SURVEY.md section 8(d) prescribes for the reference CPU path: ``device='cpu'``,
``single_seeding=True``, ``buffer_size=0``, logging off; uniform random legal actions are generated in numpy from the
reference's own task counts / tables (its ``action_space(agent).sample_nested()`` needs ``free_range_rust``, which is not
installable here) and are EXCLUDED from the timing; only ``env.step(actions)`` -- transitions, rewards, observations,
action-mapping refresh (free_range_zoo/utils/env.py:203-242) -- is timed.
"""
from __future__ import annotations

import importlib
import time

import numpy as np


def make_reference_env(domain: str, preset: str, preset_kwargs: dict, env_kwargs: dict, parallel_envs: int, seed: int):
    import torch

    from oracle import ref_shim
    ref_shim.install()
    from free_range_zoo_b200 import presets
    module = importlib.import_module(f'free_range_zoo.envs.{domain}_v0')
    structures = importlib.import_module(f'free_range_zoo.envs.{domain}.env.structures.configuration')
    configuration = getattr(presets, preset)(structures, **preset_kwargs)
    env = module.parallel_env(parallel_envs=parallel_envs, max_steps=1 << 30, configuration=configuration,
                              device=torch.device('cpu'), single_seeding=True, buffer_size=0, log_directory=None,
                              **env_kwargs)
    env.reset(seed=seed)
    return env, ref_shim.raw(env)


def legal_actions(domain: str, raw, rng) -> np.ndarray:
    """int32 [B, A, 2]: uniform over each agent's legal actions incl. the no-op (SURVEY.md 8d "Actions")."""
    B, agents = raw.parallel_envs, list(raw.agents)
    acts = np.zeros((B, len(agents), 2), dtype=np.int32)
    if domain == 'wildfire':
        for a in range(len(agents)):
            n = (raw.environment_task_count if raw.show_bad_actions else raw.agent_task_count[a]).numpy().astype(np.int64)
            k = np.minimum((rng.random(B) * (n + 1)).astype(np.int64), n)
            acts[:, a, 0], acts[:, a, 1] = k, np.where(k == n, -1, 0)
    elif domain == 'cybersecurity':
        n_att = raw.attacker_config.num_attackers
        counts, env_counts = raw.agent_task_count.numpy(), raw.environment_task_count.numpy()
        location = raw._state.location.numpy()
        for a, name in enumerate(agents):
            n = (env_counts if raw.show_bad_actions else counts[a]).astype(np.int64)
            choices = n + 1
            can_patch = np.zeros(B, bool)
            if not name.startswith('attacker'):
                can_patch = (n > 0) & (raw.show_bad_actions | (location[:, a - n_att] != -1))
                choices = choices + np.where(n > 0, 1 + can_patch, 0)
            k = np.minimum((rng.random(B) * choices).astype(np.int64), choices - 1)
            acts[:, a, 0] = k
            acts[:, a, 1] = np.where(k < n, 0, np.where(k == n, -1, np.where((k == n + 1) & can_patch, -2, -3)))
    else:  # rideshare: an agent's tasks are the unaccepted passengers and its own; the action id is the passenger state
        table = raw._state.passengers.numpy()
        for a in range(len(agents)):
            mine = (table[:, 6] == 0) | (table[:, 7] == a)
            rows = table[mine]
            order = np.argsort(rows[:, 0], kind='stable')
            rows = rows[order]
            starts = np.searchsorted(rows[:, 0], np.arange(B), side='left')
            n = np.searchsorted(rows[:, 0], np.arange(B), side='right') - starts
            k = np.minimum((rng.random(B) * (n + 1)).astype(np.int64), n)
            task = k < n
            acts[:, a, 0] = k
            acts[:, a, 1] = -1
            acts[task, a, 1] = rows[np.minimum(starts + k, len(rows) - 1)[task], 6] if len(rows) else -1
    return acts


def timed_rollout(domain: str, preset: str, preset_kwargs: dict, env_kwargs: dict, parallel_envs: int, steps: int,
                  seed: int, warmup: int = 1):
    """(executed env-steps, seconds inside ``env.step``) of one reference rollout; counts only steps that were actually
    executed (the reference's step returns early once every environment is done, utils/env.py:212)."""
    import torch
    env, raw = make_reference_env(domain, preset, preset_kwargs, env_kwargs, parallel_envs, seed)
    rng = np.random.default_rng(seed)
    seconds, executed = 0.0, 0
    for t in range(warmup + steps):
        if bool(torch.all(raw.finished)):
            break
        acts = legal_actions(domain, raw, rng)
        actions = {agent: torch.from_numpy(acts[:, i].copy()) for i, agent in enumerate(raw.agents)}
        start = time.perf_counter()
        env.step(actions)
        elapsed = time.perf_counter() - start
        if t >= warmup:
            seconds += elapsed
            executed += 1
    return executed * parallel_envs, seconds
