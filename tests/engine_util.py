"""Dump the engine's device state in the golden-fixture layout (tests/golden/gen_golden.py::*_outputs)."""
import numpy as np
import torch

PAD = -100


def cpu(t):
    return t.detach().cpu().numpy()


def padded_indices(mask: np.ndarray) -> np.ndarray:
    """bool [B, T] -> int32 [B, T]: positions of the set entries of each row, left-packed, padded with -100."""
    out = np.full(mask.shape, PAD, np.int32)
    for b in range(mask.shape[0]):
        idx = np.nonzero(mask[b])[0]
        out[b, :len(idx)] = idx
    return out


def wildfire_outputs(env) -> dict:
    raw = env.unwrapped
    A = len(raw.agents)
    HW = raw.max_y * raw.max_x
    s = raw.state()
    mask = cpu(raw.action_mask) != 0  # [B, A, HW] over env-local tasks
    counts = cpu(raw.environment_task_count)
    lit_tasks = np.arange(HW)[None, :] < counts[:, None]
    out = dict(
        fires=cpu(s.fires), intensity=cpu(s.intensity), fuel=cpu(s.fuel), suppressants=cpu(s.suppressants),
        capacity=cpu(s.capacity), equipment=cpu(s.equipment),
        rewards=np.stack([cpu(raw.rewards[a]) for a in raw.agents], axis=1),
        terminated=np.stack([cpu(raw.terminations[a]) for a in raw.agents], axis=1),
        truncated=np.stack([cpu(raw.truncations[a]) for a in raw.agents], axis=1),
        num_moves=cpu(raw.num_moves), num_burnouts=cpu(raw.num_burnouts),
        burnouts=cpu(raw.infos['burnouts']), putouts=cpu(raw.infos['putouts']),
        env_task_count=counts, agent_task_count=cpu(raw.agent_task_count).T,
        self_obs=np.stack([cpu(raw.observations[a]['self']) for a in raw.agents], axis=1),
        others_obs=np.stack([cpu(raw.observations[a]['others']) for a in raw.agents], axis=0),
        task_obs=cpu(raw.observations[raw.agents[0]]['tasks_padded']),
    )
    if raw.show_bad_actions:
        out['action_map'] = np.stack([padded_indices(lit_tasks) for _ in range(A)], axis=0)
        out['bad_map'] = np.stack([padded_indices(lit_tasks & ~mask[:, a]) for a in range(A)], axis=0)
    else:
        out['action_map'] = np.stack([padded_indices(mask[:, a]) for a in range(A)], axis=0)
    return out


def cyber_outputs(env) -> dict:
    raw = env.unwrapped
    agents = raw.agents
    s = raw.state()
    N = raw._n_nodes
    out = dict(
        network_state=cpu(s.network_state), location=cpu(s.location), presence=cpu(s.presence),
        rewards=np.stack([cpu(raw.rewards[a]) for a in agents], axis=1),
        terminated=np.stack([cpu(raw.terminations[a]) for a in agents], axis=1),
        truncated=np.stack([cpu(raw.truncations[a]) for a in agents], axis=1),
        num_moves=cpu(raw.num_moves), env_task_count=cpu(raw.environment_task_count),
        agent_task_count=cpu(raw.agent_task_count).T,
        attacker_self=np.stack([cpu(raw.observations[a]['self']) for a in agents if a.startswith('attacker')], axis=1),
        defender_self=np.stack([cpu(raw.observations[a]['self']) for a in agents if a.startswith('defender')], axis=1),
        task_obs=np.stack([cpu(raw.observations[a]['tasks']) for a in agents], axis=0),
        task_store=cpu(raw.task_store),
    )
    counts = cpu(raw.agent_task_count).T
    for i, a in enumerate(agents):
        out[f'others__{a}'] = cpu(raw.observations[a]['others'])
        out[f'action_map__{a}'] = padded_indices(np.arange(N)[None, :] < counts[:, i:i + 1])
    return out


def _widen(array: np.ndarray, axis: int, width: int) -> np.ndarray:
    """Pad ``array`` with -100 along ``axis`` up to ``width`` (golden tables are as wide as the whole schedule)."""
    if array.shape[axis] >= width:
        return array
    shape = list(array.shape)
    shape[axis] = width - array.shape[axis]
    return np.concatenate([array, np.full(shape, PAD, array.dtype)], axis=axis)


def rideshare_outputs(env, width: int) -> dict:
    raw = env.unwrapped
    agents = raw.agents
    s = raw.state()
    K = raw._capacity
    counts = cpu(raw.environment_task_count)
    valid = np.arange(K)[None, :] < counts[:, None]
    table = cpu(s.passenger_table).copy()
    table[:, :, 0] = np.arange(table.shape[0])[:, None]
    table[~valid] = PAD
    mask = cpu(raw.task_mask)  # [B, A, K]
    task_store = cpu(raw._task_obs)
    task_obs, action_map = [], []
    for i in range(len(agents)):
        per_agent = np.full_like(task_store, PAD)
        for b in range(task_store.shape[0]):
            rows = np.nonzero(mask[b, i])[0]
            per_agent[b, :len(rows)] = task_store[b, rows]
        task_obs.append(per_agent)
        action_map.append(padded_indices(mask[:, i]))
    return dict(
        agents=cpu(s.agents), passengers=_widen(table, 1, width), passenger_count=counts,
        rewards=np.stack([cpu(raw.rewards[a]) for a in agents], axis=1),
        terminated=np.stack([cpu(raw.terminations[a]) for a in agents], axis=1),
        truncated=np.stack([cpu(raw.truncations[a]) for a in agents], axis=1),
        num_moves=cpu(raw.num_moves), env_task_count=counts, agent_task_count=cpu(raw.agent_task_count).T,
        self_obs=np.stack([cpu(raw.observations[a]['self']) for a in agents], axis=1),
        others_obs=np.stack([cpu(raw.observations[a]['others']) for a in agents], axis=0),
        task_store=_widen(task_store, 1, width), task_obs=_widen(np.stack(task_obs, axis=0), 2, width),
        action_map=_widen(np.stack(action_map, axis=0), 2, width),
    )
