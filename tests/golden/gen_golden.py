"""Generate the golden trajectories under tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/gen_golden.py            # regenerates every fixture
    python tests/golden/gen_golden.py NAME ...   # only the named fixtures

For each fixture the reference environment (imported through ``ref_shim``) is rolled out on CPU with random legal
actions, with ``env.generator.generate`` replaced by a recorder that serves seeded uniforms of the shape the
reference asks for (``(events, B, *shape)``, free_range_zoo/utils/random_generator.py:87-115).  Recorded per step:
the actions, the uniforms, and every observable output of the step (state, rewards, terminations, truncations,
counters, task counts, observations, action mappings).  Index 0 of every output array is the post-``reset`` value.

The fixtures are consumed by tests/test_oracle_golden.py (oracle vs reference, CPU) and tests/test_*_gpu.py
(CUDA engine vs reference, GPU).  Jagged tensors are stored padded with -100 (the reference's own padding value,
e.g. wildfire.py:439).
"""
import json
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import ref_shim  # noqa: E402

ref_shim.install()

from free_range_zoo.envs import wildfire_v0, rideshare_v0, cybersecurity_v0  # noqa: E402
from free_range_zoo.envs.wildfire.env.structures import configuration as ref_wf_conf  # noqa: E402
from free_range_zoo.envs.rideshare.env.structures import configuration as ref_rs_conf  # noqa: E402
from free_range_zoo.envs.cybersecurity.env.structures import configuration as ref_cy_conf  # noqa: E402

from free_range_zoo_b200 import presets  # noqa: E402

PAD = -100
ONLY = set(sys.argv[1:])  # fixture names given on the command line (empty = all)


def padded(nested, width, dtype=np.int32):
    """Jagged [B, j, ...] -> dense [B, width, ...] padded with -100."""
    dense = nested.to_padded_tensor(PAD).numpy()
    out = np.full((dense.shape[0], width) + dense.shape[2:], PAD, dtype=dtype)
    out[:, :dense.shape[1]] = dense
    return out


class UniformRecorder:
    """Stand-in for RandomGenerator.generate: seeded, recorded, shape (events, B, *shape)."""

    def __init__(self, seed):
        self.gen = torch.Generator().manual_seed(seed)
        self.log = {}

    def __call__(self, parallel_envs, events, shape, key=None):
        u = torch.rand((events, parallel_envs, *shape), generator=self.gen)
        self.log.setdefault(key, []).append(u.numpy().copy())
        return u


class Trajectory:

    def __init__(self):
        self.cols = {}

    def add(self, **arrays):
        for k, v in arrays.items():
            self.cols.setdefault(k, []).append(np.asarray(v))

    def stacked(self):
        return {k: np.stack(v) for k, v in self.cols.items()}


def save(name, meta, traj, recorder, extra=None):
    data = traj.stacked()
    for key, seq in recorder.log.items():
        data[f'u_{key}'] = np.stack(seq) if seq else np.zeros((0, ))
    data.update(extra or {})
    data['meta'] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(HERE, f'{name}.npz')
    np.savez_compressed(path, **data)
    print(f'{name}: {os.path.getsize(path) / 1024:.1f} KiB, steps={meta["steps"]}')


# ------------------------------------------------------------------------------------------------ wildfire


def wildfire_outputs(raw, rewards=None):
    s = raw._state
    B, H, W = s.fires.shape
    A = len(raw.agents)
    agents = raw.agents
    out = dict(
        fires=s.fires.numpy().copy(),
        intensity=s.intensity.numpy().copy(),
        fuel=s.fuel.numpy().copy(),
        suppressants=s.suppressants.numpy().copy(),
        capacity=s.capacity.numpy().copy(),
        equipment=s.equipment.numpy().copy(),
        rewards=np.stack([(rewards[a] if rewards else torch.zeros(B)).numpy() for a in agents], axis=1),
        terminated=np.stack([raw.terminations[a].numpy() for a in agents], axis=1),
        truncated=np.stack([raw.truncations[a].numpy() for a in agents], axis=1),
        num_moves=raw.num_moves.numpy().copy(),
        num_burnouts=raw.num_burnouts.numpy().copy(),
        burnouts=(raw.infos['burnouts'].numpy() if 'burnouts' in raw.infos else np.zeros(B, np.int64)).astype(np.int32),
        putouts=(raw.infos['putouts'].numpy() if 'putouts' in raw.infos else np.zeros(B, np.int64)).astype(np.int32),
        env_task_count=raw.environment_task_count.numpy().astype(np.int32),
        agent_task_count=raw.agent_task_count.numpy().T.astype(np.int32),
        self_obs=np.stack([raw.observations[a]['self'].numpy() for a in agents], axis=1),
        others_obs=np.stack([raw.observations[a]['others'].numpy() for a in agents], axis=0),
        task_obs=padded(raw.task_store, H * W),
        action_map=np.stack([padded(raw.agent_action_mapping[a], H * W) for a in agents], axis=0),
        observation_map=padded(raw.agent_observation_mapping[agents[0]], H * W),
    )
    if raw.show_bad_actions:
        out['bad_map'] = np.stack([padded(raw.agent_bad_actions[a], H * W) for a in agents], axis=0)
    return out


def wildfire_actions(raw, gen):
    """Uniform over each agent's legal actions incl. noop: [k, 0] for k < n, [n, -1] otherwise."""
    B, A = raw.parallel_envs, len(raw.agents)
    acts = np.zeros((B, A, 2), dtype=np.int32)
    for a, agent in enumerate(raw.agents):
        n = (raw.environment_task_count if raw.show_bad_actions else raw.agent_task_count[a]).numpy().astype(np.int64)
        k = (torch.rand(B, generator=gen).numpy() * (n + 1)).astype(np.int64)
        k = np.minimum(k, n)
        acts[:, a, 0] = k
        acts[:, a, 1] = np.where(k == n, -1, 0)
    return acts


def gen_wildfire(name, preset, B, steps, seed, preset_kwargs=None, **env_kwargs):
    if ONLY and name not in ONLY:
        return
    torch.manual_seed(seed)
    preset_kwargs = preset_kwargs or {}
    config = preset(ref_wf_conf, **preset_kwargs)
    env = wildfire_v0.parallel_env(parallel_envs=B, max_steps=steps, configuration=config, device=torch.device('cpu'),
                                   single_seeding=True, **env_kwargs)
    env.reset(seed=seed)
    raw = ref_shim.raw(env)
    recorder = UniformRecorder(seed + 1)
    raw.generator.generate = recorder
    gen = torch.Generator().manual_seed(seed + 2)

    traj = Trajectory()
    traj.add(**wildfire_outputs(raw))
    actions = []
    executed = 0
    for _ in range(steps + 2):
        if torch.all(raw.finished):
            break
        acts = wildfire_actions(raw, gen)
        actions.append(acts)
        _, rewards, _, _, _ = env.step({a: torch.from_numpy(acts[:, i].copy()) for i, a in enumerate(raw.agents)})
        traj.add(**wildfire_outputs(raw, rewards))
        executed += 1

    # The 16 possible lit-neighbour patterns pushed through the reference's own conv2d: pins the fp32 summation.
    patterns = torch.zeros((16, 1, 3, 3))
    for p in range(16):
        patterns[p, 0, 0, 1] = p & 1  # north
        patterns[p, 0, 1, 0] = (p >> 1) & 1  # west
        patterns[p, 0, 1, 2] = (p >> 2) & 1  # east
        patterns[p, 0, 2, 1] = (p >> 3) & 1  # south
    with torch.no_grad():
        lut = raw.fire_spread_transition.fire_spread_filter(patterns)[:, 0, 1, 1].numpy().copy()

    meta = dict(domain='wildfire', preset=preset.__name__, preset_kwargs=preset_kwargs, B=B, steps=executed,
                max_steps=steps, seed=seed, env_kwargs=env_kwargs)
    save(name, meta, traj, recorder, extra=dict(
        actions=np.stack(actions),
        spread_weights=raw.fire_spread_weights.numpy().copy(),
        spread_lut=lut,
    ))


# ------------------------------------------------------------------------------------------------ rideshare


def rideshare_tables(raw, K):
    p = raw._state.passengers.numpy()
    B = raw.parallel_envs
    table = np.full((B, K, 11), PAD, dtype=np.int32)
    counts = np.bincount(p[:, 0], minlength=B).astype(np.int32)
    offset = 0
    for b in range(B):
        table[b, :counts[b]] = p[offset:offset + counts[b]]
        offset += counts[b]
    return table, counts


def rideshare_outputs(raw, K, rewards=None):
    B = raw.parallel_envs
    agents = raw.agents
    table, counts = rideshare_tables(raw, K)
    out = dict(
        agents=raw._state.agents.numpy().copy(),
        passengers=table,
        passenger_count=counts,
        rewards=np.stack([(rewards[a] if rewards else torch.zeros(B)).numpy() for a in agents], axis=1),
        terminated=np.stack([raw.terminations[a].numpy() for a in agents], axis=1),
        truncated=np.stack([raw.truncations[a].numpy() for a in agents], axis=1),
        num_moves=raw.num_moves.numpy().copy(),
        env_task_count=raw.environment_task_count.numpy().astype(np.int32),
        agent_task_count=raw.agent_task_count.numpy().T.astype(np.int32),
        self_obs=np.stack([raw.observations[a]['self'].numpy() for a in agents], axis=1),
        others_obs=np.stack([raw.observations[a]['others'].numpy() for a in agents], axis=0),
        task_store=padded(raw.task_store, K),
        task_obs=np.stack([padded(raw.observations[a]['tasks'], K) for a in agents], axis=0),
        action_map=np.stack([padded(raw.agent_action_mapping[a], K) for a in agents], axis=0),
    )
    return out


def rideshare_actions(raw, gen, wild):
    """Uniform over [tasks..., noop]; the action id of a task is the passenger's state (spaces/actions.py:10-50).
    With ``wild`` > 0 that fraction of task actions gets a random id in {0,1,2} instead (legal to the step code)."""
    B, A = raw.parallel_envs, len(raw.agents)
    p = raw._state.passengers.numpy()
    acts = np.zeros((B, A, 2), dtype=np.int32)
    for a in range(A):
        mine = (p[:, 6] == 0) | (p[:, 7] == a)
        for b in range(B):
            states = p[mine & (p[:, 0] == b), 6]
            n = len(states)
            k = min(int(torch.rand(1, generator=gen).item() * (n + 1)), n)
            if k == n:
                acts[b, a] = (n, -1)
            else:
                ident = int(states[k])
                if wild and torch.rand(1, generator=gen).item() < wild:
                    ident = int(torch.randint(0, 3, (1, ), generator=gen).item())
                acts[b, a] = (k, ident)
    return acts


def gen_rideshare(name, preset, B, steps, seed, wild=0.0, preset_kwargs=None, **env_kwargs):
    if ONLY and name not in ONLY:
        return
    config = preset(ref_rs_conf, **(preset_kwargs or {}))
    K = int(config.passenger_config.schedule.shape[0])
    env = rideshare_v0.parallel_env(parallel_envs=B, max_steps=steps, configuration=config, device=torch.device('cpu'),
                                    single_seeding=True, **env_kwargs)
    env.reset(seed=seed)
    raw = ref_shim.raw(env)
    recorder = UniformRecorder(seed + 1)
    gen = torch.Generator().manual_seed(seed + 2)

    traj = Trajectory()
    traj.add(**rideshare_outputs(raw, K))
    actions = []
    executed = 0
    for _ in range(steps + 2):
        if torch.all(raw.finished):
            break
        acts = rideshare_actions(raw, gen, wild)
        actions.append(acts)
        _, rewards, _, _, _ = env.step({a: torch.from_numpy(acts[:, i].copy()) for i, a in enumerate(raw.agents)})
        traj.add(**rideshare_outputs(raw, K, rewards))
        executed += 1
    meta = dict(domain='rideshare', preset=preset.__name__, B=B, steps=executed, max_steps=steps, seed=seed,
                preset_kwargs=preset_kwargs or {}, wild=wild)
    save(name, meta, traj, recorder, extra=dict(actions=np.stack(actions)))


# ------------------------------------------------------------------------------------------------ cybersecurity


def cyber_outputs(raw, rewards=None):
    s = raw._state
    B = raw.parallel_envs
    agents = raw.agents
    att = [a for a in agents if a.startswith('attacker')]
    dfd = [a for a in agents if a.startswith('defender')]
    out = dict(
        network_state=s.network_state.numpy().copy(),
        location=s.location.numpy().copy(),
        presence=s.presence.numpy().copy(),
        rewards=np.stack([(rewards[a] if rewards else torch.zeros(B)).numpy() for a in agents], axis=1),
        terminated=np.stack([raw.terminations[a].numpy() for a in agents], axis=1),
        truncated=np.stack([raw.truncations[a].numpy() for a in agents], axis=1),
        num_moves=raw.num_moves.numpy().copy(),
        env_task_count=raw.environment_task_count.numpy().astype(np.int32),
        agent_task_count=raw.agent_task_count.numpy().T.astype(np.int32),
        attacker_self=np.stack([raw.observations[a]['self'].numpy() for a in att], axis=1),
        defender_self=np.stack([raw.observations[a]['self'].numpy() for a in dfd], axis=1),
        task_obs=np.stack([raw.observations[a]['tasks'].numpy() for a in agents], axis=0).astype(np.int32),
        task_store=raw.task_store.numpy().astype(np.int32),
    )
    for a in agents:
        out[f'others__{a}'] = raw.observations[a]['others'].numpy().astype(np.float32)
        out[f'action_map__{a}'] = padded(raw.agent_action_mapping[a], raw.network_config.num_nodes)
    return out


def cyber_actions(raw, gen):
    """Uniform over the legal choices of spaces/actions.py:11-99 (attack/move k, noop -1, patch -2, monitor -3)."""
    B = raw.parallel_envs
    n_att = raw.attacker_config.num_attackers
    acts = np.zeros((B, len(raw.agents), 2), dtype=np.int32)
    counts = raw.agent_task_count.numpy()
    env_counts = raw.environment_task_count.numpy()
    loc = raw._state.location.numpy()
    for i, agent in enumerate(raw.agents):
        for b in range(B):
            n = int(env_counts[b] if raw.show_bad_actions else counts[i, b])
            if agent.startswith('attacker') or n == 0:
                choices = [(k, 0) for k in range(n)] + [(n, -1)]
            else:
                choices = [(k, 0) for k in range(n)] + [(n, -1)]
                if raw.show_bad_actions or loc[b, i - n_att] != -1:
                    choices.append((len(choices), -2))
                choices.append((len(choices), -3))
            acts[b, i] = choices[int(torch.randint(0, len(choices), (1, ), generator=gen).item())]
    return acts


def gen_cyber(name, preset, B, steps, seed, preset_kwargs=None, **env_kwargs):
    if ONLY and name not in ONLY:
        return
    preset_kwargs = preset_kwargs or {}
    config = preset(ref_cy_conf, **preset_kwargs)
    env = cybersecurity_v0.parallel_env(parallel_envs=B, max_steps=steps, configuration=config,
                                        device=torch.device('cpu'), single_seeding=True, **env_kwargs)
    env.reset(seed=seed)
    raw = ref_shim.raw(env)
    recorder = UniformRecorder(seed + 1)
    raw.generator.generate = recorder
    gen = torch.Generator().manual_seed(seed + 2)

    traj = Trajectory()
    traj.add(**cyber_outputs(raw))
    actions = []
    executed = 0
    for _ in range(steps + 2):
        if torch.all(raw.finished):
            break
        acts = cyber_actions(raw, gen)
        actions.append(acts)
        _, rewards, _, _, _ = env.step({a: torch.from_numpy(acts[:, i].copy()) for i, a in enumerate(raw.agents)})
        traj.add(**cyber_outputs(raw, rewards))
        executed += 1
    meta = dict(domain='cybersecurity', preset=preset.__name__, preset_kwargs=preset_kwargs, B=B, steps=executed,
                max_steps=steps, seed=seed, env_kwargs=env_kwargs, agents=list(raw.agents))
    # tanh of every reachable danger score as computed by torch on CPU (pins the one transcendental on the path)
    save(name, meta, traj, recorder, extra=dict(actions=np.stack(actions)))


def main():
    gen_wildfire('wildfire_profile', presets.wildfire_profile, B=16, steps=15, seed=11)
    gen_wildfire('wildfire_c1', presets.wildfire_3x3, B=48, steps=40, seed=12)
    gen_wildfire('wildfire_c4', presets.wildfire_large, B=8, steps=30, seed=13)
    gen_wildfire('wildfire_quirks', presets.wildfire_quirks, B=24, steps=40, seed=14, show_bad_actions=True,
                 observe_other_power=True, observe_other_suppressant=True)
    gen_wildfire('wildfire_quirks_good', presets.wildfire_quirks, B=24, steps=30, seed=15, show_bad_actions=False,
                 observe_other_power=False, observe_other_suppressant=True)

    # the other kernel geometries: half-warp groups with the cells in registers (7x8), one warp per environment (12x12)
    gen_wildfire('wildfire_7x8', presets.wildfire_large, B=12, steps=25, seed=16,
                 preset_kwargs=dict(height=7, width=8, num_agents=5, seed=5))
    gen_wildfire('wildfire_12x12', presets.wildfire_large, B=6, steps=20, seed=17,
                 preset_kwargs=dict(height=12, width=12, num_agents=20, seed=3))
    # a 32-cell grid with 8 agents: the largest shape of the small-grid kernels (8 lanes x 4 cells / one thread per env)
    gen_wildfire('wildfire_4x8', presets.wildfire_large, B=10, steps=25, seed=18,
                 preset_kwargs=dict(height=4, width=8, num_agents=8, seed=33))

    # the other cells-per-lane classes of the small-grid kernels (1, 3 and 4 cells per lane of the 8-lane layout, with
    # and without spare lanes feeding the agents' random words), bad actions shown, one row of 32 cells
    gen_wildfire('wildfire_2x4', presets.wildfire_large, B=12, steps=20, seed=19,
                 preset_kwargs=dict(height=2, width=4, num_agents=3, seed=41))
    gen_wildfire('wildfire_4x5_bad', presets.wildfire_large, B=10, steps=25, seed=51, show_bad_actions=True,
                 preset_kwargs=dict(height=4, width=5, num_agents=4, seed=32))
    gen_wildfire('wildfire_5x6', presets.wildfire_large, B=10, steps=25, seed=52,
                 preset_kwargs=dict(height=5, width=6, num_agents=6, seed=21))
    gen_wildfire('wildfire_1x32', presets.wildfire_large, B=8, steps=20, seed=53,
                 preset_kwargs=dict(height=1, width=32, num_agents=5, seed=42))

    gen_rideshare('rideshare_profile', presets.rideshare_profile, B=8, steps=20, seed=21)
    gen_rideshare('rideshare_c2', presets.rideshare_c2, B=24, steps=100, seed=22)
    gen_rideshare('rideshare_quirks', presets.rideshare_quirks, B=16, steps=40, seed=23,
                  preset_kwargs=dict(parallel_envs=16))
    gen_rideshare('rideshare_wild', presets.rideshare_quirks, B=16, steps=40, seed=24, wild=0.35,
                  preset_kwargs=dict(parallel_envs=16, diagonal=False))
    gen_rideshare('rideshare_fast', presets.rideshare_quirks, B=16, steps=40, seed=25,
                  preset_kwargs=dict(parallel_envs=16, fast=True))

    # the wider kernel geometries: 16 / 32 lanes per environment
    gen_rideshare('rideshare_12drivers', presets.rideshare_synthetic, B=6, steps=25, seed=26,
                  preset_kwargs=dict(drivers=12, rows=40))
    gen_rideshare('rideshare_20drivers', presets.rideshare_synthetic, B=4, steps=25, seed=27,
                  preset_kwargs=dict(drivers=20, rows=64))

    # the eight-driver variants of the tiled kernel (one thread per environment): 32- and 64-bit row sets
    gen_rideshare('rideshare_7drivers', presets.rideshare_synthetic, B=8, steps=25, seed=28,
                  preset_kwargs=dict(drivers=7, rows=24))
    gen_rideshare('rideshare_8drivers', presets.rideshare_synthetic, B=6, steps=25, seed=29,
                  preset_kwargs=dict(drivers=8, rows=40))
    gen_rideshare('rideshare_6drivers', presets.rideshare_synthetic, B=8, steps=25, seed=30,
                  preset_kwargs=dict(drivers=6, rows=16))
    gen_rideshare('rideshare_2drivers', presets.rideshare_synthetic, B=6, steps=25, seed=38, wild=0.2,
                  preset_kwargs=dict(drivers=2, rows=64))

    gen_cyber('cyber_profile', presets.cyber_profile, B=8, steps=20, seed=31)
    gen_cyber('cyber_c3', presets.cyber_c3, B=32, steps=60, seed=32, show_bad_actions=False, partially_observable=True)
    gen_cyber('cyber_quirks', presets.cyber_quirks, B=24, steps=50, seed=33, show_bad_actions=True,
              partially_observable=True, observe_other_location=True, observe_other_presence=True,
              observe_other_power=False)
    gen_cyber('cyber_quirks_open', presets.cyber_quirks, B=16, steps=30, seed=34, show_bad_actions=False,
              partially_observable=False, observe_other_location=True, observe_other_presence=False,
              observe_other_power=True)
    # the other size classes of the tiled kernel: (8, 4, 4), (16, 8, 8), and a network beyond them (direct kernel)
    gen_cyber('cyber_6nodes', presets.cyber_synthetic, B=12, steps=25, seed=35,
              preset_kwargs=dict(nodes=6, attackers=3, defenders=3, seed=71), show_bad_actions=True)
    gen_cyber('cyber_10nodes', presets.cyber_synthetic, B=8, steps=25, seed=36,
              preset_kwargs=dict(nodes=10, attackers=5, defenders=4, seed=77), show_bad_actions=True)
    gen_cyber('cyber_20nodes', presets.cyber_synthetic, B=6, steps=20, seed=37,
              preset_kwargs=dict(nodes=20, attackers=10, defenders=9, seed=79), show_bad_actions=True)


if __name__ == '__main__':
    main()
