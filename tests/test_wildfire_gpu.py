"""GPU parity tests of the fused wildfire step (through the C ABI, via the public Parallel API).

1. against the trajectories recorded from the unmodified reference (tests/golden/wildfire_*.npz),
2. against the CPU oracle on larger seeded cases with injected uniforms,
3. size-independent properties at BASELINE.json's full size with in-kernel Philox randomness.
Integer state, masks and dones are compared bit-exactly; float rewards within 1e-5 relative (golden_util).
"""
import numpy as np
import pytest
import torch

from free_range_zoo_b200 import presets
from tests import golden_util as G
from tests.engine_util import cpu, wildfire_outputs

pytestmark = pytest.mark.gpu


def make_env(config, B, max_steps, **kwargs):
    from free_range_zoo_b200.envs import wildfire_v0
    return wildfire_v0.parallel_env(parallel_envs=B, max_steps=max_steps, configuration=config,
                                    device=torch.device('cuda'), **kwargs)


def tiled_kernel_serves(config) -> bool:
    """Grids the one-thread-per-environment kernel steps (csrc/frz_wildfire_tile.cuh); others ignore step_kernel."""
    return config.grid_height * config.grid_width <= 32 and config.agent_config.agents.shape[0] <= 8


# 'groups': lanes per environment (every grid); 'tiles': one thread per environment (small grids, forced at any batch size)
KERNELS = ['groups', 'tiles']


@pytest.mark.parametrize('kernel', KERNELS)
@pytest.mark.parametrize('name', G.fixtures('wildfire'))
def test_matches_reference_trajectory(name, kernel):
    meta, gold = G.load(name)
    config = getattr(presets, meta['preset'])(**meta.get('preset_kwargs', {}))
    if kernel == 'tiles' and not tiled_kernel_serves(config):
        pytest.skip('grid too large for the tiled kernel')
    env = make_env(config, meta['B'], meta['max_steps'], step_kernel=kernel, **meta['env_kwargs'])
    env.reset(seed=0)
    G.compare(wildfire_outputs(env), gold, 0, context=name)
    agents = env.agents
    for t in range(meta['steps']):
        env.unwrapped.inject_uniforms(torch.from_numpy(gold['u_field'][t]), torch.from_numpy(gold['u_agent'][t]))
        actions = torch.from_numpy(gold['actions'][t]).cuda()
        env.step({a: actions[:, i] for i, a in enumerate(agents)})
        G.compare(wildfire_outputs(env), gold, t + 1, context=name)
    env.unwrapped.check_errors()


@pytest.mark.parametrize('preset,B,steps,kwargs', [
    ('wildfire_large', 1024, 25, {}),
    ('wildfire_large', 333, 12, dict(show_bad_actions=True)),
    ('wildfire_quirks', 2048, 30, dict(show_bad_actions=True, observe_other_power=True)),
    ('wildfire_3x3', 4099, 40, {}),
    ('wildfire_profile', 1000, 15, {}),
    ('wildfire_large', 1, 10, {}),  # a single environment
    ('wildfire_3x3', 3, 12, {}),  # fewer environments than one warp holds
    # small grids (both kernels): all four cells-per-lane classes of the 8-lane layout, 8 agents, bad actions shown
    (('wildfire_large', dict(height=2, width=4, num_agents=3, seed=41)), 333, 20, {}),
    (('wildfire_large', dict(height=4, width=5, num_agents=4, seed=32)), 450, 20, dict(show_bad_actions=True)),
    (('wildfire_large', dict(height=5, width=6, num_agents=6, seed=21)), 300, 20, {}),
    (('wildfire_large', dict(height=4, width=8, num_agents=8, seed=33)), 257, 20, {}),
    (('wildfire_large', dict(height=1, width=32, num_agents=5, seed=42)), 200, 15, {}),  # one row of 32 cells
    # the kernel's other geometries: 2 and 8 cells per lane, rows of >= 32 cells (word-crossing neighbours), 20 agents
    (('wildfire_large', dict(height=7, width=8, num_agents=5, seed=5)), 700, 20, {}),
    (('wildfire_large', dict(height=12, width=16, num_agents=12, seed=6)), 300, 20, {}),
    (('wildfire_large', dict(height=4, width=40, num_agents=6, seed=7)), 300, 20, {}),
    (('wildfire_large', dict(height=3, width=33, num_agents=20, seed=8)), 300, 20, dict(show_bad_actions=True)),
    (('wildfire_large', dict(height=2, width=100, num_agents=7, seed=9)), 200, 15, {}),
    (('wildfire_large', dict(height=3, width=33, num_agents=6, seed=10)), 300, 20, {}),  # half-warp groups, rows >= 32 cells
    (('wildfire_large', dict(height=5, width=9, num_agents=16, seed=11)), 300, 20, {}),  # half-warp groups, 16 agents
    (('wildfire_large', dict(height=16, width=16, num_agents=32, seed=12)), 100, 12, {}),  # the engine's size limits
])
@pytest.mark.parametrize('kernel', KERNELS)
def test_matches_oracle_on_random_rollouts(preset, B, steps, kwargs, kernel):
    from oracle.wildfire import WildfireOracle
    config = getattr(presets, preset)() if isinstance(preset, str) else getattr(presets, preset[0])(**preset[1])
    if kernel == 'tiles' and not tiled_kernel_serves(config):
        pytest.skip('grid too large for the tiled kernel')
    oracle = WildfireOracle(config, B, steps, **kwargs)
    oracle.reset()
    env = make_env(config, B, steps, step_kernel=kernel, **kwargs)
    env.reset(seed=1)
    rng = np.random.default_rng(7)
    H, W, A = oracle.H, oracle.W, oracle.A
    keys = ('fires', 'intensity', 'fuel', 'suppressants', 'capacity', 'equipment', 'rewards', 'terminated', 'truncated',
            'num_moves', 'num_burnouts', 'burnouts', 'putouts', 'env_task_count', 'agent_task_count', 'self_obs',
            'task_obs', 'action_map', 'bad_map')
    for t in range(steps):
        counts = oracle.environment_task_count[:, None] if kwargs.get('show_bad_actions') else oracle.agent_task_count
        k = np.minimum((rng.random((B, A)) * (counts + 1)).astype(np.int64), counts)
        actions = np.stack([k, np.where(k == counts, -1, 0)], axis=2).astype(np.int32)
        u_field = rng.random((3, B, H, W), dtype=np.float32)
        u_agent = rng.random((5, B, A), dtype=np.float32)
        if not oracle.step(actions, u_field, u_agent):
            break
        env.unwrapped.inject_uniforms(torch.from_numpy(u_field), torch.from_numpy(u_agent))
        env.step(torch.from_numpy(actions).cuda())
        want = {key: value[None] for key, value in oracle.outputs().items()}
        G.compare({k_: v for k_, v in wildfire_outputs(env).items() if k_ in keys}, want, 0, context=f'{preset} t={t}')
    env.unwrapped.check_errors()


def rollout(env, steps, sampler_seed=5):
    raw = env.unwrapped
    for _ in range(steps):
        raw.sample_actions(sampler_seed)
        raw.step_all()
    torch.cuda.synchronize()


def test_philox_rollout_is_deterministic_and_shard_invariant():
    config = presets.wildfire_large()
    full = make_env(config, 512, 50)
    full.reset(seed=123)
    rollout(full, 20)
    again = make_env(config, 512, 50)
    again.reset(seed=123)
    rollout(again, 20)
    for name in ('fires', 'intensity', 'fuel', 'suppressants', 'equipment'):
        assert torch.equal(getattr(full.state(), name), getattr(again.state(), name)), name
    # the second half of the batch run as its own shard (env_offset) reproduces the same environments
    shard = make_env(config, 256, 50, env_offset=256)
    shard.reset(seed=123)
    rollout(shard, 20)
    for name in ('fires', 'intensity', 'fuel', 'suppressants', 'equipment'):
        assert torch.equal(getattr(full.state(), name)[256:], getattr(shard.state(), name)), name
    other = make_env(config, 512, 50)
    other.reset(seed=124)
    rollout(other, 20)
    assert not torch.equal(full.state().fires, other.state().fires)


def test_full_size_invariants():
    """BASELINE config C4 at its full size (65,536 envs): properties that hold for every step of any rollout."""
    B = 65536
    env = make_env(presets.wildfire_large(), B, 30)
    env.reset(seed=9)
    raw = env.unwrapped
    total = torch.zeros((B, 10), device='cuda')
    prev_terminated = raw.terminated.clone()
    for _ in range(30):
        raw.sample_actions(77)
        raw.step_all()
        s = raw.state()
        lit = (s.fires > 0).flatten(1)
        assert torch.equal(lit.sum(1).int(), raw.environment_task_count)
        assert torch.equal(raw.action_mask.sum(2).int(), raw._agent_task_count)
        assert (raw._agent_task_count <= raw.environment_task_count[:, None]).all()
        # masks only cover existing tasks; padded task rows are -100 and real rows are the lit cells in row-major order
        steps_ = torch.arange(100, device='cuda')[None, :]
        assert ((raw.action_mask != 0) <= (steps_ < raw.environment_task_count[:, None])[:, None, :]).all()
        real = steps_ < raw.environment_task_count[:, None]
        assert (raw._task_obs[~real] == -100).all()
        cells = raw._task_obs[..., 0] * 10 + raw._task_obs[..., 1]
        assert torch.equal(cells[real], lit.nonzero()[:, 1].int())
        assert (raw._task_obs[..., 2][real] > 0).all()
        # termination is monotone and zeroes the fire grid (wildfire.py:570)
        assert (raw.terminated >= prev_terminated).all()
        assert (s.fires[raw.terminated] == 0).all()
        prev_terminated = raw.terminated.clone()
        assert (s.suppressants >= 0).all() and (s.equipment >= 0).all() and (s.equipment <= 2).all()
        assert (s.intensity >= 0).all() and (s.intensity <= 4).all()
        total += raw._rewards
    assert torch.allclose(total, raw._cumulative, rtol=1e-5, atol=1e-4)
    assert (raw.num_moves == 30).all() and raw.truncated.all()
    raw.check_errors()
    # every env is now truncated: a further step is a device-side no-op (utils/env.py:212)
    before = raw.state().clone()
    raw.sample_actions(78)
    raw.step_all()
    assert torch.equal(before.fires, raw.state().fires) and (raw.num_moves == 30).all()


def test_cuda_graph_replay_matches_eager():
    config = presets.wildfire_3x3()
    eager = make_env(config, 2048, 100)
    eager.reset(seed=3)
    rollout(eager, 12, sampler_seed=2026)
    graphed = make_env(config, 2048, 100)
    graphed.reset(seed=3)
    raw = graphed.unwrapped
    raw.capture_graph(sample=True, sampler_seed=2026)
    raw.reset(seed=3)  # capture warm-up stepped once; start over with the same seed
    for _ in range(12):
        raw.replay()
    torch.cuda.synchronize()
    for name in ('fires', 'intensity', 'fuel', 'suppressants', 'capacity', 'equipment'):
        assert torch.equal(getattr(eager.state(), name), getattr(graphed.state(), name)), name
    assert torch.equal(eager.unwrapped._cumulative, raw._cumulative)


def test_partial_reset_restores_selected_envs():
    env = make_env(presets.wildfire_large(), 256, 100)
    env.reset(seed=4)
    raw = env.unwrapped
    initial = raw.state().clone()
    rollout(env, 10)
    moved = raw.state().clone()
    picked = torch.tensor([0, 5, 77, 255], device='cuda')
    env.reset_batches(picked)
    torch.cuda.synchronize()
    keep = torch.ones(256, dtype=torch.bool, device='cuda')
    keep[picked] = False
    for name in ('fires', 'intensity', 'fuel', 'suppressants', 'capacity', 'equipment'):
        assert torch.equal(getattr(raw.state(), name)[picked], getattr(initial, name)[picked]), name
        assert torch.equal(getattr(raw.state(), name)[keep], getattr(moved, name)[keep]), name
    assert (raw.num_moves[picked] == 0).all() and (raw.num_moves[keep] == 10).all()
    assert (raw._cumulative[picked] == 0).all() and (raw.num_burnouts[picked] == 0).all()
    lit = (raw.state().fires > 0).flatten(1).sum(1).int()
    assert torch.equal(lit, raw.environment_task_count)


def test_action_mapping_wrapper_returns_the_reference_mappings():
    """wrappers/action_task.py: observations come back as (observation, {'agent_action_mapping': jagged indices});
    the mappings equal the reference's recorded ones on a golden trajectory."""
    from free_range_zoo_b200.wrappers import action_mapping_wrapper_v0
    meta, gold = G.load('wildfire_c4')
    env = action_mapping_wrapper_v0(make_env(getattr(presets, meta['preset'])(**meta.get('preset_kwargs', {})), meta['B'],
                                             meta['max_steps'], **meta['env_kwargs']))
    observations, _ = env.reset(seed=0)
    agents = env.agents
    for t in range(3):
        for i, agent in enumerate(agents):
            observation, extra = observations[agent]
            mapping = extra['agent_action_mapping'].to_padded_tensor(-100).cpu().numpy()
            want = gold['action_map'][t][i]
            assert np.array_equal(mapping, want[:, :mapping.shape[1]]) and (want[:, mapping.shape[1]:] == -100).all()
            assert observation['self'].shape == (meta['B'], 4)
        env.unwrapped.inject_uniforms(torch.from_numpy(gold['u_field'][t]), torch.from_numpy(gold['u_agent'][t]))
        actions = torch.from_numpy(gold['actions'][t]).cuda()
        observations, _, _, _, _ = env.step({a: actions[:, i] for i, a in enumerate(agents)})


def test_burnt_out_and_live_environments_share_a_warp():
    """Sub-warp groups: an environment with nothing lit (a termination candidate, which triggers the fuel reduction)
    next to a live one in the same warp must step like the oracle -- every lane has to take part in the group-wide
    shuffles whatever its own environment looks like."""
    from oracle.wildfire import WildfireOracle
    config = presets.wildfire_large(height=7, width=8, num_agents=5, seed=5)
    B, steps = 64, 6
    oracle = WildfireOracle(config, B, steps)
    oracle.reset()
    env = make_env(config, B, steps)
    env.reset(seed=1)
    state = env.unwrapped.state()
    state.fires[::2] = -state.fires[::2].abs()  # every other environment: everything put out, fuel left
    state.intensity[::2] = 0
    oracle.fires[::2] = -np.abs(oracle.fires[::2])
    oracle.intensity[::2] = 0
    env.unwrapped.update_observations()
    env.unwrapped.update_actions()
    oracle.update_observations()
    oracle.update_actions()
    rng = np.random.default_rng(3)
    H, W, A = oracle.H, oracle.W, oracle.A
    for t in range(steps):
        counts = oracle.agent_task_count
        k = np.minimum((rng.random((B, A)) * (counts + 1)).astype(np.int64), counts)
        actions = np.stack([k, np.where(k == counts, -1, 0)], axis=2).astype(np.int32)
        u_field, u_agent = rng.random((3, B, H, W), dtype=np.float32), rng.random((5, B, A), dtype=np.float32)
        assert oracle.step(actions, u_field, u_agent)
        env.unwrapped.inject_uniforms(torch.from_numpy(u_field), torch.from_numpy(u_agent))
        env.step(torch.from_numpy(actions).cuda())
        want = {key: value[None] for key, value in oracle.outputs().items()}
        G.compare({k_: v for k_, v in wildfire_outputs(env).items() if k_ in ('fires', 'intensity', 'fuel', 'rewards', 'terminated', 'env_task_count', 'agent_task_count')},
                  want, 0, context=f'mixed t={t}')


def test_sampled_actions_are_uniform_over_the_legal_choices():
    """wildfire_sample_kernel: an agent with n reachable fires picks each of its n + 1 choices (the fires, then the
    task-agnostic noop / refill slot) with equal probability -- chi-square over 262 144 environments that share one
    state (same initial grid, no step taken), per agent."""
    from scipy.stats import chi2
    B = 262144
    env = make_env(presets.wildfire_large(), B, 100)
    env.reset(seed=9)
    raw = env.unwrapped
    raw.sample_actions(31)
    actions, counts = raw._actions.cpu().numpy(), raw._agent_task_count.cpu().numpy()
    assert (counts == counts[0]).all()  # identical environments
    for a, n in enumerate(counts[0]):
        picks = actions[:, a, 0]
        assert picks.min() >= 0 and picks.max() <= n
        observed = np.bincount(picks, minlength=n + 1)
        statistic = ((observed - B / (n + 1))**2 / (B / (n + 1))).sum()
        assert statistic < chi2.ppf(1 - 1e-6, df=max(int(n), 1)), (a, n, observed)
        assert (actions[picks == n, a, 1] == -1).all() and (actions[picks < n, a, 1] == 0).all()


@pytest.mark.parametrize('spec,B,kwargs', [
    ('wildfire_3x3', 50000, {}),  # more than one round of tiles per CTA, partial last tile
    (dict(height=5, width=6, num_agents=6, seed=21), 3000, dict(show_bad_actions=True)),
    (dict(height=2, width=3, num_agents=3, seed=31), 1111, {}),
])
def test_tiled_and_group_kernels_draw_the_same_trajectories(spec, B, kwargs):
    """In production (Philox) mode the two step kernels of a small grid are interchangeable: same random streams, so
    the same trajectory bit for bit -- the batch size (and with it how a batch is sharded over GPUs) only selects
    which one runs."""
    config = presets.wildfire_3x3() if spec == 'wildfire_3x3' else presets.wildfire_large(**spec)
    envs = {kernel: make_env(config, B, 25, step_kernel=kernel, **kwargs) for kernel in KERNELS}
    for env in envs.values():
        env.reset(seed=77)
    groups, tiles = envs['groups'].unwrapped, envs['tiles'].unwrapped
    names = ('fires', 'intensity', 'fuel', 'suppressants', 'capacity', 'equipment')
    for t in range(30):  # runs past max_steps: the early-out of a finished batch included
        groups.sample_actions(13)
        tiles.sample_actions(13)
        assert torch.equal(groups._actions, tiles._actions), t
        groups.step_all()
        tiles.step_all()
        for name in names:
            assert torch.equal(getattr(groups.state(), name), getattr(tiles.state(), name)), (name, t)
        for name in ('_rewards', '_cumulative', '_task_obs', '_self_obs', '_agent_task_count', '_action_mask', '_burnouts',
                     '_putouts', 'num_moves', 'num_burnouts', 'environment_task_count', '_terminated', '_truncated'):
            assert torch.equal(getattr(groups, name), getattr(tiles, name)), (name, t)
        mine, theirs = groups.control_block(), tiles.control_block()
        for field in ('seed', 'step', 'alive', 'agents_with_tasks', 'error_word'):
            assert mine[field] == theirs[field], (field, t)
    groups.check_errors()
    tiles.check_errors()
