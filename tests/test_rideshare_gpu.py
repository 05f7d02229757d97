"""GPU parity tests of the fused rideshare step (through the C ABI, via the public Parallel API)."""
import numpy as np
import pytest
import torch

from free_range_zoo_b200 import presets
from tests import golden_util as G
from tests.engine_util import cpu, rideshare_outputs

pytestmark = pytest.mark.gpu


def make_env(config, B, max_steps, **kwargs):
    from free_range_zoo_b200.envs import rideshare_v0
    return rideshare_v0.parallel_env(parallel_envs=B, max_steps=max_steps, configuration=config,
                                     device=torch.device('cuda'), **kwargs)


KERNELS = ['groups', 'tiles']  # a group of lanes per environment / one thread per environment (include/frz.h)


@pytest.mark.parametrize('kernel', KERNELS)
@pytest.mark.parametrize('name', G.fixtures('rideshare'))
def test_matches_reference_trajectory(name, kernel):
    meta, gold = G.load(name)
    config = getattr(presets, meta['preset'])(**meta['preset_kwargs'])
    env = make_env(config, meta['B'], meta['max_steps'], step_kernel=kernel)
    env.reset(seed=0)
    width = gold['passengers'].shape[2]
    G.compare(rideshare_outputs(env, width), gold, 0, context=name)
    for t in range(meta['steps']):
        actions = torch.from_numpy(gold['actions'][t]).cuda()
        env.step({a: actions[:, i] for i, a in enumerate(env.agents)})
        G.compare(rideshare_outputs(env, width), gold, t + 1, context=name)
    env.unwrapped.check_errors()


def policy_actions(oracle, rng, wild):
    """Random task or noop per agent; the action id is the passenger's state, or (with probability ``wild``) any of
    accept / pick / drop -- ids the reference's step code accepts without complaint."""
    B, A = oracle.B, oracle.A
    acts = np.zeros((B, A, 2), np.int32)
    for b in range(B):
        for a in range(A):
            mine = oracle._task_list(b, a)
            k = min(int(rng.random() * (len(mine) + 1)), len(mine))
            if k == len(mine):
                acts[b, a] = (k, -1)
            else:
                ident = oracle.tables[b][mine[k]][6]
                if rng.random() < wild:
                    ident = int(rng.integers(0, 3))
                acts[b, a] = (k, ident)
    return acts


@pytest.mark.parametrize('preset,preset_kwargs,B,steps,wild', [
    ('rideshare_c2', {}, 512, 90, 0.0),
    ('rideshare_c2', {}, 300, 60, 0.3),
    ('rideshare_quirks', dict(parallel_envs=200), 200, 40, 0.2),
    ('rideshare_quirks', dict(parallel_envs=128, diagonal=False, fast=True), 128, 40, 0.1),
    ('rideshare_profile', {}, 100, 20, 0.0),
    ('rideshare_c2', {}, 1, 30, 0.0),  # a single environment
    ('rideshare_c2', dict(rows=48, horizon=30), 150, 50, 0.2),  # more than 32 rows per table: two rows per lane
    ('rideshare_c2', dict(rows=61, horizon=20), 100, 40, 0.2),  # ... and a row count that is not a multiple of 4
    ('rideshare_synthetic', dict(drivers=20, rows=64), 60, 30, 0.2),  # 32-lane groups: 20 drivers, a full 64-row table
    ('rideshare_synthetic', dict(drivers=12, rows=40), 80, 30, 0.2),  # 16-lane groups, 4 rows per lane
    ('rideshare_synthetic', dict(drivers=32, rows=24), 50, 30, 0.1),  # the engine's driver limit
    # shapes the tiled kernel serves (tables with a multiple of four rows, at most eight drivers)
    ('rideshare_synthetic', dict(drivers=7, rows=24), 200, 30, 0.2),  # eight-driver variant, odd driver count
    ('rideshare_synthetic', dict(drivers=8, rows=40), 130, 30, 0.2),  # ... with 64-bit row sets, four drivers per store
    ('rideshare_synthetic', dict(drivers=6, rows=16), 97, 30, 0.2),  # ... two drivers per load, a partial last tile
    ('rideshare_synthetic', dict(drivers=2, rows=64), 70, 30, 0.1),  # four-driver variant, a full 64-row table
    ('rideshare_c2', {}, 2100, 90, 0.1),  # more tiles than one CTA's warps, partial last tile
])
@pytest.mark.parametrize('kernel', KERNELS)
def test_matches_oracle_on_random_rollouts(preset, preset_kwargs, B, steps, wild, kernel):
    from oracle.rideshare import RideshareOracle
    config = getattr(presets, preset)(**preset_kwargs)
    oracle = RideshareOracle(config, B, steps)
    oracle.reset()
    env = make_env(config, B, steps, step_kernel=kernel)
    env.reset(seed=1)
    rng = np.random.default_rng(17)
    width = oracle.K
    for t in range(steps):
        actions = policy_actions(oracle, rng, wild)
        assert oracle.step(actions)
        env.step(torch.from_numpy(actions).cuda())
        want = {key: value[None] for key, value in oracle.outputs().items()}
        G.compare(rideshare_outputs(env, width), want, 0, context=f'{preset} t={t}')
    env.unwrapped.check_errors()


@pytest.mark.parametrize('B,steps', [(16384, 6), (32768, 4)], ids=['c2_named_size_groups', 'tiles_by_batch_size'])
def test_full_size_oracle_parity(B, steps):
    """C2 at its named batch size (16 384 environments, the group kernel) and at a size the dispatcher gives to the
    tiled kernel: every output bit-exact against the oracle, actions sampled on the device."""
    from oracle.rideshare import RideshareOracle
    config = presets.rideshare_c2()
    oracle = RideshareOracle(config, B, 100)
    oracle.reset()
    env = make_env(config, B, 100)
    env.reset(seed=8)
    raw = env.unwrapped
    for t in range(steps):
        raw.sample_actions(23)
        actions = raw._actions.cpu().numpy().copy()
        assert oracle.step(actions)
        raw.step_all()
        want = {key: value[None] for key, value in oracle.outputs().items()}
        G.compare(rideshare_outputs(env, oracle.K), want, 0, context=f'B={B} t={t}')
    raw.check_errors()


@pytest.mark.parametrize('kernel', KERNELS)
def test_full_size_sampler_rollout_properties(kernel):
    """C2 at its full size (16,384 envs): invariants of any legal rollout, driven by the on-device sampler."""
    B = 16384
    env = make_env(presets.rideshare_c2(), B, 100, step_kernel=kernel)
    env.reset(seed=3)
    raw = env.unwrapped
    total = torch.zeros((B, 4), device='cuda')
    delivered = torch.zeros(B, device='cuda')
    for t in range(100):
        before = raw.environment_task_count.clone()
        raw.sample_actions(41)
        raw.step_all()
        total += raw._rewards
        s = raw.state()
        K = raw._capacity
        valid = torch.arange(K, device='cuda')[None, :] < s.passenger_count[:, None]
        table = s.passenger_table
        # every present passenger and every driver stays on the 10x10 grid
        assert ((table[..., 1:5][valid] >= 0) & (table[..., 1:5][valid] <= 9)).all()
        assert ((s.agents >= 0) & (s.agents <= 9)).all()
        states = table[..., 6]
        assert ((states[valid] >= 0) & (states[valid] <= 2)).all()
        # accepted / riding passengers are associated with a driver, unaccepted ones are not
        assert (table[..., 7][valid & (states > 0)] >= 0).all() and (table[..., 7][valid & (states == 0)] == -1).all()
        # rows are in entry order (stable compaction) and padding rows of the observation are -100
        entered = torch.where(valid, table[..., 8], 10**6)
        assert (entered[:, 1:] >= entered[:, :-1]).all()
        assert (raw._task_obs[~valid] == -100).all()
        # the agent task lists are exactly: unaccepted or own
        mask = raw.task_mask
        expect = valid[:, None, :] & ((states[:, None, :] == 0) | (table[..., 7][:, None, :] == torch.arange(4, device='cuda')[None, :, None]))
        assert torch.equal(mask, expect)
        assert torch.equal(mask.sum(2).int(), raw._agent_task_count)
        # self observation counts
        for a in range(4):
            own = valid & (table[..., 7] == a)
            assert torch.equal(raw._self_obs[:, a, 2], (own & (states == 1)).sum(1).int())
            assert torch.equal(raw._self_obs[:, a, 3], (own & (states == 2)).sum(1).int())
    assert torch.allclose(total, raw._cumulative, rtol=1e-5, atol=1e-3)
    assert (raw.num_moves == 100).all() and raw.truncated.all() and not raw.terminated.any()
    raw.check_errors()


def test_flat_table_view_and_partial_reset():
    env = make_env(presets.rideshare_c2(), 64, 100)
    env.reset(seed=4)
    raw = env.unwrapped
    first = raw.state().passengers.clone()
    assert first.shape[1] == 11 and (first[:, 8] == 0).all()
    for _ in range(15):
        raw.sample_actions(9)
        raw.step_all()
    flat = raw.state().passengers
    assert flat.shape[0] == int(raw.environment_task_count.sum())
    assert (flat[1:, 0] >= flat[:-1, 0]).all()  # sorted by environment like the reference's flat table
    moved = raw.state().clone()
    picked = torch.tensor([1, 7, 63], device='cuda')
    env.reset_batches(picked)
    torch.cuda.synchronize()
    assert (raw.num_moves[picked] == 0).all() and (raw.num_moves[0] == 15)
    assert torch.equal(raw.state().agents[picked], raw._init_agents[picked])
    assert torch.equal(raw.state().agents[0], moved.agents[0])
    again = raw.state().passengers
    assert torch.equal(again[again[:, 0] == 7][:, 1:], first[first[:, 0] == 7][:, 1:])
    assert torch.equal(raw.state().passenger_table[0, :int(moved.passenger_count[0])],
                       moved.passenger_table[0, :int(moved.passenger_count[0])])


def test_reset_reaches_every_driver_when_the_table_is_shorter_than_the_driver_list():
    """A one-row schedule (11 table words per environment) with 12 drivers (24 coordinates): the restore kernel must
    still put every driver back and zero every reward."""
    config = presets.rideshare_synthetic(drivers=12, rows=1, horizon=3, height=6, width=7)
    B = 300
    env = make_env(config, B, 50)
    env.reset(seed=0)
    raw = env.unwrapped
    start = raw.state().agents.clone()
    for _ in range(6):
        raw.sample_actions(4)
        raw.step_all()
    raw._cumulative.fill_(3.0)
    raw.state().agents.add_(1)
    picked = torch.arange(0, B, 3, device='cuda')
    env.reset_batches(picked)
    torch.cuda.synchronize()
    assert torch.equal(raw.state().agents[picked], start[picked])
    assert (raw._cumulative[picked] == 0).all() and (raw._rewards[picked] == 0).all()
    keep = torch.ones(B, dtype=torch.bool, device='cuda')
    keep[picked] = False
    assert (raw._cumulative[keep] == 3.0).all()
    env.reset(seed=0)
    torch.cuda.synchronize()
    assert torch.equal(raw.state().agents, start) and (raw._cumulative == 0).all()


def test_action_mapping_wrapper_returns_the_reference_mappings():
    """wrappers/action_task.py on rideshare: every observation comes back with the driver's action -> task mapping
    (jagged environment-local passenger indices); equal to the reference's recorded mappings on a golden trajectory."""
    from free_range_zoo_b200.wrappers import action_mapping_wrapper_v0
    meta, gold = G.load('rideshare_c2')
    env = action_mapping_wrapper_v0(make_env(getattr(presets, meta['preset'])(**meta['preset_kwargs']), meta['B'],
                                             meta['max_steps']))
    observations, _ = env.reset(seed=0)
    agents = env.agents
    for t in range(12):
        for i, agent in enumerate(agents):
            observation, extra = observations[agent]
            mapping = extra['agent_action_mapping'].to_padded_tensor(-100).cpu().numpy()
            want = gold['action_map'][t][i]
            assert np.array_equal(mapping, want[:, :mapping.shape[1]]) and (want[:, mapping.shape[1]:] == -100).all(), (t, agent)
            assert observation['self'].shape == (meta['B'], 4)
        actions = torch.from_numpy(gold['actions'][t]).cuda()
        observations, _, _, _, _ = env.step({a: actions[:, i] for i, a in enumerate(agents)})


def test_sampled_actions_are_uniform_over_the_legal_choices():
    """rideshare_sample_kernel: a driver with n tasks picks each of its n + 1 choices (the tasks, then noop) with equal
    probability -- chi-square over 131 072 environments in one state, and every task action carries the passenger's
    state as its id (spaces/actions.py:10-50)."""
    from scipy.stats import chi2
    B = 131072
    env = make_env(presets.rideshare_c2(), B, 100)
    env.reset(seed=1)
    raw = env.unwrapped
    for _ in range(6):  # a state with several passengers, identical in every environment (wildcard schedule)
        raw.sample_actions(3)
        raw.step_all()
    # the environments have diverged: test within groups of equal task count
    raw.sample_actions(77)
    actions, counts = raw._actions.cpu().numpy(), raw._agent_task_count.cpu().numpy()
    checked = 0
    for a in range(actions.shape[1]):
        for n in np.unique(counts[:, a]):
            picks = actions[counts[:, a] == n, a, 0]
            if len(picks) < 20 * (n + 1):
                continue
            assert picks.min() >= 0 and picks.max() <= n
            observed = np.bincount(picks, minlength=n + 1)
            statistic = ((observed - len(picks) / (n + 1))**2 / (len(picks) / (n + 1))).sum()
            assert statistic < chi2.ppf(1 - 1e-6, df=max(n, 1)), (a, n, observed)
            checked += 1
    assert checked >= 8
    noop = actions[..., 0] == counts
    assert (actions[..., 1][noop] == -1).all() and (actions[..., 1][~noop] >= 0).all()
