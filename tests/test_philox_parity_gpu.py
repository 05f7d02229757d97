"""GPU parity of the PRODUCTION (in-kernel Philox) step kernels against the CPU oracle.

The golden / oracle tests of test_wildfire_gpu.py and test_cyber_gpu.py inject uniforms, which selects the
``INJECTED=true`` instantiations.  The benchmarked kernels are the ``INJECTED=false`` ones: they draw Philox words
in-kernel with a geometry-dependent word layout (split layout, spare-lane agent words, shared increase / decrease and
decrease / refill words).  Here an independent host Philox (tests/philox_ref.py) reproduces that layout, expands it
into reference-shaped uniforms (wildfire.py:409-410, random_generator.py:87-146), feeds them to the oracle and
compares the Philox-mode rollout bit-for-bit (ints / masks / dones) and within 1e-5 relative (rewards).
Covers every geometry the wildfire dispatcher can pick, every cybersecurity size class, and the named configs at
their full batch sizes (C4 at 65 536, C1 at 1 024, C3 at 16 384).
"""
import numpy as np
import pytest
import torch

from free_range_zoo_b200 import presets
from tests import golden_util as G
from tests import philox_ref as P
from tests.engine_util import cpu, cyber_outputs, wildfire_outputs
from tests.test_cyber_gpu import legal_actions

pytestmark = pytest.mark.gpu

WILDFIRE_KEYS = ('fires', 'intensity', 'fuel', 'suppressants', 'capacity', 'equipment', 'rewards', 'terminated',
                 'truncated', 'num_moves', 'num_burnouts', 'burnouts', 'putouts', 'env_task_count', 'agent_task_count',
                 'self_obs', 'task_obs', 'action_map', 'bad_map')


def wildfire_env(config, B, max_steps, **kwargs):
    from free_range_zoo_b200.envs import wildfire_v0
    return wildfire_v0.parallel_env(parallel_envs=B, max_steps=max_steps, configuration=config,
                                    device=torch.device('cuda'), **kwargs)


def _config(preset):
    return getattr(presets, preset)() if isinstance(preset, str) else getattr(presets, preset[0])(**preset[1])


def fast_wildfire_compare(raw, oracle, context):
    """Vectorised comparison of everything the step writes (for batches too large for the golden-layout dump)."""
    s = raw.state()
    B, A, HW = oracle.B, oracle.A, oracle.H * oracle.W
    exact = dict(fires=s.fires, intensity=s.intensity, fuel=s.fuel, equipment=s.equipment, num_moves=raw.num_moves,
                 num_burnouts=raw.num_burnouts, burnouts=raw._burnouts, putouts=raw._putouts,
                 terminated=raw.terminated, truncated=raw.truncated, environment_task_count=raw.environment_task_count,
                 agent_task_count=raw._agent_task_count, task_obs=raw._task_obs)
    for name, tensor in exact.items():
        want = getattr(oracle, name)
        assert np.array_equal(cpu(tensor).reshape(want.shape), want), f'{context}: {name}'
    for name, tensor in dict(suppressants=s.suppressants, capacity=s.capacity, self_obs=raw._self_obs).items():
        assert np.array_equal(cpu(tensor), getattr(oracle, name)), f'{context}: {name}'
    for name, tensor in dict(rewards=raw._rewards, cumulative_rewards=raw._cumulative).items():
        assert np.allclose(cpu(tensor), getattr(oracle, name), rtol=G.FLOAT_RTOL, atol=1e-4), f'{context}: {name}'
    # action mask: byte [b, a, t] of the engine (t = env-local task) == oracle.available[b, a, cell of task t]
    mask = cpu(raw.action_mask) != 0
    lit, rank = oracle._lit, oracle._rank
    b_index, cells = np.nonzero(lit)
    expected = np.zeros((B, A, HW), bool)
    expected[b_index, :, rank[b_index, cells]] = oracle.available[b_index, :, cells]
    assert np.array_equal(mask, expected), f'{context}: action_mask'


# every geometry pick_geometry() can return: (lanes per environment, cells per lane), with and without the agents'
# words coming from spare lanes
WILDFIRE_GEOMETRIES = [
    (dict(height=2, width=3, num_agents=3, seed=31), 500, (8, 1)),
    ('wildfire_3x3', 1024, (8, 2)),  # C1 at its named batch size
    (dict(height=4, width=5, num_agents=4, seed=32), 500, (8, 3)),
    (dict(height=5, width=6, num_agents=6, seed=21), 500, (8, 4)),
    (dict(height=4, width=8, num_agents=8, seed=33), 300, (8, 4)),  # full last row: no spare lanes
    (dict(height=4, width=4, num_agents=10, seed=34), 300, (16, 1)),
    (dict(height=4, width=8, num_agents=10, seed=35), 300, (16, 2)),
    (dict(height=5, width=9, num_agents=16, seed=11), 300, (16, 3)),
    (dict(height=7, width=8, num_agents=5, seed=5), 700, (16, 4)),
    (dict(height=8, width=10, num_agents=10, seed=36), 300, (16, 5)),  # full last row, cells in shared memory
    (dict(height=9, width=10, num_agents=9, seed=37), 300, (16, 6)),
    (dict(height=10, width=10, num_agents=10, seed=1234), 1000, (16, 7)),  # C4's kernel, split layout
    (dict(height=10, width=11, num_agents=16, seed=38), 300, (16, 7)),  # split layout, too few spare lanes
    (dict(height=8, width=15, num_agents=12, seed=39), 300, (16, 8)),
    (dict(height=4, width=8, num_agents=20, seed=40), 200, (32, 1)),
    (dict(height=3, width=20, num_agents=20, seed=41), 200, (32, 2)),
    (dict(height=3, width=33, num_agents=20, seed=8), 200, (32, 4)),
    (dict(height=12, width=16, num_agents=12, seed=6), 200, (32, 8)),
    (dict(height=16, width=16, num_agents=32, seed=12), 100, (32, 8)),
]


@pytest.mark.parametrize('kernel', ['groups', 'tiles'])
@pytest.mark.parametrize('spec,B,geometry', WILDFIRE_GEOMETRIES, ids=lambda v: str(v).replace(' ', ''))
def test_wildfire_philox_rollout_matches_oracle(spec, B, geometry, kernel):
    from oracle.wildfire import WildfireOracle
    config = presets.wildfire_3x3() if spec == 'wildfire_3x3' else presets.wildfire_large(**spec)
    if kernel == 'tiles' and geometry[0] != 8:  # the one-thread-per-environment kernel steps the 8-lane grids only
        pytest.skip('grid too large for the tiled kernel')
    steps, seed, offset = 14, 0x1234_5678_9ABC_DEF0 + B, 3 * B + 1
    oracle = WildfireOracle(config, B, steps)
    oracle.reset()
    H, W, A = oracle.H, oracle.W, oracle.A
    assert P.wildfire_geometry(H, W, A) == geometry
    env = wildfire_env(config, B, steps, env_offset=offset, step_kernel=kernel)
    env.reset(seed=seed)
    rng = np.random.default_rng(17)
    envs = offset + np.arange(B)
    stepped = 0
    for t in range(steps):
        counts = oracle.agent_task_count
        k = np.minimum((rng.random((B, A)) * (counts + 1)).astype(np.int64), counts)
        actions = np.stack([k, np.where(k == counts, -1, 0)], axis=2).astype(np.int32)
        u_field, u_agent = P.wildfire_uniforms(seed, t, envs, H, W, A)
        if not oracle.step(actions, u_field, u_agent):
            break
        env.step(torch.from_numpy(actions).cuda())
        stepped += 1
        want = {key: value[None] for key, value in oracle.outputs().items()}
        G.compare({k_: v for k_, v in wildfire_outputs(env).items() if k_ in WILDFIRE_KEYS}, want, 0,
                  context=f'{spec} philox t={t}')
    assert stepped >= 5 and env.unwrapped.control_block()['step'] == stepped
    env.unwrapped.check_errors()


@pytest.mark.parametrize('kwargs', [dict(show_bad_actions=True, observe_other_power=True)])
def test_wildfire_quirks_philox_rollout_matches_oracle(kwargs):
    """Every reward / termination branch (localized put-outs, scaled burn-out penalty, bad actions) in Philox mode."""
    from oracle.wildfire import WildfireOracle
    config, B, steps, seed = presets.wildfire_quirks(), 1536, 25, 99
    oracle = WildfireOracle(config, B, steps, **kwargs)
    oracle.reset()
    H, W, A = oracle.H, oracle.W, oracle.A
    env = wildfire_env(config, B, steps, **kwargs)
    env.reset(seed=seed)
    rng = np.random.default_rng(5)
    for t in range(steps):
        counts = oracle.environment_task_count[:, None] + np.zeros((1, A), np.int32)
        k = np.minimum((rng.random((B, A)) * (counts + 1)).astype(np.int64), counts)
        actions = np.stack([k, np.where(k == counts, -1, 0)], axis=2).astype(np.int32)
        if not oracle.step(actions, *P.wildfire_uniforms(seed, t, np.arange(B), H, W, A)):
            break
        env.step(torch.from_numpy(actions).cuda())
        want = {key: value[None] for key, value in oracle.outputs().items()}
        G.compare({k_: v for k_, v in wildfire_outputs(env).items() if k_ in WILDFIRE_KEYS}, want, 0,
                  context=f'quirks philox t={t}')
    env.unwrapped.check_errors()


def test_wildfire_c4_full_size_sampled_rollout_matches_oracle():
    """The benchmarked flow itself -- [wildfire_sample_kernel -> wildfire_step_kernel<16,7,0,0>] at 65 536 envs: the
    on-device sampler's actions equal the host restatement, and every output of every step equals the oracle's."""
    from oracle.wildfire import WildfireOracle
    config, B, steps, seed, sampler_seed = presets.wildfire_large(), 65536, 4, 2026, 2026
    oracle = WildfireOracle(config, B, 1 << 30)
    oracle.reset()
    env = wildfire_env(config, B, 1 << 30)
    env.reset(seed=seed)
    raw = env.unwrapped
    envs = np.arange(B)
    for t in range(steps):
        actions = P.sampled_actions(sampler_seed, t, envs, oracle.agent_task_count)
        raw.sample_actions(sampler_seed)
        assert np.array_equal(cpu(raw._actions), actions), f'sampler t={t}'
        assert oracle.step(actions, *P.wildfire_uniforms(seed, t, envs, 10, 10, 10))
        raw.step_all()
        fast_wildfire_compare(raw, oracle, f'C4 x 65536 t={t}')
    raw.check_errors()


def test_wildfire_c4_full_size_injected_rollout_matches_oracle():
    """C4 at 65 536 envs with injected uniforms (the reference-shaped parity mode) against the oracle."""
    from oracle.wildfire import WildfireOracle
    config, B, steps = presets.wildfire_large(), 65536, 3
    oracle = WildfireOracle(config, B, 1 << 30)
    oracle.reset()
    env = wildfire_env(config, B, 1 << 30)
    env.reset(seed=3)
    raw = env.unwrapped
    rng = np.random.default_rng(23)
    for t in range(steps):
        counts = oracle.agent_task_count
        k = np.minimum((rng.random(counts.shape) * (counts + 1)).astype(np.int64), counts)
        actions = np.stack([k, np.where(k == counts, -1, 0)], axis=2).astype(np.int32)
        u_field, u_agent = rng.random((3, B, 10, 10), dtype=np.float32), rng.random((5, B, 10), dtype=np.float32)
        assert oracle.step(actions, u_field, u_agent)
        raw.inject_uniforms(torch.from_numpy(u_field), torch.from_numpy(u_agent))
        raw.step_all(torch.from_numpy(actions).cuda())
        fast_wildfire_compare(raw, oracle, f'C4 x 65536 injected t={t}')
    raw.check_errors()


# ------------------------------------------------------------------------------------------------ cybersecurity


@pytest.mark.parametrize('preset,B,steps,kwargs', [
    ('cyber_c3', 16384, 30, dict(show_bad_actions=False, partially_observable=True)),  # C3 at its named batch size
    ('cyber_c3', 129, 10, dict(show_bad_actions=False, partially_observable=True)),  # a full tile + 1 (cooperative copies)
    ('cyber_quirks', 3000, 20, dict(show_bad_actions=True, observe_other_location=True)),  # (8, 4, 4)
    (('cyber_synthetic', dict(nodes=10, attackers=5, defenders=4)), 1500, 15, dict(show_bad_actions=True)),  # (16, 8, 8)
    (('cyber_synthetic', dict(nodes=17, attackers=1, defenders=1)), 700, 15, dict(show_bad_actions=True)),  # runtime loops
    (('cyber_synthetic', dict(nodes=20, attackers=9, defenders=9)), 300, 12, dict(show_bad_actions=True)),
    (('cyber_synthetic', dict(nodes=32, attackers=16, defenders=16)), 150, 10, dict(show_bad_actions=True)),  # direct kernel
])
def test_cyber_philox_rollout_matches_oracle(preset, B, steps, kwargs):
    from free_range_zoo_b200.envs import cybersecurity_v0
    from oracle.cybersecurity import CybersecurityOracle
    config = _config(preset)
    seed, offset = (0xABCD << 32) | 77, 5 * B + 3
    oracle = CybersecurityOracle(config, B, steps, **kwargs)
    oracle.reset()
    env = cybersecurity_v0.parallel_env(parallel_envs=B, max_steps=steps, configuration=config,
                                        device=torch.device('cuda'), env_offset=offset, **kwargs)
    env.reset(seed=seed)
    rng = np.random.default_rng(13)
    envs = offset + np.arange(B)
    for t in range(steps):
        actions = legal_actions(oracle, rng)
        assert oracle.step(actions, *P.cyber_uniforms(seed, t, envs, oracle.N, oracle.n_agents))
        env.step(torch.from_numpy(actions).cuda())
        want = {key: value[None] for key, value in oracle.outputs(env.agents).items()}
        G.compare(cyber_outputs(env), want, 0, context=f'{preset} philox t={t}')
    assert not oracle.faults
    env.unwrapped.check_errors()
