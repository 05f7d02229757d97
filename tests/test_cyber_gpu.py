"""GPU parity tests of the fused cybersecurity step (through the C ABI, via the public Parallel API)."""
import numpy as np
import pytest
import torch

from free_range_zoo_b200 import presets
from tests import golden_util as G
from tests.engine_util import cpu, cyber_outputs

pytestmark = pytest.mark.gpu


def make_env(config, B, max_steps, **kwargs):
    from free_range_zoo_b200.envs import cybersecurity_v0
    return cybersecurity_v0.parallel_env(parallel_envs=B, max_steps=max_steps, configuration=config,
                                         device=torch.device('cuda'), **kwargs)


@pytest.mark.parametrize('name', G.fixtures('cyber'))
def test_matches_reference_trajectory(name):
    meta, gold = G.load(name)
    env = make_env(getattr(presets, meta['preset'])(**meta.get('preset_kwargs', {})), meta['B'], meta['max_steps'], **meta['env_kwargs'])
    env.reset(seed=0)
    assert list(env.agents) == meta['agents']
    G.compare(cyber_outputs(env), gold, 0, context=name)
    for t in range(meta['steps']):
        env.unwrapped.inject_uniforms(torch.from_numpy(gold['u_network'][t]), torch.from_numpy(gold['u_agent'][t]))
        actions = torch.from_numpy(gold['actions'][t]).cuda()
        env.step({a: actions[:, i] for i, a in enumerate(env.agents)})
        G.compare(cyber_outputs(env), gold, t + 1, context=name)
    env.unwrapped.check_errors()


def legal_actions(oracle, rng):
    """Uniform over the legal choices of each agent (same rule as tests/golden/gen_golden.py::cyber_actions)."""
    B, n = oracle.B, oracle.n_agents
    acts = np.zeros((B, n, 2), np.int32)
    for a in range(n):
        count = np.full(B, oracle.N) if oracle.show_bad_actions else oracle.agent_task_count[:, a]
        choices = count + 1
        can_patch = np.zeros(B, bool)
        if a >= oracle.n_att:
            can_patch = (count > 0) & (oracle.show_bad_actions | (oracle.location[:, a - oracle.n_att] != -1))
            choices = choices + np.where(count > 0, 1 + can_patch, 0)
        k = np.minimum((rng.random(B) * choices).astype(np.int64), choices - 1)
        ident = np.where(k < count, 0, np.where(k == count, -1, np.where((k == count + 1) & can_patch, -2, -3)))
        acts[:, a, 0], acts[:, a, 1] = k, ident
    return acts


@pytest.mark.parametrize('preset,B,steps,kwargs', [
    ('cyber_c3', 16384, 40, dict(show_bad_actions=False, partially_observable=True)),
    ('cyber_quirks', 5000, 30, dict(show_bad_actions=True, observe_other_location=True)),
    ('cyber_profile', 777, 12, {}),
    ('cyber_c3', 1, 10, dict(show_bad_actions=False, partially_observable=True)),  # a single environment
    ('cyber_c3', 129, 10, dict(show_bad_actions=False, partially_observable=True)),  # one full tile + 1
    # the kernel's other size classes: (16, 8, 8) tiled; runtime-loop tiled; too large for a tile -> direct kernel
    (('cyber_synthetic', dict(nodes=10, attackers=5, defenders=4)), 1500, 20, dict(show_bad_actions=True)),
    (('cyber_synthetic', dict(nodes=17, attackers=1, defenders=1)), 700, 20, dict(show_bad_actions=True)),
    (('cyber_synthetic', dict(nodes=20, attackers=9, defenders=9)), 300, 15, dict(show_bad_actions=True)),
    (('cyber_synthetic', dict(nodes=32, attackers=16, defenders=16)), 150, 10, dict(show_bad_actions=True)),  # size limits
])
def test_matches_oracle_on_random_rollouts(preset, B, steps, kwargs):
    from oracle.cybersecurity import CybersecurityOracle
    config = getattr(presets, preset)() if isinstance(preset, str) else getattr(presets, preset[0])(**preset[1])
    oracle = CybersecurityOracle(config, B, steps, **kwargs)
    oracle.reset()
    env = make_env(config, B, steps, **kwargs)
    env.reset(seed=1)
    rng = np.random.default_rng(11)
    for t in range(steps):
        actions = legal_actions(oracle, rng)
        u_network = rng.random((1, B, oracle.N), dtype=np.float32)
        u_agent = rng.random((1, B, oracle.n_agents), dtype=np.float32)
        assert oracle.step(actions, u_network, u_agent)
        env.unwrapped.inject_uniforms(torch.from_numpy(u_network), torch.from_numpy(u_agent))
        env.step(torch.from_numpy(actions).cuda())
        want = {key: value[None] for key, value in oracle.outputs(env.agents).items()}
        G.compare(cyber_outputs(env), want, 0, context=f'{preset} t={t}')
    assert not oracle.faults
    env.unwrapped.check_errors()


def test_invalid_actions_raise_lazily():
    env = make_env(presets.cyber_c3(), 64, 10, show_bad_actions=False)
    env.reset(seed=2)
    actions = torch.zeros((64, 4, 2), dtype=torch.int32, device='cuda')
    actions[:, :, 1] = -1
    actions[3, 0] = torch.tensor([7, 0])  # attack a node that does not exist (reference: ValueError, :341)
    env.step(actions)
    with pytest.raises(ValueError, match='target outside'):
        env.unwrapped.check_errors()
    env.unwrapped.check_errors()  # cleared


def test_full_size_philox_rollout_properties():
    """C3 at 16,384 envs and the C5 size 524,288: presence statistics and invariants under in-kernel randomness."""
    for B in (16384, 524288):
        env = make_env(presets.cyber_c3(), B, 100, show_bad_actions=False, partially_observable=True)
        env.reset(seed=5)
        raw = env.unwrapped
        total = torch.zeros((B, 4), device='cuda')
        for _ in range(60):
            raw.sample_actions(31)
            raw.step_all()
            total += raw._rewards
        s = raw.state()
        assert (s.network_state >= 0).all() and (s.network_state <= 4).all()
        assert ((s.location >= -1) & (s.location <= 2)).all()
        assert torch.equal(raw._agent_task_count, s.presence.int() * 3)
        # stationary presence of the two-state chain: return / (return + 1 - persist) = 0.5 / 0.6
        assert abs(s.presence.float().mean().item() - 0.5 / 0.6) < 0.01
        assert torch.allclose(total, raw._cumulative, rtol=1e-5, atol=1e-3)
        # attackers and defenders receive opposite network rewards (patch_reward = 0 in this config)
        assert torch.equal(raw._rewards[:, 0], -raw._rewards[:, 2])
        assert (raw.num_moves == 60).all() and not raw.terminated.any()
        raw.check_errors()


def test_action_mapping_wrapper_returns_the_reference_mappings():
    """wrappers/action_task.py on cybersecurity: every observation comes back with the agent's action -> task mapping
    (all N subnetworks while the agent is present, nothing while it is away: cybersecurity.py:414-446); equal to the
    reference's recorded mappings on a golden trajectory with presence openness."""
    from free_range_zoo_b200.wrappers import action_mapping_wrapper_v0
    meta, gold = G.load('cyber_c3')
    env = action_mapping_wrapper_v0(make_env(getattr(presets, meta['preset'])(**meta.get('preset_kwargs', {})), meta['B'],
                                             meta['max_steps'], **meta['env_kwargs']))
    observations, _ = env.reset(seed=0)
    agents = env.agents
    for t in range(15):
        for agent in agents:
            observation, extra = observations[agent]
            mapping = extra['agent_action_mapping'].to_padded_tensor(-100).cpu().numpy()
            want = gold[f'action_map__{agent}'][t]
            assert np.array_equal(mapping, want[:, :mapping.shape[1]]) and (want[:, mapping.shape[1]:] == -100).all(), (t, agent)
        env.unwrapped.inject_uniforms(torch.from_numpy(gold['u_network'][t]), torch.from_numpy(gold['u_agent'][t]))
        actions = torch.from_numpy(gold['actions'][t]).cuda()
        observations, _, _, _, _ = env.step({a: actions[:, i] for i, a in enumerate(agents)})


def test_sampled_actions_are_uniform_over_the_legal_choices():
    """cyber_sample_kernel: chi-square of the sampled choice index over 262 144 environments, per agent and per number
    of legal choices (attackers: N nodes + noop; defenders: N moves + noop + monitor [+ patch when not at home])."""
    from scipy.stats import chi2
    B = 262144
    env = make_env(presets.cyber_c3(), B, 100, show_bad_actions=False, partially_observable=True)
    env.reset(seed=5)
    raw = env.unwrapped
    for _ in range(4):
        raw.sample_actions(11)
        raw.step_all()
    raw.sample_actions(12)
    actions = raw._actions.cpu().numpy()
    present = raw.state().presence.cpu().numpy().astype(bool)
    location = raw.state().location.cpu().numpy()
    n_att, N = raw._n_att, raw._n_nodes
    for a in range(actions.shape[1]):
        ident, k = actions[:, a, 1], actions[:, a, 0]
        assert (ident[~present[:, a]] == -1).all()  # an absent agent can only pass
        groups = [present[:, a]]
        if a >= n_att:  # defenders at home cannot patch
            home = location[:, a - n_att] == -1
            groups = [present[:, a] & home, present[:, a] & ~home]
        for members in groups:
            if members.sum() < 1000:
                continue
            # the choice as one category: node 0 .. N-1, then the task-agnostic ids
            category = np.where(ident[members] == 0, k[members], N - 1 - ident[members])
            observed = np.bincount(category)
            observed = observed[observed > 0]
            expected = members.sum() / len(observed)
            statistic = ((observed - expected)**2 / expected).sum()
            assert statistic < chi2.ppf(1 - 1e-6, df=len(observed) - 1), (a, observed)
            assert len(observed) >= N + 1
