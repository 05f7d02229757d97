"""CPU tests of the host side of the engine: configuration structures, flattening into the POD parameter structs,
the Space shim, containers, the agent selector and the state container (no GPU, no kernel launches)."""
import numpy as np
import pytest
import torch

from free_range_zoo_b200 import _lib, presets
from free_range_zoo_b200.envs.cybersecurity.env import cybersecurity as cy_env
from free_range_zoo_b200.envs.rideshare.env import rideshare as rs_env
from free_range_zoo_b200.envs.wildfire.env import wildfire as wf_env
from free_range_zoo_b200.envs.wildfire.env.structures import configuration as wf_conf
from free_range_zoo_b200.envs.wildfire.env.structures.state import WildfireState
from free_range_zoo_b200.utils.containers import (LazyDict, ObservationDict, jagged_from_padded,
                                                  jagged_indices_from_mask, jagged_rows_from_mask)
from free_range_zoo_b200.utils.selector import AgentSelector
from free_range_zoo_b200.utils.spaces import BatchedActionSpace, Space
from tests import golden_util as G

# ------------------------------------------------------------------------------------------------ configurations


def test_presets_validate_and_expose_reference_properties():
    config = presets.wildfire_large()
    assert config.agent_config.num_agents == 10 and config.fire_config.burned_out == 4
    assert config.fire_spread_weights.shape == (1, 1, 3, 3)
    assert config.fire_random_spread_weight == pytest.approx(0.01)
    assert presets.cyber_c3().num_agents == 4
    assert presets.rideshare_c2().max_fare <= 10


def test_invalid_configurations_raise_value_error():
    good = presets.wildfire_profile()
    with pytest.raises(ValueError, match='mutually exclusive'):
        wf_conf.RewardConfiguration(fire_rewards=torch.zeros(2, 3), bad_attack_penalty=0, burnout_penalty=-1.0,
                                    burnout_penalty_scaled=True)
    with pytest.raises(ValueError, match='realistic fire spread'):
        wf_conf.StochasticConfiguration(**{**vars(good.stochastic_config), 'realistic_fire_spread': True})
    with pytest.raises(ValueError, match='grid_width'):
        wf_conf.WildfireConfiguration(grid_width=0, grid_height=2, fire_config=good.fire_config,
                                      agent_config=good.agent_config, reward_config=good.reward_config,
                                      stochastic_config=good.stochastic_config)


def test_configuration_to_moves_tensors():
    config = presets.cyber_quirks().to('cpu')
    assert config.attacker_config.threat.device.type == 'cpu'


# ------------------------------------------------------------------------------------------------ flattening


@pytest.mark.parametrize('name', ['wildfire_c1', 'wildfire_c4', 'wildfire_quirks'])
def test_wildfire_flattening_matches_reference_tables(name):
    """spread LUT == the reference conv2d outputs recorded in the fixture; range masks == the oracle's range test."""
    from oracle.wildfire import WildfireOracle
    meta, gold = G.load(name)
    config = getattr(presets, meta['preset'])(**meta.get('preset_kwargs', {}))
    params, cell_reward, cell_ignition, range_mask = wf_env.flatten_configuration(config, 7, True, env_offset=5)
    np.testing.assert_array_equal(np.array(list(params.spread_lut), np.float32), gold['spread_lut'])
    assert params.max_steps == 7 and params.env_offset == 5 and params.flags & _lib.WF_SHOW_BAD_ACTIONS
    H, W, A = params.height, params.width, params.num_agents
    np.testing.assert_array_equal(cell_reward.reshape(H, W), config.reward_config.fire_rewards.numpy())
    oracle = WildfireOracle(config, 1, 1)
    oracle.reset()
    for e in range(params.num_equipment_states):
        oracle.fires[:] = 1  # every cell lit, full suppressant: availability == the pure range test
        oracle.suppressants[:] = 1
        oracle.equipment[:] = e
        oracle.update_actions()
        for a in range(A):
            bits = [(int(range_mask[a, e, c >> 5]) >> (c & 31)) & 1 for c in range(H * W)]
            np.testing.assert_array_equal(np.array(bits, bool), oracle.available[0, a])


def test_wildfire_limits_are_enforced():
    with pytest.raises(ValueError, match='engine limits'):
        wf_env.flatten_configuration(presets.wildfire_large(height=20, width=20), 10, False)


def test_cyber_score_table_matches_oracle_scores():
    from oracle.cybersecurity import danger_score
    config = presets.cyber_quirks()
    params, lut = cy_env.flatten_configuration(config, 10, True)
    assert params.lut_bits == 5 and lut.shape == (32, )
    threat, mitigation = config.attacker_config.threat.numpy(), config.defender_config.mitigation.numpy()
    for index in (0, 1, 0b00110, 0b11000, 0b10101, 31):
        attacks = np.float32(0)
        for a in range(3):
            if index >> a & 1:
                attacks = np.float32(attacks + threat[a])
        patches = np.float32(0)
        for d in range(2):
            if index >> (3 + d) & 1:
                patches = np.float32(patches + mitigation[d])
        want = danger_score(np.array([patches]), np.array([attacks]), config.network_config.temperature)[0]
        assert lut[index].item() == want


def test_rideshare_schedule_is_time_sorted_and_capacity_counts_applicable_rows():
    config = presets.rideshare_quirks(parallel_envs=16)
    params, schedule = rs_env.flatten_configuration(config, 10, 16)
    assert (np.diff(schedule[:, 0]) >= 0).all()
    original = config.passenger_config.schedule.numpy()
    for t in np.unique(original[:, 0]):  # rows of one step keep their schedule order
        np.testing.assert_array_equal(schedule[schedule[:, 0] == t], original[original[:, 0] == t])
    wildcard = int((original[:, 1] == -1).sum())
    addressed = np.bincount(original[original[:, 1] >= 0, 1], minlength=16).max()
    assert params.capacity == wildcard + addressed
    shard_params, _ = rs_env.flatten_configuration(config, 10, 4, env_offset=12)
    assert shard_params.env_offset == 12 and shard_params.capacity <= params.capacity


# ------------------------------------------------------------------------------------------------ spaces


def test_space_shim_matches_reference_constructions():
    """Expected constructions of the reference's space tests (tests/.../spaces/test_action_space.py)."""
    noop_only = Space.OneOf([Space.Discrete(1, start=-1)])
    three = Space.OneOf([*[Space.Discrete(1, start=0) for _ in range(3)], Space.Discrete(1, start=-1)])
    counts = torch.tensor([0, 3], dtype=torch.int32)
    slots = torch.arange(5, dtype=torch.int32).unsqueeze(0)
    batched = BatchedActionSpace(torch.where(slots == counts.unsqueeze(1), -1, 0).to(torch.int32), counts + 1)
    assert batched.spaces == [noop_only, three]
    assert batched == Space.Vector([noop_only, three])
    assert three.spaces[1].start == 0 and three.spaces[1].n == 1 and three.spaces[-1].start == -1
    assert hash(Space.Box([0, 0], [1, 2])) == hash(Space.Box([0, 0], [1, 2]))
    assert Space.Dict({'a': noop_only}).spaces['a'] == noop_only


def test_batched_action_space_samples_legal_actions():
    counts = torch.tensor([0, 1, 4, 9], dtype=torch.int32)
    slots = torch.arange(10, dtype=torch.int32).unsqueeze(0)
    space = BatchedActionSpace(torch.where(slots == counts.unsqueeze(1), -1, 0).to(torch.int32), counts + 1)
    generator = torch.Generator().manual_seed(0)
    seen_noop = torch.zeros(4, dtype=torch.bool)
    for _ in range(200):
        sample = space.sample_tensor(generator)
        assert sample.shape == (4, 2) and sample.dtype == torch.int32
        assert (sample[:, 0] >= 0).all() and (sample[:, 0] <= counts).all()
        assert ((sample[:, 1] == -1) == (sample[:, 0] == counts)).all()
        seen_noop |= sample[:, 1] == -1
    assert seen_noop.all()
    nested = space.sample_nested()
    assert len(nested) == 4 and nested[0] == [0, -1]


# ------------------------------------------------------------------------------------------------ containers


def test_lazy_and_observation_dict():
    calls = []
    lazy = LazyDict({'x': lambda: calls.append(1) or torch.ones(2), 'y': torch.zeros(1)})
    assert not calls and torch.equal(lazy['x'], torch.ones(2)) and torch.equal(lazy['x'], torch.ones(2))
    assert calls == [1] and set(lazy) == {'x', 'y'}
    obs = ObservationDict({'self': torch.zeros(3, 4)}, batch_size=[3], device='cpu')
    obs['agent_action_mapping'] = torch.ones(3)  # what action_mapping_wrapper_v0 adds
    assert obs.batch_size == torch.Size([3]) and 'agent_action_mapping' in obs and len(obs.clone()) == 2


def test_jagged_helpers_reproduce_reference_nested_layout():
    padded = torch.tensor([[[1, 1], [2, 2], [-100, -100]], [[-100, -100]] * 3, [[5, 5], [6, 6], [7, 7]]])
    counts = torch.tensor([2, 0, 3])
    nested = jagged_from_padded(padded, counts, torch.int64)
    want = torch.nested.as_nested_tensor([padded[0, :2], padded[1, :0], padded[2, :3]], layout=torch.jagged)
    assert torch.equal(nested.to_padded_tensor(-100), want.to_padded_tensor(-100)) and nested.dtype == torch.int64
    mask = torch.tensor([[True, False, True], [False, False, False], [False, True, True]])
    assert torch.equal(jagged_indices_from_mask(mask).to_padded_tensor(-100), torch.tensor([[0, 2], [-100, -100], [1, 2]]))
    rows = jagged_rows_from_mask(padded, mask)
    assert torch.equal(rows.to_padded_tensor(-100)[0], torch.tensor([[1, 1], [-100, -100]]))


# ------------------------------------------------------------------------------------------------ runtime pieces


def test_agent_selector_cycle():
    selector = AgentSelector(['a', 'b', 'c'])
    assert selector.reset() == 'a' and selector.is_first() and not selector.is_last()
    assert selector.next() == 'b' and selector.next() == 'c' and selector.is_last()
    assert selector.next() == 'a' and selector.is_first()


def test_state_snapshots_restore_selected_rows():
    def make():
        return WildfireState(fires=torch.arange(12, dtype=torch.int32).reshape(3, 2, 2), intensity=torch.zeros(3, 2, 2),
                             fuel=torch.zeros(3, 2, 2), agents=torch.zeros(2, 2), suppressants=torch.ones(3, 2),
                             capacity=torch.ones(3, 2), equipment=torch.ones(3, 2))

    state = make()
    state.save_initial()
    state.fires += 100
    state.save_checkpoint()
    state.fires += 100
    assert len(state) == 3
    state.restore_initial(torch.tensor([1]))
    assert torch.equal(state.fires[1], make().fires[1]) and torch.equal(state.fires[0], make().fires[0] + 200)
    state.restore_from_checkpoint()
    assert torch.equal(state.fires, make().fires + 100)
    with pytest.raises(ValueError):
        make().restore_initial()
    stacked = WildfireState.cat([make(), make()], dim=0)
    assert stacked.fires.shape[0] == 6 and stacked.agents.shape == (2, 2)


def test_seed_folding_is_deterministic():
    from free_range_zoo_b200.utils.env import _seed_to_u64
    assert _seed_to_u64(7) == 7
    assert _seed_to_u64([1, 2, 3]) == _seed_to_u64(torch.tensor([1, 2, 3])) != _seed_to_u64([3, 2, 1])
    assert 0 <= _seed_to_u64(None) < 2**64


def test_state_per_environment_access_and_hash():
    """State.unwrap / __getitem__ / to_dataframe / __hash__ (reference utils/state.py:180-238)."""
    import torch

    from free_range_zoo_b200.envs.cybersecurity.env.structures.state import CybersecurityState
    from free_range_zoo_b200.envs.rideshare.env.structures.state import RideshareState
    from free_range_zoo_b200.envs.wildfire.env.structures.state import WildfireState
    i32 = torch.int32
    wildfire = WildfireState(fires=torch.arange(12, dtype=i32).view(3, 2, 2), intensity=torch.zeros(3, 2, 2, dtype=i32),
                             fuel=torch.ones(3, 2, 2, dtype=i32), agents=torch.tensor([[0, 1], [1, 1]], dtype=i32),
                             suppressants=torch.ones(3, 2), capacity=torch.ones(3, 2), equipment=torch.zeros(3, 2, dtype=i32))
    singles = wildfire.unwrap()
    assert len(wildfire) == 3 and len(singles) == 3
    assert torch.equal(singles[2].fires, wildfire.fires[2]) and singles[2].agents is wildfire.agents  # shared field
    frame = wildfire.to_dataframe()
    assert list(frame.columns) == ['fires', 'intensity', 'fuel', 'suppressants', 'capacity', 'equipment', 'agents']
    assert frame.shape == (3, 7) and frame['fires'][1] == '[[4, 5], [6, 7]]' and frame['agents'][2] == '[[0, 1], [1, 1]]'
    assert hash(wildfire) == hash(wildfire.clone()) and hash(wildfire) != hash(wildfire[torch.tensor([0, 0, 0])])
    assert wildfire == wildfire and wildfire != wildfire.clone()  # identity, not element-wise tensor comparison

    rideshare = RideshareState(agents=torch.zeros(2, 4, 2, dtype=i32), passenger_table=torch.ones(2, 3, 11, dtype=i32),
                               passenger_count=torch.tensor([1, 2], dtype=i32))
    assert [int(s.passenger_count) for s in rideshare.unwrap()] == [1, 2]
    assert rideshare.passengers.shape == (3, 11) and rideshare.passengers[:, 0].tolist() == [0, 1, 1]
    cyber = CybersecurityState(network_state=torch.zeros(4, 3, dtype=i32), location=torch.zeros(4, 2, dtype=i32),
                               presence=torch.ones(4, 4, dtype=torch.bool))
    assert len(cyber.unwrap()) == 4 and cyber.to_dataframe().shape == (4, 3)
    assert isinstance(hash(cyber), int)
