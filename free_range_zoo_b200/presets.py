"""Named configurations used by the parity fixtures and by ``bench.py``.

Every builder takes ``ns``: a namespace holding the configuration classes to instantiate.  By default that is
this package's own structures; ``tests/golden/gen_golden.py`` passes the *reference's* configuration modules so the
reference and the engine are configured from one description.  The in-repo reference fixtures
(``tests/utils/*_configs.py``, ``envs/wildfire/configs/aaai_2024.py``) are re-stated here value by value because
``/root/reference`` does not exist on the GPU box.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch


def _own(domain: str):
    import importlib
    return importlib.import_module(f'free_range_zoo_b200.envs.{domain}.env.structures.configuration')


# ------------------------------------------------------------------------------------------------ wildfire


def _wildfire(ns, *, fire_types, lit, fire_rewards, ignition_temp, agents, power, attack_range, fire, agent, reward,
              stochastic):
    ns = ns or _own('wildfire')
    fire_types = torch.as_tensor(fire_types, dtype=torch.int32)
    height, width = fire_types.shape
    return ns.WildfireConfiguration(
        grid_width=width,
        grid_height=height,
        fire_config=ns.FireConfiguration(fire_types=fire_types,
                                         lit=torch.as_tensor(lit, dtype=torch.bool),
                                         ignition_temp=torch.as_tensor(ignition_temp, dtype=torch.int32),
                                         **fire),
        agent_config=ns.AgentConfiguration(agents=torch.as_tensor(agents, dtype=torch.int32),
                                           fire_reduction_power=torch.as_tensor(power),
                                           attack_range=torch.as_tensor(attack_range),
                                           **agent),
        reward_config=ns.RewardConfiguration(fire_rewards=torch.as_tensor(fire_rewards, dtype=torch.float32), **reward),
        stochastic_config=ns.StochasticConfiguration(**stochastic),
    )


_NEUTRAL_EQUIPMENT = [[0.0, 0.0, 0.0]] * 3


def _wf_agent(**overrides):
    base = dict(
        suppressant_states=3,
        initial_suppressant=2,
        suppressant_decrease_probability=1.0,
        suppressant_refill_probability=1.0,
        initial_equipment_state=2,
        equipment_states=torch.tensor(_NEUTRAL_EQUIPMENT, dtype=torch.float32),
        repair_probability=1.0,
        degrade_probability=1.0,
        critical_error_probability=0.0,
        initial_capacity=2,
        tank_switch_probability=1.0,
        possible_capacities=torch.tensor([1, 2, 3], dtype=torch.float32),
        capacity_probabilities=torch.tensor([0.0, 1.0, 0.0], dtype=torch.float32),
    )
    base.update(overrides)
    return base


def _wf_switches(flip=(), default=False):
    names = ('special_burnout_probability', 'suppressant_refill', 'suppressant_decrease', 'tank_switch',
             'critical_error', 'degrade', 'repair', 'fire_increase', 'fire_decrease', 'fire_spread',
             'realistic_fire_spread', 'random_fire_ignition', 'fire_fuel')
    return {name: (name in flip) != default for name in names}


def wildfire_profile(ns=None):
    """C1a: the config the reference's own profiler uses -- tests/utils/wildfire_configs.py:9-92 (2x3, 3 agents)."""
    return _wildfire(
        ns,
        fire_types=[[0, 0, 0], [1, 2, 1]],
        lit=[[0, 0, 0], [1, 1, 1]],
        fire_rewards=[[0, 0, 0], [20.0, 50.0, 20.0]],
        ignition_temp=[[2, 2, 2], [2, 2, 2]],
        agents=[[0, 0], [0, 1], [0, 2]],
        power=torch.tensor([1, 1, 1], dtype=torch.int32),
        attack_range=torch.tensor([1, 1, 1], dtype=torch.int32),
        fire=dict(num_fire_states=5, intensity_increase_probability=1.0, intensity_decrease_probability=1.0,
                  extra_power_decrease_bonus=0.0, burnout_probability=1.0, base_spread_rate=3.0, max_spread_rate=67.0,
                  random_ignition_probability=0.0, cell_size=200.0, wind_direction=0.0, initial_fuel=2),
        agent=_wf_agent(),
        reward=dict(bad_attack_penalty=-100.0, burnout_penalty=-1.0, termination_reward=0.0, termination_kappa=0.0,
                    localize_putouts=False),
        stochastic=_wf_switches(),
    )


def wildfire_3x3(ns=None):
    """C1b "3x3 grid, 3 agents": aaai_2025_ol_config(2) (envs/wildfire/configs/aaai_2024.py:16-102) extended by one
    row of cells and a third agent, exactly as SURVEY.md section 8(d) prescribes."""
    return _wildfire(
        ns,
        fire_types=[[0, 0, 0], [1, 2, 1], [1, 2, 1]],
        lit=[[0, 0, 0], [0, 1, 0], [0, 0, 0]],
        fire_rewards=[[0, 0, 0], [20.0, 50.0, 20.0], [20.0, 50.0, 20.0]],
        ignition_temp=[[2, 2, 2]] * 3,
        agents=[[0, 0], [0, 2], [0, 1]],
        power=torch.tensor([1, 1, 1], dtype=torch.int32),
        attack_range=torch.tensor([1, 1, 1], dtype=torch.int32),
        fire=dict(num_fire_states=5, intensity_increase_probability=1.0, intensity_decrease_probability=0.8,
                  extra_power_decrease_bonus=0.12, burnout_probability=4 * 0.167 * 67.0 / 200.0, base_spread_rate=50.0,
                  max_spread_rate=67.0, random_ignition_probability=0.0, cell_size=200.0, wind_direction=0.25 * math.pi,
                  initial_fuel=2),
        agent=_wf_agent(suppressant_decrease_probability=1.0 / 3, suppressant_refill_probability=1.0 / 3),
        reward=dict(bad_attack_penalty=-100.0, burnout_penalty=-1.0, termination_reward=0.0, termination_kappa=0.0,
                    localize_putouts=False),
        stochastic=_wf_switches(flip=('special_burnout_probability', 'suppressant_refill', 'suppressant_decrease',
                                    'fire_spread', 'realistic_fire_spread', 'fire_increase', 'fire_decrease')),
    )


def wildfire_large(ns=None, height: int = 10, width: int = 10, num_agents: int = 10, seed: int = 1234):
    """C4: 10x10 grid, 10 agents, agent + task + frame openness all on."""
    gen = torch.Generator().manual_seed(seed)
    fire_types = torch.randint(1, 4, (height, width), generator=gen, dtype=torch.int32)
    lit = torch.rand((height, width), generator=gen) < 0.15
    agents = torch.stack([
        torch.randint(0, height, (num_agents, ), generator=gen, dtype=torch.int32),
        torch.randint(0, width, (num_agents, ), generator=gen, dtype=torch.int32)
    ], dim=1)
    return _wildfire(
        ns,
        fire_types=fire_types,
        lit=lit,
        fire_rewards=10.0 * fire_types.float(),
        ignition_temp=torch.full((height, width), 2, dtype=torch.int32),
        agents=agents,
        power=torch.ones(num_agents, dtype=torch.float32),
        attack_range=torch.full((num_agents, ), 2, dtype=torch.int32),
        fire=dict(num_fire_states=5, intensity_increase_probability=0.6, intensity_decrease_probability=0.8,
                  extra_power_decrease_bonus=0.12, burnout_probability=4 * 0.167 * 67.0 / 200.0, base_spread_rate=25.0,
                  max_spread_rate=67.0, random_ignition_probability=0.01, cell_size=200.0, wind_direction=0.25 * math.pi,
                  initial_fuel=2),
        agent=_wf_agent(suppressant_states=4, suppressant_decrease_probability=1.0 / 3,
                        suppressant_refill_probability=1.0 / 3,
                        equipment_states=torch.tensor([[-1.0, -0.5, -1.0], [0.0, 0.0, 0.0], [1.0, 0.5, 1.0]],
                                                      dtype=torch.float32),
                        repair_probability=0.5, degrade_probability=0.2, critical_error_probability=0.05,
                        tank_switch_probability=0.5,
                        capacity_probabilities=torch.tensor([0.25, 0.5, 0.25], dtype=torch.float32)),
        reward=dict(bad_attack_penalty=-100.0, burnout_penalty=-1.0, termination_reward=10.0, termination_kappa=1.0,
                    localize_putouts=False),
        stochastic=_wf_switches(default=True),
    )


def wildfire_quirks(ns=None):
    """Parity-only config that turns on every branch the named configs leave off: localized put-out rewards, scaled
    burnout penalty, deterministic fire dynamics mixed with stochastic equipment, non-square 4x5 grid, unequal
    agents, fractional equipment bonuses."""
    return _wildfire(
        ns,
        fire_types=[[1, 0, 2, 3, 1], [2, 1, 0, 1, 2], [0, 3, 1, 2, 0], [1, 2, 2, 0, 1]],
        lit=[[1, 0, 0, 1, 0], [0, 1, 0, 0, 0], [0, 0, 1, 0, 0], [0, 1, 0, 0, 1]],
        fire_rewards=[[5.0, 0.0, 7.5, 30.0, 2.0], [11.0, 3.0, 0.0, 4.0, 8.0], [0.0, 21.0, 6.0, 9.0, 0.0],
                      [1.5, 2.5, 3.5, 0.0, 4.5]],
        ignition_temp=[[1, 2, 2, 1, 2], [2, 1, 2, 2, 1], [2, 2, 1, 2, 2], [1, 2, 2, 2, 1]],
        agents=[[0, 0], [1, 3], [3, 1], [2, 2]],
        power=torch.tensor([1.0, 2.0, 0.5, 1.5], dtype=torch.float32),
        attack_range=torch.tensor([1, 2, 1, 3], dtype=torch.int32),
        fire=dict(num_fire_states=6, intensity_increase_probability=0.7, intensity_decrease_probability=0.55,
                  extra_power_decrease_bonus=0.15, burnout_probability=0.4, base_spread_rate=0.08, max_spread_rate=67.0,
                  random_ignition_probability=0.03, cell_size=200.0, wind_direction=1.1, initial_fuel=3),
        agent=_wf_agent(suppressant_states=5, initial_suppressant=3, suppressant_decrease_probability=0.6,
                        suppressant_refill_probability=0.7,
                        equipment_states=torch.tensor(
                            [[-1.0, -0.5, -1.0], [-0.5, 0.0, 0.0], [0.0, 0.25, 0.5], [1.0, 1.0, 1.0]],
                            dtype=torch.float32),
                        initial_equipment_state=3, repair_probability=0.6, degrade_probability=0.3,
                        critical_error_probability=0.1, initial_capacity=3, tank_switch_probability=0.6,
                        possible_capacities=torch.tensor([1, 2, 3, 4], dtype=torch.float32),
                        capacity_probabilities=torch.tensor([0.125, 0.375, 0.25, 0.25], dtype=torch.float32)),
        reward=dict(bad_attack_penalty=-7.0, burnout_penalty=0.0, burnout_penalty_scaled=True, termination_reward=25.0,
                    termination_kappa=3.0, localize_putouts=True),
        stochastic=_wf_switches(default=True, flip=('realistic_fire_spread', )),
    )


# ------------------------------------------------------------------------------------------------ rideshare


def _rideshare(ns, *, height, width, starts, pool_limit, diagonal, fast, schedule, reward):
    ns = ns or _own('rideshare')
    base = dict(pick_cost=-0.1, move_cost=-0.8, drop_cost=0.0, noop_cost=-1, accept_cost=0.0, pool_limit_cost=-2.0,
                use_pooling_rewards=False, use_variable_move_cost=True, use_waiting_costs=False,
                wait_limit=torch.tensor([1, 2, 3]), long_wait_time=10, general_wait_cost=-.1, long_wait_cost=-.2)
    base.update(reward)
    return ns.RideshareConfiguration(
        grid_height=height,
        grid_width=width,
        agent_config=ns.AgentConfiguration(start_positions=torch.as_tensor(starts), pool_limit=pool_limit,
                                           use_fast_travel=fast, use_diagonal_travel=diagonal),
        reward_config=ns.RewardConfiguration(**base),
        passenger_config=ns.PassengerConfiguration(schedule=torch.as_tensor(schedule, dtype=torch.int32)),
    )


def synthetic_schedule(rows: int, horizon: int, height: int, width: int, seed: int, batch_rows: int = 0,
                       parallel_envs: int = 1) -> torch.Tensor:
    """Random schedule rows (t, batch|-1, y, x, dest_y, dest_x, fare) with position != destination."""
    gen = torch.Generator().manual_seed(seed)
    sched = torch.empty((rows, 7), dtype=torch.int32)
    sched[:, 0] = torch.randint(0, horizon, (rows, ), generator=gen)
    sched[:, 1] = -1
    if batch_rows:
        sched[-batch_rows:, 1] = torch.randint(0, parallel_envs, (batch_rows, ), generator=gen).int()
    cells = height * width
    pos = torch.randint(0, cells, (rows, ), generator=gen)
    dest = (pos + torch.randint(1, cells, (rows, ), generator=gen)) % cells
    sched[:, 2], sched[:, 3] = pos // width, pos % width
    sched[:, 4], sched[:, 5] = dest // width, dest % width
    sched[:, 6] = torch.randint(1, 11, (rows, ), generator=gen)
    return sched


def rideshare_profile(ns=None):
    """tests/utils/rideshare_configs.py:7-48 verbatim (10x10, 4 drivers, 3-row schedule with one batch-specific row)."""
    return _rideshare(ns, height=10, width=10, starts=[[0, 0], [9, 9], [0, 9], [9, 0]], pool_limit=4, diagonal=False,
                      fast=False, schedule=[[0, -1, 1, 1, 1, 1, 1], [1, -1, 1, 1, 1, 1, 2], [2, 1, 1, 1, 1, 1, 3]],
                      reward={})


def rideshare_c2(ns=None, rows: int = 32, horizon: int = 80):
    """C2: reference test config with a 32-row wildcard schedule and waiting costs on."""
    return _rideshare(ns, height=10, width=10, starts=[[0, 0], [9, 9], [0, 9], [9, 0]], pool_limit=4, diagonal=False,
                      fast=False, schedule=synthetic_schedule(rows, horizon, 10, 10, seed=4321),
                      reward=dict(use_waiting_costs=True))


def rideshare_quirks(ns=None, parallel_envs: int = 16, diagonal: bool = True, fast: bool = False):
    """Parity-only: 8-direction travel, batch-specific rows, tight pool limit, fixed move cost, non-zero costs."""
    return _rideshare(ns, height=6, width=7, starts=[[0, 0], [5, 6], [2, 3]], pool_limit=2, diagonal=diagonal,
                      fast=fast,
                      schedule=synthetic_schedule(20, 25, 6, 7, seed=99, batch_rows=8, parallel_envs=parallel_envs),
                      reward=dict(use_waiting_costs=True, use_variable_move_cost=False, drop_cost=0.25,
                                  accept_cost=-0.05, noop_cost=-0.3, wait_limit=torch.tensor([2, 3, 4]),
                                  long_wait_time=5, general_wait_cost=-0.15, long_wait_cost=-0.4))


def rideshare_synthetic(ns=None, drivers: int = 20, rows: int = 64, horizon: int = 25, height: int = 12, width: int = 9,
                        seed: int = 55):
    """Parity-only: many drivers on a non-square grid with a dense schedule (exercises the widest kernel geometry)."""
    gen = torch.Generator().manual_seed(seed)
    starts = torch.stack([torch.randint(0, height, (drivers, ), generator=gen),
                          torch.randint(0, width, (drivers, ), generator=gen)], dim=1).tolist()
    return _rideshare(ns, height=height, width=width, starts=starts, pool_limit=2, diagonal=True, fast=False,
                      schedule=synthetic_schedule(rows, horizon, height, width, seed=seed + 1),
                      reward=dict(use_waiting_costs=True, use_variable_move_cost=True))


# ------------------------------------------------------------------------------------------------ cybersecurity


def _cyber(ns, *, threat, mitigation, att_presence, def_presence, def_location, att_probs, def_probs, states,
           temperature, initial_state, adjacency, state_rewards, stochastic, patch_reward=0.0, bad_action_penalty=-100.0):
    ns = ns or _own('cybersecurity')
    f32 = lambda v: torch.as_tensor(v, dtype=torch.float32)
    return ns.CybersecurityConfiguration(
        attacker_config=ns.AttackerConfiguration(initial_presence=torch.as_tensor(att_presence, dtype=torch.bool),
                                                 threat=f32(threat), persist_probs=f32(att_probs[0]),
                                                 return_probs=f32(att_probs[1])),
        defender_config=ns.DefenderConfiguration(initial_location=torch.as_tensor(def_location, dtype=torch.int32),
                                                 initial_presence=torch.as_tensor(def_presence, dtype=torch.bool),
                                                 mitigation=f32(mitigation), persist_probs=f32(def_probs[0]),
                                                 return_probs=f32(def_probs[1])),
        network_config=ns.NetworkConfiguration(patched_states=states[0], vulnerable_states=states[1],
                                               exploited_states=states[2], temperature=temperature,
                                               initial_state=torch.as_tensor(initial_state, dtype=torch.int32),
                                               adj_matrix=torch.as_tensor(adjacency, dtype=torch.bool)),
        reward_config=ns.RewardConfiguration(bad_action_penalty=bad_action_penalty, patch_reward=patch_reward,
                                             network_state_rewards=f32(state_rewards)),
        stochastic_config=ns.StochasticConfiguration(network_state=stochastic),
    )


_TRIANGLE = [[0, 1, 1], [1, 0, 1], [1, 1, 0]]


def cyber_profile(ns=None):
    """tests/utils/cybersecurity_configs.py:15-61 verbatim (3 nodes, 2 attackers + 2 defenders, deterministic)."""
    return _cyber(ns, threat=[1.0, 1.0], mitigation=[1.0, 1.0], att_presence=[True, True], def_presence=[True, True],
                  def_location=[0, 1], att_probs=([1.0, 1.0], [1.0, 1.0]), def_probs=([1.0, 1.0], [1.0, 1.0]),
                  states=(1, 1, 3), temperature=1.0, initial_state=[0, 0, 0], adjacency=_TRIANGLE,
                  state_rewards=[4.0, 0.0, -2.0, -4.0, -8.0], stochastic=False)


def cyber_c3(ns=None):
    """C3: same topology with agent presence openness and stochastic network state."""
    return _cyber(ns, threat=[1.0, 1.0], mitigation=[1.0, 1.0], att_presence=[True, True], def_presence=[True, True],
                  def_location=[0, 1], att_probs=([0.9, 0.9], [0.5, 0.5]), def_probs=([0.9, 0.9], [0.5, 0.5]),
                  states=(1, 1, 3), temperature=1.0, initial_state=[0, 0, 0], adjacency=_TRIANGLE,
                  state_rewards=[4.0, 0.0, -2.0, -4.0, -8.0], stochastic=True)


def cyber_quirks(ns=None):
    """Parity-only: 5-node asymmetric graph, 3 attackers + 2 defenders, unequal powers, temperature != 1."""
    return _cyber(ns, threat=[0.75, 1.5, 2.25], mitigation=[1.25, 2.0], att_presence=[True, False, True],
                  def_presence=[True, False], def_location=[-1, 3], att_probs=([0.8, 0.7, 0.95], [0.4, 0.6, 0.3]),
                  def_probs=([0.85, 0.6], [0.55, 0.45]), states=(2, 2, 2), temperature=2.5,
                  initial_state=[0, 3, 1, 5, 2],
                  adjacency=[[0, 1, 0, 0, 1], [1, 0, 1, 1, 0], [0, 0, 0, 1, 0], [1, 1, 1, 0, 1], [0, 0, 0, 1, 0]],
                  state_rewards=[3.0, 1.5, 0.0, -1.0, -2.5, -6.0], stochastic=True, patch_reward=-0.5,
                  bad_action_penalty=-9.0)


def cyber_synthetic(ns=None, nodes: int = 10, attackers: int = 5, defenders: int = 4, seed: int = 77):
    """Parity-only: random graph / powers / presence dynamics of a given size (exercises the kernel's size classes)."""
    gen = torch.Generator().manual_seed(seed)
    rand = lambda n, lo, hi: (lo + (hi - lo) * torch.rand(n, generator=gen)).tolist()
    adjacency = (torch.rand((nodes, nodes), generator=gen) < 0.4).int()
    adjacency.fill_diagonal_(0)
    adjacency[torch.arange(nodes), (torch.arange(nodes) + 1) % nodes] = 1  # every node has a neighbour
    return _cyber(ns, threat=rand(attackers, 0.5, 2.5), mitigation=rand(defenders, 0.5, 2.5),
                  att_presence=(torch.rand(attackers, generator=gen) < 0.7).tolist(),
                  def_presence=(torch.rand(defenders, generator=gen) < 0.7).tolist(),
                  def_location=torch.randint(-1, nodes, (defenders, ), generator=gen).tolist(),
                  att_probs=(rand(attackers, 0.6, 0.95), rand(attackers, 0.3, 0.7)),
                  def_probs=(rand(defenders, 0.6, 0.95), rand(defenders, 0.3, 0.7)), states=(2, 1, 2), temperature=1.75,
                  initial_state=torch.randint(0, 5, (nodes, ), generator=gen).tolist(), adjacency=adjacency.tolist(),
                  state_rewards=[2.0, 1.0, 0.0, -1.5, -3.0], stochastic=True, patch_reward=-0.25)


PRESETS = SimpleNamespace(
    wildfire=dict(profile=wildfire_profile, c1=wildfire_3x3, c4=wildfire_large, quirks=wildfire_quirks),
    rideshare=dict(profile=rideshare_profile, c2=rideshare_c2, quirks=rideshare_quirks),
    cybersecurity=dict(profile=cyber_profile, c3=cyber_c3, quirks=cyber_quirks),
)
