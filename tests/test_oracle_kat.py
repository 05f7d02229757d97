"""Known-answer tests: the oracle's transition functions against the reference's OWN transition unit tests.

tests/golden/kat_*.npz holds every call that tests/free_range_zoo/envs/*/env/transitions/test_*.py of the reference
make into the reference's transition modules (module buffers, inputs, outputs), recorded by
tests/golden/gen_kat.py while those tests ran and passed.  Each record is replayed through the oracle.
"""
import json
import os

import numpy as np
import pytest

from oracle import cybersecurity as cy
from oracle import rideshare as rs
from oracle import wildfire as wf

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def records(domain):
    data = np.load(os.path.join(GOLDEN, f'kat_{domain}.npz'))
    index = json.loads(bytes(data['index']).decode())
    out = []
    for i, entry in enumerate(index):
        arrays = {key: data[f'r{i}/{key}'] for key in entry['keys']}
        out.append(pytest.param(entry['transition'], arrays, id=f"{i}-{entry['test'].split('.')[-1]}"))
    return out


def same(got, want, what):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, f'{what}: shape {got.shape} != {want.shape}'
    if np.issubdtype(want.dtype, np.floating):
        np.testing.assert_allclose(got.astype(np.float64), want.astype(np.float64), rtol=1e-6, atol=0, err_msg=what)
    else:
        np.testing.assert_array_equal(got.astype(np.int64), want.astype(np.int64), err_msg=what)


# ------------------------------------------------------------------------------------------------ wildfire


@pytest.mark.parametrize('transition,r', records('wildfire'))
def test_wildfire_transition(transition, r):
    fires, intensity, fuel = (r[f'in.state.{k}'].copy() for k in ('fires', 'intensity', 'fuel'))
    supp, cap, equip = (r[f'in.state.{k}'].copy() for k in ('suppressants', 'capacity', 'equipment'))
    u = r['in.randomness_source']
    if transition == 'SuppressantDecreaseTransition':
        supp = wf.suppressant_decrease(supp, r['in.used_suppressants'], u, bool(r['buf.stochastic_decrease']),
                                       r['buf.decrease_probability'])
    elif transition == 'EquipmentTransition':
        equip = wf.equipment_transition(equip, u, r['buf.equipment_states'].shape[0], bool(r['buf.stochastic_repair']),
                                        r['buf.repair_probability'], bool(r['buf.stochastic_degrade']),
                                        r['buf.degrade_probability'], bool(r['buf.critical_error']),
                                        r['buf.critical_error_probability'])
    elif transition == 'SuppressantRefillTransition':
        supp, _ = wf.suppressant_refill(supp, cap, equip, r['in.refilled_suppressants'], u,
                                        bool(r['buf.stochastic_refill']), r['buf.refill_probability'],
                                        r['buf.equipment_bonuses'])
    elif transition == 'CapacityTransition':
        supp, cap = wf.capacity_transition(supp, cap, r['in.targets'], u[0], u[1], bool(r['buf.stochastic_switch']),
                                           r['buf.tank_switch_probability'], r['buf.possible_capacities'],
                                           r['buf.capacity_probabilities'])
    elif transition == 'FireIncreaseTransition':
        wf.fire_increase(fires, intensity, fuel, r['in.attack_counts'].astype(np.float32), u,
                         int(r['buf.burnout_state']) + 1, bool(r['buf.stochastic_increase']),
                         r['buf.intensity_increase_probability'], bool(r['buf.stochastic_burnouts']),
                         r['buf.burnout_probability'])
    elif transition == 'FireDecreaseTransition':
        wf.fire_decrease(fires, intensity, fuel, r['in.attack_counts'].astype(np.float32), u,
                         bool(r['buf.stochastic_decrease']), r['buf.decrease_probability'],
                         r['buf.extra_power_decrease_bonus'])
    elif transition == 'FireSpreadTransition':
        lut = wf.spread_lut(r['buf.fire_spread_filter.weight'])
        wf.fire_spread(fires, intensity, fuel, u, lut, r['buf.fire_random_spread_weight'],
                       r['buf.ignition_temperatures'], bool(r['buf.use_fire_fuel']))
    else:
        raise AssertionError(f'unhandled reference transition {transition}')
    prefix = 'out.0.' if 'out.0.fires' in r else 'out.'
    for name, value in (('fires', fires), ('intensity', intensity), ('fuel', fuel), ('suppressants', supp),
                        ('capacity', cap), ('equipment', equip)):
        same(value, r[prefix + name], f'{transition}.{name}')


# ------------------------------------------------------------------------------------------------ cybersecurity


@pytest.mark.parametrize('transition,r', records('cybersecurity'))
def test_cybersecurity_transition(transition, r):
    network, location, presence = (r[f'in.state.{k}'].copy() for k in ('network_state', 'location', 'presence'))
    if transition == 'MovementTransition':
        location = cy.movement(location, r['in.movement_targets'], r['in.movement_mask'])
    elif transition == 'PresenceTransition':
        presence, location = cy.presence_transition(presence, location, r['in.randomness_source'],
                                                    r['buf.persist_probs'], r['buf.return_probs'],
                                                    int(r['buf.num_attackers']))
    elif transition == 'SubnetworkTransition':
        states = int(r['buf.patched_states']) + int(r['buf.vulnerable_states']) + int(r['buf.exploited_states'])
        network = cy.subnetwork(network, r['in.patches'], r['in.attacks'], r['in.randomness_source'],
                                r['buf.temperature'], bool(r['buf.stochastic_state']), states)
    else:
        raise AssertionError(f'unhandled reference transition {transition}')
    same(network, r['out.network_state'], f'{transition}.network_state')
    same(location, r['out.location'], f'{transition}.location')
    same(presence, r['out.presence'], f'{transition}.presence')


# ------------------------------------------------------------------------------------------------ rideshare


def bare_rideshare_oracle(agents, flat, directions=None, fast=False, schedule=None):
    """RideshareOracle without a configuration: per-environment tables rebuilt from the reference's flat table."""
    oracle = object.__new__(rs.RideshareOracle)
    oracle.B, oracle.A = agents.shape[0], agents.shape[1]
    oracle.agents = agents.astype(np.int32).copy()
    oracle.tables = [[] for _ in range(oracle.B)]
    offsets = np.zeros(oracle.B, np.int64)
    if flat is not None:
        for row in flat:
            oracle.tables[int(row[0])].append([int(v) for v in row])
        counts = np.bincount(flat[:, 0].astype(np.int64), minlength=oracle.B) if len(flat) else np.zeros(oracle.B, np.int64)
        offsets = np.cumsum(counts) - counts
    oracle.fast = fast
    oracle.directions = directions if directions is not None else rs.CARDINAL
    oracle.schedule = schedule
    return oracle, offsets


def flat_table(oracle):
    rows = [row for table in oracle.tables for row in table]
    return np.asarray(rows, np.int64).reshape(-1, 11)


def local_targets(targets, offsets):
    return [[int(t - offsets[b]) if t != rs.PAD else rs.PAD for t in targets[b]] for b in range(len(targets))]


def as_vectors(vectors, b):
    return [tuple(int(v) for v in vectors[b, a]) for a in range(vectors.shape[1])]


@pytest.mark.parametrize('transition,r', records('rideshare'))
def test_rideshare_transition(transition, r):
    agents = r['in.state.agents']
    flat = r.get('in.state.passengers')
    if transition == 'MovementTransition':
        directions = [tuple(int(v) for v in d) for d in r['buf.directions']]
        oracle, _ = bare_rideshare_oracle(agents, flat, directions, bool(r['buf.fast_travel']))
        cost = np.stack([oracle.movement(b, as_vectors(r['in.vectors'], b)) for b in range(oracle.B)])
        same(oracle.agents, r['out.0.agents'], 'movement.agents')
        same(flat_table(oracle), r['out.0.passengers'], 'movement.passengers')
        same(cost, r['out.1'], 'movement.distances')
    elif transition == 'PassengerEntryTransition':
        oracle, _ = bare_rideshare_oracle(agents, flat, schedule=r['buf.schedule'].astype(np.int32))
        oracle._entry(r['in.timesteps'])
        same(flat_table(oracle), r['out.passengers'], 'entry.passengers')
    elif transition == 'PassengerStateTransition':
        oracle, offsets = bare_rideshare_oracle(agents, flat)
        targets = local_targets(r['in.targets'], offsets)
        for b in range(oracle.B):
            oracle.passenger_state(b, r['in.accepts'][b], r['in.picks'][b], targets[b], as_vectors(r['in.vectors'], b),
                                   int(r['in.timesteps'][b]))
        same(flat_table(oracle), r['out.passengers'], 'state.passengers')
    elif transition == 'PassengerExitTransition':
        oracle, offsets = bare_rideshare_oracle(agents, flat)
        targets = local_targets(r['in.targets'], offsets)
        fares = []
        for b in range(oracle.B):
            vectors = as_vectors(r['in.vectors'], b)
            dist = [np.inf if all(x == rs.PAD for x in v) else rs._norm(v[0] - v[2], v[1] - v[3]) for v in vectors]
            fares.append(oracle.passenger_exit(b, r['in.drops'][b], targets[b], dist))
        same(flat_table(oracle), r['out.0.passengers'], 'exit.passengers')
        same(np.stack(fares), r['out.1'], 'exit.fares')
    else:
        raise AssertionError(f'unhandled reference transition {transition}')
