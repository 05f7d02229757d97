"""The CPU oracle must reproduce every trajectory recorded from the unmodified reference (tests/golden/*.npz)."""
import numpy as np
import pytest

from free_range_zoo_b200 import presets
from oracle.wildfire import WildfireOracle, spread_lut
from tests import golden_util as G


@pytest.mark.parametrize('name', G.fixtures('wildfire'))
def test_wildfire_oracle_matches_reference(name):
    meta, gold = G.load(name)
    config = getattr(presets, meta['preset'])(**meta.get('preset_kwargs', {}))
    oracle = WildfireOracle(config, meta['B'], meta['max_steps'], **meta['env_kwargs'])
    np.testing.assert_array_equal(oracle.spread_weights, gold['spread_weights'])
    np.testing.assert_array_equal(oracle.spread_lut, gold['spread_lut'])
    oracle.reset()
    G.compare(oracle.outputs(), gold, 0, context=name)
    for t in range(meta['steps']):
        stepped = oracle.step(gold['actions'][t], gold['u_field'][t], gold['u_agent'][t])
        assert stepped
        G.compare(oracle.outputs(), gold, t + 1, context=name)


@pytest.mark.parametrize('name', G.fixtures('cyber'))
def test_cybersecurity_oracle_matches_reference(name):
    from oracle.cybersecurity import CybersecurityOracle
    meta, gold = G.load(name)
    config = getattr(presets, meta['preset'])(**meta.get('preset_kwargs', {}))
    oracle = CybersecurityOracle(config, meta['B'], meta['max_steps'], **meta['env_kwargs'])
    oracle.reset()
    G.compare(oracle.outputs(meta['agents']), gold, 0, context=name)
    for t in range(meta['steps']):
        assert oracle.step(gold['actions'][t], gold['u_network'][t], gold['u_agent'][t])
        G.compare(oracle.outputs(meta['agents']), gold, t + 1, context=name)
    assert not oracle.faults


@pytest.mark.parametrize('name', G.fixtures('rideshare'))
def test_rideshare_oracle_matches_reference(name):
    from oracle.rideshare import RideshareOracle
    meta, gold = G.load(name)
    config = getattr(presets, meta['preset'])(**meta['preset_kwargs'])
    oracle = RideshareOracle(config, meta['B'], meta['max_steps'])
    oracle.reset()
    G.compare(oracle.outputs(), gold, 0, context=name)
    for t in range(meta['steps']):
        assert oracle.step(gold['actions'][t])
        G.compare(oracle.outputs(), gold, t + 1, context=name)
