"""The CPU oracle must reproduce every trajectory recorded from the unmodified reference (tests/golden/*.npz)."""
import numpy as np
import pytest

from free_range_zoo_b200 import presets
from oracle.wildfire import WildfireOracle, spread_lut
from tests import golden_util as G


@pytest.mark.parametrize('name', G.fixtures('wildfire'))
def test_wildfire_oracle_matches_reference(name):
    meta, gold = G.load(name)
    config = getattr(presets, meta['preset'])()
    oracle = WildfireOracle(config, meta['B'], meta['max_steps'], **meta['env_kwargs'])
    np.testing.assert_array_equal(oracle.spread_weights, gold['spread_weights'])
    np.testing.assert_array_equal(oracle.spread_lut, gold['spread_lut'])
    oracle.reset()
    G.compare(oracle.outputs(), gold, 0, context=name)
    for t in range(meta['steps']):
        stepped = oracle.step(gold['actions'][t], gold['u_field'][t], gold['u_agent'][t])
        assert stepped
        G.compare(oracle.outputs(), gold, t + 1, context=name)
