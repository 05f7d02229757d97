"""Helpers shared by the oracle (CPU) and engine (GPU) parity tests: load golden fixtures, compare output dicts."""
import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

# float outputs are compared with the tolerance north_star states (1e-5 relative); everything else bit-exact
FLOAT_RTOL = 1e-5
FLOAT_ATOL = 1e-6


def fixtures(domain: str):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, f'{domain}_*.npz')))


def load(name: str):
    data = dict(np.load(os.path.join(GOLDEN_DIR, f'{name}.npz')))
    meta = json.loads(bytes(data.pop('meta')).decode())
    return meta, data


def compare(got: dict, want: dict, step: int, keys=None, context: str = ''):
    """Assert ``got[key] == want[key][step]`` for every golden key (ints/bools exact, floats within tolerance)."""
    checked = 0
    for key, value in got.items():
        if key not in want or (keys is not None and key not in keys):
            continue
        expected = want[key][step]
        value = np.asarray(value)
        assert value.shape == expected.shape, f'{context} step {step} {key}: shape {value.shape} != {expected.shape}'
        if np.issubdtype(expected.dtype, np.floating):
            ok = np.isclose(value.astype(np.float64), expected.astype(np.float64), rtol=FLOAT_RTOL, atol=FLOAT_ATOL)
        else:
            ok = value.astype(np.int64) == expected.astype(np.int64)
        if not ok.all():
            bad = np.argwhere(~ok)[:5]
            raise AssertionError(f'{context} step {step} {key}: {(~ok).sum()} mismatches, first at {bad.tolist()} '
                                 f'got {value[tuple(bad[0])]} want {expected[tuple(bad[0])]}')
        checked += 1
    assert checked > 0
    return checked
