"""CPU checks of the host Philox restatement the Philox-mode GPU parity tests rely on (tests/philox_ref.py)."""
import numpy as np
import pytest

from free_range_zoo_b200 import presets
from tests import philox_ref as P
from tests.test_philox_parity_gpu import WILDFIRE_GEOMETRIES


@pytest.mark.parametrize('key,counter,expected', [
    # known-answer vectors of Philox4x32-10 published with Random123 (kat_vectors)
    ((0, 0), (0, 0, 0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff), (0xffffffff, ) * 4, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0xa4093822, 0x299f31d0), (0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
])
def test_philox_known_answers(key, counter, expected):
    assert tuple(int(word) for word in P.philox4x32_10(key, counter)) == expected


def test_u01_support():
    u = P.u01(np.array([0, 0xff, 0x100, 0xffffffff], dtype=np.uint32))
    assert u.dtype == np.float32 and u[0] == 0 and u[1] == 0 and u[2] == np.float32(2.0 ** -24) and u[3] < 1


@pytest.mark.parametrize('spec,B,geometry', WILDFIRE_GEOMETRIES, ids=lambda v: str(v).replace(' ', ''))
def test_wildfire_layout_never_reuses_a_word(spec, B, geometry):
    """Every (cell, event) and (agent, event) draw of an environment-step comes from its own Philox word -- except the
    two intended sharings (increase / decrease of a cell, suppressant decrease / refill of an agent)."""
    config = presets.wildfire_3x3() if spec == 'wildfire_3x3' else presets.wildfire_large(**spec)
    H, W, A = config.grid_height, config.grid_width, config.agent_config.agents.shape[0]
    assert P.wildfire_geometry(H, W, A) == geometry
    layout = P.wildfire_layout(H, W, A)
    words = [tuple(w) for table in (layout['grow'], layout['spread'], layout['agent'].reshape(-1, 2)) for w in table.tolist()]
    assert len(set(words)) == len(words) == 2 * H * W + 4 * A
    u_field, u_agent = P.wildfire_uniforms(7, 3, np.array([0, 5, 2 ** 33 + 1]), H, W, A)
    assert u_field.shape == (3, 3, H, W) and u_agent.shape == (5, 3, A)
    assert np.array_equal(u_field[0], u_field[1]) and np.array_equal(u_agent[0], u_agent[2])
    assert (u_field >= 0).all() and (u_field < 1).all()
    # environments and steps are independent streams
    again = P.wildfire_uniforms(7, 4, np.array([0, 5, 2 ** 33 + 1]), H, W, A)[0]
    assert not np.array_equal(again, u_field) and not np.array_equal(u_field[:, 0], u_field[:, 1])


def test_wildfire_layout_covers_both_agent_word_sources():
    fed = [P.wildfire_layout(*P_) ['spare_lanes_feed_agents'] for P_ in ((10, 10, 10), (8, 10, 10), (10, 11, 16), (4, 8, 8))]
    assert fed == [True, False, False, False]
    assert P.wildfire_layout(10, 10, 10)['split'] and not P.wildfire_layout(9, 10, 9)['split']


def test_cyber_uniform_layout():
    u_network, u_agent = P.cyber_uniforms(11, 2, np.arange(5), 10, 9)
    assert u_network.shape == (1, 5, 10) and u_agent.shape == (1, 5, 9)
    raw = P._calls(11, np.arange(5), 2, np.array([1], dtype=np.uint32))
    assert np.array_equal(u_network[0, :, 4:8], P.u01(raw[:, 0, :]))


def test_tiled_wildfire_kernel_uses_the_group_kernels_word_layout():
    """frz_wildfire_tile_random_layout (the table the one-thread-per-environment kernel draws from) == the word layout of
    the group kernel as restated in tests/philox_ref.py, for EVERY small grid (<= 32 cells, <= 8 agents): the two
    kernels consume the same Philox words for the same events, so the batch size -- which selects the kernel -- never
    changes a trajectory."""
    import ctypes

    from free_range_zoo_b200 import _lib
    from tests import philox_ref as P
    lib = _lib.library()
    streams, destinations = (ctypes.c_uint32 * 24)(), (ctypes.c_int8 * 96)()
    checked = 0
    for H in range(1, 9):
        for W in range(1, 33):
            if H * W > 32:
                continue
            for A in range(1, 9):
                params = _lib.WildfireParams()
                params.height, params.width, params.num_agents = H, W, A
                calls = lib.frz_wildfire_tile_random_layout(ctypes.byref(params), streams, destinations)
                assert calls > 0, (H, W, A, lib.frz_last_error())
                layout = P.wildfire_layout(H, W, A)
                assert layout['G'] == 8 and not layout['split']
                HW = H * W
                slot = {}  # destination slot -> (stream, word)
                for i in range(calls):
                    for j in range(4):
                        if destinations[4 * i + j] >= 0:
                            assert destinations[4 * i + j] not in slot
                            slot[destinations[4 * i + j]] = (streams[i], j)
                for c in range(HW):
                    assert slot[c] == tuple(layout['grow'][c]), (H, W, A, c)
                    assert slot[HW + c] == tuple(layout['spread'][c]), (H, W, A, c)
                for a in range(A):
                    for j in range(4):
                        assert slot[2 * HW + 4 * a + j] == tuple(layout['agent'][a][j]), (H, W, A, a, j)
                assert len(slot) == 2 * HW + 4 * A
                checked += 1
    assert checked > 500
    params = _lib.WildfireParams()
    params.height, params.width, params.num_agents = 10, 10, 10
    assert lib.frz_wildfire_tile_random_layout(ctypes.byref(params), streams, destinations) < 0  # not a small grid
