"""CPU-side checks of the C ABI: the library loads, exports every symbol include/frz.h declares, the ctypes mirrors
have the C layout, and host-detectable errors follow the status / frz_last_error convention (no GPU needed)."""
import ctypes
import os
import subprocess

import pytest

from free_range_zoo_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STRUCTS = {
    'FrzControl': _lib.Control,
    'FrzWildfireParams': _lib.WildfireParams,
    'FrzWildfireBuffers': _lib.WildfireBuffers,
    'FrzCyberParams': _lib.CyberParams,
    'FrzCyberBuffers': _lib.CyberBuffers,
    'FrzRideshareParams': _lib.RideshareParams,
    'FrzRideshareBuffers': _lib.RideshareBuffers,
    'FrzHostStep': _lib.HostStep,
    'FrzGatherArray': _lib.GatherArray,
}


@pytest.fixture(scope='module')
def lib():
    if not os.path.exists(_lib.LIBRARY_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _lib.library()


def test_every_declared_symbol_is_exported(lib):
    declared = _lib.exported_symbols()
    assert len(declared) == 27  # 8 library-wide + 6 per domain + the wildfire tile kernel's random layout
    missing = [name for name in declared if not hasattr(lib, name)]
    assert not missing, missing
    assert lib.frz_version() == _lib.ABI_VERSION == 4


def test_ctypes_mirrors_match_the_c_layout(tmp_path):
    """Compile a probe against include/frz.h with gcc and compare sizeof / offsetof with the ctypes structures."""
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "frz.h"', 'int main(void) {']
    for c_name, mirror in STRUCTS.items():
        lines.append(f'  printf("{c_name} %zu\\n", sizeof({c_name}));')
        for field, _ in mirror._fields_:
            lines.append(f'  printf("{c_name}.{field} %zu\\n", offsetof({c_name}, {field}));')
    lines += ['  return 0;', '}']
    source = tmp_path / 'probe.c'
    source.write_text('\n'.join(lines))
    binary = tmp_path / 'probe'
    subprocess.run(['gcc', '-I', os.path.join(ROOT, 'include'), str(source), '-o', str(binary)], check=True)
    layout = dict(line.split() for line in subprocess.run([str(binary)], capture_output=True, text=True,
                                                          check=True).stdout.splitlines())
    for c_name, mirror in STRUCTS.items():
        assert int(layout[c_name]) == ctypes.sizeof(mirror), c_name
        for field, _ in mirror._fields_:
            assert int(layout[f'{c_name}.{field}']) == getattr(mirror, field).offset, f'{c_name}.{field}'


def test_null_and_shape_errors_are_reported_without_a_gpu(lib):
    params, io = _lib.WildfireParams(), _lib.WildfireBuffers()
    assert lib.frz_wildfire_step(None, None, 4, None) == 1  # FRZ_ERR_NULL
    assert b'NULL' in lib.frz_last_error()
    assert lib.frz_wildfire_step(ctypes.byref(params), ctypes.byref(io), 4, None) == 1  # control / fires are NULL
    cy_params, cy_io = _lib.CyberParams(), _lib.CyberBuffers()
    assert lib.frz_cyber_step(ctypes.byref(cy_params), ctypes.byref(cy_io), 4, None) == 1
    rs_params, rs_io = _lib.RideshareParams(), _lib.RideshareBuffers()
    assert lib.frz_rideshare_step(ctypes.byref(rs_params), ctypes.byref(rs_io), 4, None) == 1
    # the host-buffer step checks its own block too
    assert lib.frz_wildfire_step_host(None, None, 4, None, None) == 1
    assert lib.frz_cyber_step_host(ctypes.byref(cy_params), ctypes.byref(cy_io), 4, None, None) == 1
    assert lib.frz_rideshare_step_host(ctypes.byref(rs_params), ctypes.byref(rs_io), 4, None, None) == 1
    # non-NULL pointers but an unsupported shape -> FRZ_ERR_SHAPE before anything is launched
    dummy = ctypes.create_string_buffer(64)
    address = ctypes.addressof(dummy)
    io.control, io.fires, io.actions, io.cell_agents, io.range_mask = (address,) * 5
    params.height, params.width, params.num_agents = 40, 40, 3  # 1600 cells > FRZ_MAX_CELLS
    assert lib.frz_wildfire_step(ctypes.byref(params), ctypes.byref(io), 4, None) == 2  # FRZ_ERR_SHAPE
    assert b'unsupported shape' in lib.frz_last_error()
    assert lib.frz_wildfire_step(ctypes.byref(params), ctypes.byref(io), 0, None) == 2
    with pytest.raises(RuntimeError, match='status 2'):
        _lib.check(2, 'probe')


def test_there_is_no_cpu_fallback():
    import torch

    from free_range_zoo_b200 import presets
    from free_range_zoo_b200.envs import cybersecurity_v0, rideshare_v0, wildfire_v0
    for module, preset in ((wildfire_v0, presets.wildfire_profile), (rideshare_v0, presets.rideshare_profile),
                           (cybersecurity_v0, presets.cyber_profile)):
        with pytest.raises(RuntimeError, match='no CPU fallback'):
            module.parallel_env(parallel_envs=2, configuration=preset(), device=torch.device('cpu'))


def test_product_code_never_imports_the_oracle():
    offenders = []
    for folder, _, files in os.walk(os.path.join(ROOT, 'free_range_zoo_b200')):
        for name in files:
            if name.endswith('.py'):
                text = open(os.path.join(folder, name)).read()
                if 'import oracle' in text or 'from oracle' in text:
                    offenders.append(os.path.join(folder, name))
    assert not offenders, offenders


def test_host_step_slices(lib):
    """frz_host_slices: what frz_<domain>_step_host does with a batch (pure host logic, no GPU)."""
    assert _lib.host_slices(65536, 1) == [0, 65536]
    assert _lib.host_slices(65536, 5) == [0, 8192, 22528, 36864, 51200, 65536]  # first slice half as long
    assert _lib.host_slices(1000, 4) == [0, 1000]  # too small to cut: one slice
    assert _lib.host_slices(3000, 2) == [0, 1024, 3000]
    for B in (1, 1023, 1024, 1025, 4100, 70000, 524288, 4194304):
        for chunks in range(1, _lib.MAX_CHUNKS + 1):
            bounds = _lib.host_slices(B, chunks)
            assert bounds[0] == 0 and bounds[-1] == B and len(bounds) - 1 <= chunks
            assert all(a < b for a, b in zip(bounds, bounds[1:]))
            assert all(b % 1024 == 0 for b in bounds[1:-1])
            sizes = [b - a for a, b in zip(bounds, bounds[1:])]
            if len(sizes) > 2:
                assert sizes[0] <= max(sizes[1:-1])
    with pytest.raises(RuntimeError):
        _lib.host_slices(0, 2)
    with pytest.raises(RuntimeError):
        _lib.host_slices(100, _lib.MAX_CHUNKS + 1)


def test_buffer_bytes_matches_the_host_allocations(lib):
    """frz_<domain>_buffer_bytes (pure host logic): the sizes a C caller needs equal what the Python host allocates --
    checked against the shapes / dtypes documented in include/frz.h for one configuration per domain."""
    import torch

    from free_range_zoo_b200 import presets
    from free_range_zoo_b200.envs.cybersecurity.env import cybersecurity
    from free_range_zoo_b200.envs.rideshare.env import rideshare
    from free_range_zoo_b200.envs.wildfire.env import wildfire
    B = 37
    params = wildfire.flatten_configuration(presets.wildfire_large(), 100, False)[0]
    sizes = {name: lib.frz_wildfire_buffer_bytes(ctypes.byref(params), B, name.encode())
             for name, _ in _lib.WildfireBuffers._fields_ if name not in ('mask_stride', 'mask_words')}
    assert sizes['fires'] == sizes['intensity'] == sizes['init_fuel'] == 4 * B * 100
    assert sizes['action_mask'] == B * 10 * 100 and sizes['task_obs'] == 16 * B * 100 and sizes['self_obs'] == 16 * B * 10
    assert sizes['actions'] == 8 * B * 10 and sizes['terminated'] == B and sizes['control'] == 64
    assert sizes['range_mask'] == 4 * 10 * 3 * 4 and sizes['cell_agents'] == 4 * 3 * 100
    assert all(value > 0 for value in sizes.values()), sizes
    assert lib.frz_wildfire_buffer_bytes(ctypes.byref(params), B, b'no_such_field') == -1
    assert b'unknown buffer field' in lib.frz_last_error()
    assert lib.frz_wildfire_buffer_bytes(None, B, b'fires') == -1

    cy_params = cybersecurity.flatten_configuration(presets.cyber_c3(), 100, False)[0]
    cy = {name: lib.frz_cyber_buffer_bytes(ctypes.byref(cy_params), B, name.encode()) for name, _ in _lib.CyberBuffers._fields_}
    assert cy['network_state'] == 4 * B * 3 and cy['presence'] == B * 4 and cy['defender_self'] == 12 * B * 2
    assert cy['task_obs'] == 8 * B * 3 and cy['score_lut'] == 4 * 16 and min(cy.values()) > 0

    rs_params = rideshare.flatten_configuration(presets.rideshare_c2(), 100, 0)[0]
    rs = {name: lib.frz_rideshare_buffer_bytes(ctypes.byref(rs_params), B, name.encode())
          for name, _ in _lib.RideshareBuffers._fields_}
    K = rs_params.capacity
    assert rs['passengers'] == 4 * B * K * 11 and rs['task_mask'] == B * 4 * K and rs['task_obs'] == 32 * B * K
    assert rs['schedule'] == 4 * 32 * 7 and min(rs.values()) > 0


def test_host_pipeline_handle_argument_checks(lib):
    assert lib.frz_host_pipeline_create(None) == 1  # FRZ_ERR_NULL
    assert lib.frz_host_pipeline_destroy(None) == 0
    assert lib.frz_control_restore(None, 1, 2, None) == 1
