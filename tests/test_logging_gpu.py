"""GPU test of the asynchronous CSV logging tap.

tests/golden/logs/<fixture>/<env>.csv were written by the reference's own CSVLogger
(free_range_zoo/utils/logging_handlers.py:36-111, see tests/golden/gen_logs.py) while it rolled out the trajectory of
the .npz fixture of the same name.  Here the engine replays that trajectory (same actions, same injected uniforms) with
``log_directory`` set; its tap must produce the same files byte for byte -- same columns, order and cell formatting.
"""
import glob
import importlib
import os

import pytest
import torch

from free_range_zoo_b200 import presets
from tests import golden_util as G

pytestmark = pytest.mark.gpu

LOGS = os.path.join(G.GOLDEN_DIR, 'logs')
MODULES = {'wildfire': 'wildfire_v0', 'rideshare': 'rideshare_v0', 'cybersecurity': 'cybersecurity_v0'}


@pytest.mark.parametrize('name', sorted(os.listdir(LOGS)))
def test_log_files_match_the_reference_logger(name, tmp_path):
    meta, gold = G.load(name)
    module = importlib.import_module(f'free_range_zoo_b200.envs.{MODULES[meta["domain"]]}')
    config = getattr(presets, meta['preset'])(**meta.get('preset_kwargs', {}))
    directory = str(tmp_path / 'logs')
    env = module.parallel_env(parallel_envs=meta['B'], max_steps=meta['max_steps'], configuration=config,
                              device=torch.device('cuda'), log_directory=directory, **meta.get('env_kwargs', {}))
    env.reset(seed=0)
    raw = env.unwrapped
    for t in range(meta['steps']):
        if meta['domain'] == 'wildfire':
            raw.inject_uniforms(torch.from_numpy(gold['u_field'][t]), torch.from_numpy(gold['u_agent'][t]))
        elif meta['domain'] == 'cybersecurity':
            raw.inject_uniforms(torch.from_numpy(gold['u_network'][t]), torch.from_numpy(gold['u_agent'][t]))
        actions = torch.from_numpy(gold['actions'][t]).cuda()
        env.step({agent: actions[:, i] for i, agent in enumerate(env.agents)})
    raw.flush_logs()
    assert len(os.listdir(directory)) == meta['B']  # one file per environment
    expected = sorted(glob.glob(os.path.join(LOGS, name, '*.csv')))
    assert expected
    for path in expected:
        want = open(path).read()
        got = open(os.path.join(directory, os.path.basename(path))).read()
        if got != want:
            for line, (g, w) in enumerate(zip(got.splitlines(), want.splitlines())):
                assert g == w, f'{name}/{os.path.basename(path)} line {line}:\n got  {g}\n want {w}'
            assert got == want, f'{name}/{os.path.basename(path)}: different number of lines'


def test_logging_does_not_block_the_step_path(tmp_path):
    """The tap only enqueues copies: a burst of steps returns before the rows are written, flush() then drains it, and
    a non-empty directory is refused like the reference's CSVLogger does."""
    from free_range_zoo_b200.envs import wildfire_v0
    directory = str(tmp_path / 'burst')
    env = wildfire_v0.parallel_env(parallel_envs=256, max_steps=50, configuration=presets.wildfire_3x3(),
                                   device=torch.device('cuda'), log_directory=directory)
    env.reset(seed=3)
    raw = env.unwrapped
    for _ in range(3):
        raw.sample_actions(7)
        raw.step_all()
    raw.flush_logs()
    rows = open(os.path.join(directory, '255.csv')).read().splitlines()
    assert len(rows) == 1 + 1 + 3  # header, reset row, three steps
    assert rows[1].split(',')[-1] == 'NULL' and ',-1,' in rows[1]
    with pytest.raises(FileExistsError):
        wildfire_v0.parallel_env(parallel_envs=4, max_steps=5, configuration=presets.wildfire_3x3(),
                                 device=torch.device('cuda'), log_directory=directory).reset(seed=1)
