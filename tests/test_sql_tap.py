"""SQL half of the logging tap (utils/sql_tap.py): the tables of the reference's SQLAlchemy models
(free_range_zoo/utils/sql_logging.py:12-104) written with sqlite3, rows as SQLLogger writes them
(free_range_zoo/utils/logging_handlers.py:116-241)."""
import csv
import os
import sqlite3

import numpy as np
import pytest

from free_range_zoo_b200.utils.sql_tap import SqliteSink, sqlite_path

# table -> columns in declaration order, transcribed from the reference's models (sql_logging.py)
REFERENCE_TABLES = {
    'simulation': ['id', 'name', 'description', 'timestamp'],  # :12-20
    'environment': ['id', 'simulation_id', 'simulation_index'],  # :23-32
    'agent': ['id', 'environment_id', 'name'],  # :35-43
    'environment_timestep': ['environment_id', 'id', 'timestep'],  # :46-58
    'wildfire_environment_log': ['id', 'simulation_timestep_id', 'fires', 'intensity', 'fuel', 'suppressants', 'capacity',
                                 'equipment', 'agents'],  # :61-74
    'rideshare_environment_log': ['id', 'simulation_timestep_id', 'agents', 'passengers'],  # :77-84
    'cybersecurity_environment_log': ['id', 'simulation_timestep_id', 'network_state', 'location', 'presence',
                                      'adj_matrix'],  # :87-96
    'agent_log': ['id', 'simulation_timestep_id', 'agent_id', 'reward', 'action_field', 'task_field', 'action_map',
                  'observation_map'],  # :99-110
}


def test_sqlite_urls():
    assert sqlite_path('sqlite:///logs/run.db') == 'logs/run.db'
    assert sqlite_path('sqlite:////tmp/run.db') == '/tmp/run.db'
    assert sqlite_path('sqlite://') == ':memory:'
    with pytest.raises(NotImplementedError):
        sqlite_path('postgresql://user@host/db')
    with pytest.raises(ValueError):
        sqlite_path('sqlite://host/db')


def test_schema_and_rows_follow_the_reference_logger(tmp_path):
    path = tmp_path / 'log.db'
    sink = SqliteSink(f'sqlite:///{path}', 'wildfire_v0', parallel_envs=2)
    agents = ('firefighter_1', 'firefighter_2')
    with pytest.raises(RuntimeError):  # logging_handlers.py:172-173
        sink.write({}, reset=False)
    sink.reset('run', 'first', agents)

    def record(step):
        state = {name: [f'{name}{step}e0', f'{name}{step}e1']
                 for name in ('fires', 'intensity', 'fuel', 'suppressants', 'capacity', 'equipment', 'agents')}
        per_agent = {agent: dict(reward=np.array([1.9, -2.7], np.float32), action_field=np.array([0, -1]),
                                 task_field=np.array([3, 0]), action_map=['[0, 1]', '[]'], observation_map=['[0, 1]', '[0]'])
                     for agent in agents}
        return dict(timestep=np.array([step, step]), state=state, agents=per_agent)

    sink.write(record(0), reset=True)
    sink.write(record(1), reset=False)
    sink.reset(None, None, agents)  # a second episode: new simulation, new environments, new agents
    sink.write(record(0), reset=True)
    sink.close()

    db = sqlite3.connect(path)
    tables = {name for (name, ) in db.execute("SELECT name FROM sqlite_master WHERE type='table'")}
    assert tables == set(REFERENCE_TABLES)
    for table, columns in REFERENCE_TABLES.items():
        assert [row[1] for row in db.execute(f'PRAGMA table_info({table})')] == columns, table
    count = lambda table: db.execute(f'SELECT COUNT(*) FROM {table}').fetchone()[0]
    assert count('simulation') == 2 and count('environment') == 4 and count('agent') == 8
    assert count('environment_timestep') == 6 and count('wildfire_environment_log') == 6
    assert count('agent_log') == 4  # one per (agent, environment) of the one non-reset step
    assert db.execute('SELECT name, description FROM simulation ORDER BY id').fetchall() == [('run', 'first'),
                                                                                           ('simulation', None)]
    # agents are registered agent-major (logging_handlers.py:152-158): ids 1, 2 = firefighter_1 in environments 1, 2
    assert db.execute('SELECT name, environment_id FROM agent WHERE id <= 4 ORDER BY id').fetchall() == [
        ('firefighter_1', 1), ('firefighter_1', 2), ('firefighter_2', 1), ('firefighter_2', 2)]
    # rewards are truncated by int() like the reference (:229)
    assert db.execute('SELECT reward, action_field, task_field, action_map FROM agent_log ORDER BY id').fetchall()[:2] == [
        (1, 0, 3, '[0, 1]'), (1, 0, 3, '[0, 1]')]
    assert db.execute('SELECT reward, action_field FROM agent_log WHERE agent_id IN (2, 4) ORDER BY id').fetchall() == [
        (-2, -1), (-2, -1)]
    assert db.execute('SELECT fires, agents FROM wildfire_environment_log ORDER BY id').fetchall()[2] == ('fires1e0', 'agents1e0')


@pytest.mark.gpu
@pytest.mark.parametrize('domain,preset,kwargs', [
    ('wildfire', 'wildfire_3x3', {}),
    ('rideshare', 'rideshare_c2', {}),
    ('cybersecurity', 'cyber_c3', dict(show_bad_actions=False, partially_observable=True)),
])
def test_sql_rows_equal_the_csv_rows_of_the_same_rollout(domain, preset, kwargs, tmp_path):
    """The CSV files are pinned to the reference's CSVLogger (tests/test_logging_gpu.py); the SQL tables must carry the
    same cells for the same seeded rollout."""
    import importlib

    import torch

    from free_range_zoo_b200 import presets
    module = importlib.import_module(f'free_range_zoo_b200.envs.{domain}_v0')
    B, steps = 3, 4
    database, directory = tmp_path / 'log.db', tmp_path / 'csv'

    def rollout(log_directory):
        env = module.parallel_env(parallel_envs=B, max_steps=20, configuration=getattr(presets, preset)(),
                                  device=torch.device('cuda'), log_directory=log_directory, **kwargs)
        env.reset(seed=5, options={'log_label': 'parity', 'log_description': 'same rollout'})
        raw = env.unwrapped
        for _ in range(steps):
            raw.sample_actions(17)
            raw.step_all()
        raw.flush_logs()
        return raw

    raw = rollout(f'sqlite:///{database}')
    rollout(str(directory))
    db = sqlite3.connect(database)
    assert db.execute('SELECT name, description FROM simulation').fetchall() == [('parity', 'same rollout')]
    table = f'{domain}_environment_log'
    state_columns = [row[1] for row in db.execute(f'PRAGMA table_info({table})')][2:]
    for env_index in range(B):
        rows = list(csv.DictReader(open(os.path.join(directory, f'{env_index}.csv'))))
        assert len(rows) == steps + 1
        logged = db.execute(
            f'SELECT t.id, t.timestep, {", ".join("l." + c for c in state_columns)} FROM environment_timestep t '
            f'JOIN environment e ON e.id = t.environment_id JOIN {table} l ON l.simulation_timestep_id = t.id '
            'WHERE e.simulation_index = ? ORDER BY t.id', (env_index, )).fetchall()
        assert len(logged) == steps + 1
        for step, (csv_row, sql_row) in enumerate(zip(rows, logged)):
            assert sql_row[1] == step  # num_moves; the CSV writes -1 on the reset row (logging_handlers.py:90 vs :181)
            for column, cell in zip(state_columns, sql_row[2:]):
                assert cell == csv_row[column], (column, step)
            agent_rows = db.execute(
                'SELECT a.name, g.reward, g.action_field, g.task_field, g.action_map, g.observation_map FROM agent_log g '
                'JOIN agent a ON a.id = g.agent_id WHERE g.simulation_timestep_id = ? ORDER BY g.id', (sql_row[0], )).fetchall()
            if step == 0:
                assert agent_rows == []  # no agent rows right after a reset (:221)
                continue
            assert [row[0] for row in agent_rows] == list(raw.possible_agents)
            for name, reward, action_field, task_field, action_map, observation_map in agent_rows:
                assert reward == int(float(csv_row[f'{name}_rewards']))
                assert [task_field, action_field] == eval(csv_row[f'{name}_action'])
                assert action_map == csv_row[f'{name}_action_map']
                assert observation_map == csv_row[f'{name}_observation_map']
