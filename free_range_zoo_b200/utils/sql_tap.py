"""SQLite sink of the asynchronous logging tap (the SQL half).

The reference's ``SQLLogger`` (free_range_zoo/utils/logging_handlers.py:116-241) writes, through SQLAlchemy models
(free_range_zoo/utils/sql_logging.py:12-117), one ``simulation`` row per ``reset``, one ``environment`` row per parallel
environment, one ``agent`` row per (agent, environment), and per logged step one ``environment_timestep`` row, one
``<domain>_environment_log`` row and -- except right after a reset -- one ``agent_log`` row per agent, all synchronously
on the step path with a ``session.flush()`` per row.  SQLAlchemy is not a dependency of this engine: the same tables
(same names, columns, types, keys -- the DDL ``Base.metadata.create_all`` emits for those models) are created and
filled with the standard library's ``sqlite3``, on the tap's writer thread, one transaction per logged step.

Cells have the reference's formats: ``str(tensor.tolist())`` for state and mapping cells, ``int(...)`` (truncation)
for ``agent_log.reward`` / ``action_field`` / ``task_field`` (logging_handlers.py:226-234), today's date for
``simulation.timestamp``.  Only ``sqlite://`` URLs are served (``postgresql://`` needs a driver this image does not
have).  One deliberate difference: the reference indexes rideshare's FLAT passenger table with the environment index
(``state.passengers[env_idx]``, logging_handlers.py:200 -- a single passenger row, or an IndexError); here
``rideshare_environment_log.passengers`` holds the environment's own rows, the evident intent.
"""
from __future__ import annotations

import datetime
import sqlite3
from typing import Any, Dict, List, Optional, Sequence

# (table, [(column, declaration)], [table-level constraints]) -- what create_all emits for sql_logging.py:12-104
SCHEMA = [
    ('simulation', [('id', 'INTEGER NOT NULL'), ('name', 'TEXT NOT NULL'), ('description', 'TEXT'),
                    ('timestamp', 'DATE NOT NULL')], ['PRIMARY KEY (id)']),
    ('environment', [('id', 'INTEGER NOT NULL'), ('simulation_id', 'INTEGER NOT NULL'), ('simulation_index', 'INTEGER')],
     ['PRIMARY KEY (id)', 'FOREIGN KEY(simulation_id) REFERENCES simulation (id)']),
    ('agent', [('id', 'INTEGER NOT NULL'), ('environment_id', 'INTEGER NOT NULL'), ('name', 'TEXT NOT NULL')],
     ['PRIMARY KEY (id)', 'FOREIGN KEY(environment_id) REFERENCES environment (id)']),
    ('environment_timestep', [('environment_id', 'INTEGER NOT NULL'), ('id', 'INTEGER NOT NULL'), ('timestep', 'INTEGER')],
     ['PRIMARY KEY (id)', 'FOREIGN KEY(environment_id) REFERENCES environment (id)']),
    ('wildfire_environment_log',
     [('id', 'INTEGER NOT NULL'), ('simulation_timestep_id', 'INTEGER NOT NULL'), ('fires', 'TEXT'), ('intensity', 'TEXT'),
      ('fuel', 'TEXT'), ('suppressants', 'TEXT'), ('capacity', 'TEXT'), ('equipment', 'TEXT'), ('agents', 'TEXT')],
     ['PRIMARY KEY (id)', 'FOREIGN KEY(simulation_timestep_id) REFERENCES environment_timestep (id)']),
    ('rideshare_environment_log',
     [('id', 'INTEGER NOT NULL'), ('simulation_timestep_id', 'INTEGER NOT NULL'), ('agents', 'TEXT'), ('passengers', 'TEXT')],
     ['PRIMARY KEY (id)', 'FOREIGN KEY(simulation_timestep_id) REFERENCES environment_timestep (id)']),
    ('cybersecurity_environment_log',
     [('id', 'INTEGER NOT NULL'), ('simulation_timestep_id', 'INTEGER NOT NULL'), ('network_state', 'TEXT'),
      ('location', 'TEXT'), ('presence', 'TEXT'), ('adj_matrix', 'TEXT')],
     ['PRIMARY KEY (id)', 'FOREIGN KEY(simulation_timestep_id) REFERENCES environment_timestep (id)']),
    ('agent_log',
     [('id', 'INTEGER NOT NULL'), ('simulation_timestep_id', 'INTEGER NOT NULL'), ('agent_id', 'INTEGER NOT NULL'),
      ('reward', 'INTEGER'), ('action_field', 'INTEGER'), ('task_field', 'INTEGER'), ('action_map', 'TEXT'),
      ('observation_map', 'TEXT')],
     ['PRIMARY KEY (id)', 'FOREIGN KEY(simulation_timestep_id) REFERENCES environment_timestep (id)',
      'FOREIGN KEY(agent_id) REFERENCES agent (id)']),
]
DOMAIN_TABLES = {
    'wildfire': ('wildfire_environment_log', ('fires', 'intensity', 'fuel', 'suppressants', 'capacity', 'equipment', 'agents')),
    'rideshare': ('rideshare_environment_log', ('agents', 'passengers')),
    'cybersecurity': ('cybersecurity_environment_log', ('network_state', 'location', 'presence', 'adj_matrix')),
}


def sqlite_path(connection_string: str) -> str:
    """``sqlite:///relative.db`` / ``sqlite:////absolute.db`` / ``sqlite://`` (in memory), as SQLAlchemy reads them."""
    if not connection_string.startswith('sqlite://'):
        raise NotImplementedError(f'only sqlite:// connection strings are served (got {connection_string!r})')
    rest = connection_string[len('sqlite://'):]
    if rest in ('', '/'):
        return ':memory:'
    if not rest.startswith('/'):
        raise ValueError(f'malformed sqlite URL {connection_string!r}: expected sqlite:///<path>')
    return rest[1:]


class SqliteSink:
    """Writes the reference's SQL log tables with ``sqlite3``.  Used from ONE thread (the tap's writer thread)."""

    def __init__(self, connection_string: str, domain: str, parallel_envs: int):
        self.path = sqlite_path(connection_string)
        self.domain = domain.split('_')[0]  # metadata name "wildfire_v0" -> "wildfire" (logging_handlers.py:184)
        if self.domain not in DOMAIN_TABLES:
            raise NotImplementedError(f'Environment {domain} does not have an implemented log_environment function.')
        self.parallel_envs = parallel_envs
        self._connection: Optional[sqlite3.Connection] = None
        self._environment_ids: List[int] = []
        self._agent_ids: Dict[Any, int] = {}

    def _db(self) -> sqlite3.Connection:
        if self._connection is None:  # opened lazily, on the thread that uses it
            self._connection = sqlite3.connect(self.path)
            for table, columns, constraints in SCHEMA:
                body = ', '.join([f'{name} {declaration}' for name, declaration in columns] + constraints)
                self._connection.execute(f'CREATE TABLE IF NOT EXISTS {table} ({body})')
            self._connection.commit()
        return self._connection

    def reset(self, label: Optional[str], description: Optional[str], agents: Sequence[str]) -> None:
        """SQLLogger.reset (logging_handlers.py:131-158): a new simulation, its environments and their agents."""
        db = self._db()
        with db:
            simulation = db.execute('INSERT INTO simulation (name, description, timestamp) VALUES (?, ?, ?)',
                                    (label or 'simulation', description, datetime.date.today().isoformat())).lastrowid
            self._environment_ids = [
                db.execute('INSERT INTO environment (simulation_id, simulation_index) VALUES (?, ?)', (simulation, i)).lastrowid
                for i in range(self.parallel_envs)
            ]
            self._agent_ids = {}
            for agent in agents:
                for environment in self._environment_ids:
                    self._agent_ids[(agent, environment)] = db.execute(
                        'INSERT INTO agent (name, environment_id) VALUES (?, ?)', (agent, environment)).lastrowid

    def write(self, record: Dict[str, Any], reset: bool) -> None:
        """SQLLogger.log_environment (logging_handlers.py:160-241).  ``record``: ``timestep`` int per environment,
        ``state`` {column: cell per environment} and ``agents`` {name: {reward, action_field, task_field, action_map,
        observation_map: one value per environment}}."""
        if not self._environment_ids:
            raise RuntimeError('SQLLogger: reset() must be called before logging. _env_ids is None.')
        table, columns = DOMAIN_TABLES[self.domain]
        db = self._db()
        with db:  # one transaction per logged step
            for index, environment in enumerate(self._environment_ids):
                timestep = db.execute('INSERT INTO environment_timestep (environment_id, timestep) VALUES (?, ?)',
                                      (environment, int(record['timestep'][index]))).lastrowid
                cells = [record['state'][column][index] for column in columns]
                db.execute(f'INSERT INTO {table} (simulation_timestep_id, {", ".join(columns)}) '
                           f'VALUES ({", ".join("?" * (len(columns) + 1))})', [timestep] + cells)
                if reset:
                    continue
                for agent, fields in record['agents'].items():
                    agent_id = self._agent_ids.get((agent, environment))
                    if agent_id is None:
                        continue
                    db.execute(
                        'INSERT INTO agent_log (simulation_timestep_id, agent_id, reward, action_field, task_field, '
                        'action_map, observation_map) VALUES (?, ?, ?, ?, ?, ?, ?)',
                        (timestep, agent_id, int(fields['reward'][index]), int(fields['action_field'][index]),
                         int(fields['task_field'][index]), fields['action_map'][index], fields['observation_map'][index]))

    def close(self) -> None:
        if self._connection is not None:
            self._connection.close()
            self._connection = None
