"""Asynchronous logging tap: CSV files, or the SQL tables of utils/sql_tap.py.

The reference logs synchronously on the step path: after every environment step ``CSVLogger.log_environment``
(free_range_zoo/utils/logging_handlers.py:52-111) pulls every state tensor to the host with ``.tolist()``, builds a
pandas frame and appends one row to one file per environment -- O(B) Python work and several device synchronisations
per step.  Here the step path only ENQUEUES work:

  main stream   device-to-device snapshot of the live buffers into a ring slot (the kernels update state in place, so
                the next step may run as soon as this copy is queued), then an event;
  side stream   waits for that event, copies the slot to pinned host memory, records a second event;
  writer thread waits for the second event, formats the rows and appends them to ``<log_directory>/<env>.csv``.

The host never waits for the GPU on the step path unless the ring (``depth`` slots) is full.  The files are the
reference's: same columns, same order, same cell formatting (``str(tensor.tolist())`` cells, ``NULL`` for missing
values, minimal quoting), one file per environment, header rewritten by every ``reset`` -- pinned by
tests/golden/logs/*, which were written by the reference's own CSVLogger.
"""
from __future__ import annotations

import csv
import os
import queue
import threading
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch


class _Slot:

    def __init__(self, live: Dict[str, torch.Tensor]):
        self.device = {name: torch.empty_like(tensor) for name, tensor in live.items()}
        self.host = {name: torch.empty(tensor.shape, dtype=tensor.dtype, pin_memory=True) for name, tensor in live.items()}
        self.free = threading.Event()
        self.free.set()


class LoggingTap:
    """Ring-buffered, stream-ordered CSV logger for one environment object."""

    def __init__(self, log_directory: str, parallel_envs: int, device: torch.device,
                 snapshot: Callable[[], Dict[str, torch.Tensor]],
                 columns: Callable[[Dict[str, np.ndarray], bool], Dict[str, Sequence]], depth: int = 4,
                 override_initialization_check: bool = False, sink=None, agents: Sequence[str] = ()):
        """
        Args:
            log_directory: directory that receives ``<env index>.csv`` (created; must be empty unless overridden,
                like the reference's CSVLogger, logging_handlers.py:41-49)
            snapshot: returns the live device tensors to record, called on the step path (no copies, no syncs)
            columns: turns the host copy of one snapshot into ordered CSV columns, each a length-B sequence; runs on the
                writer thread
            depth: ring slots; ``capture`` blocks only when all of them are still being written
            sink: optional row sink replacing the CSV files (utils/sql_tap.py::SqliteSink); ``columns`` then returns the
                sink's record instead of CSV columns, ``agents`` are the names registered by every reset
        """
        self._sink, self._agents = sink, tuple(agents)
        if sink is None:
            if not override_initialization_check and os.path.exists(log_directory) and os.listdir(log_directory):
                raise FileExistsError('The logging output directory already exists. Set override_initialization_check or rename.')
            os.makedirs(log_directory, exist_ok=True)
        self.log_directory, self.parallel_envs, self.device = log_directory, parallel_envs, device
        self._snapshot, self._columns, self._depth = snapshot, columns, depth
        self._slots: List[_Slot] = []
        self._next = 0
        self._copy_stream = torch.cuda.Stream(device)
        self._queue: 'queue.Queue' = queue.Queue()
        self._error: Optional[BaseException] = None
        self._writer = threading.Thread(target=self._write_loop, name='frz-logging-tap', daemon=True)
        self._writer.start()

    # ------------------------------------------------------------------------------------------ step path

    def capture(self, reset: bool, description: Optional[str], label: Optional[str] = None) -> None:
        """Queue one log row per environment for the current device state."""
        if self._error is not None:
            raise RuntimeError('the logging tap failed') from self._error
        live = self._snapshot()
        index = self._next % self._depth
        if index == len(self._slots):  # the ring is allocated on first use, shaped like the live buffers
            self._slots.append(_Slot(live))
        slot = self._slots[index]
        self._next += 1
        slot.free.wait()  # back-pressure: only when the writer is `depth` steps behind
        slot.free.clear()
        main = torch.cuda.current_stream(self.device)
        for name, tensor in live.items():
            slot.device[name].copy_(tensor, non_blocking=True)
        staged = torch.cuda.Event()
        staged.record(main)
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(staged)
            for name in live:
                slot.host[name].copy_(slot.device[name], non_blocking=True)
            landed = torch.cuda.Event()
            landed.record(self._copy_stream)
        self._queue.put((slot, landed, reset, description, label))

    def flush(self) -> None:
        """Block until every queued row is on disk."""
        self._queue.join()
        if self._error is not None:
            raise RuntimeError('the logging tap failed') from self._error

    def close(self) -> None:
        self.flush()
        self._queue.put(None)
        self._writer.join()
        if self._sink is not None:
            self._sink.close()

    # ------------------------------------------------------------------------------------------ writer thread

    @staticmethod
    def _cell(value) -> str:
        if value is None:
            return 'NULL'
        if isinstance(value, (np.floating, float)):
            return repr(float(value))
        if isinstance(value, (np.bool_, bool)):
            return str(bool(value))
        if isinstance(value, np.integer):
            return str(int(value))
        return str(value)

    def _write_loop(self) -> None:
        while True:
            item = self._queue.get()
            if item is None:
                self._queue.task_done()
                return
            slot, landed, reset, description, label = item
            try:
                landed.synchronize()
                host = {name: tensor.numpy() for name, tensor in slot.host.items()}
                if self._sink is not None:
                    if reset:
                        self._sink.reset(label, description, self._agents)
                    self._sink.write(self._columns(host, reset), reset)
                    continue
                columns = dict(self._columns(host, reset))
                columns['description'] = [description] * self.parallel_envs
                names = list(columns)
                for env in range(self.parallel_envs):
                    path = os.path.join(self.log_directory, f'{env}.csv')
                    with open(path, 'w' if reset else 'a', newline='') as handle:
                        writer = csv.writer(handle, quoting=csv.QUOTE_MINIMAL, lineterminator='\n')
                        if reset:
                            writer.writerow(names)
                        writer.writerow([self._cell(columns[name][env]) for name in names])
            except BaseException as error:  # surfaced by the next capture / flush
                self._error = error
            finally:
                slot.free.set()
                self._queue.task_done()


def nested(array: np.ndarray) -> str:
    """``str(tensor.tolist())`` of one environment's slice -- the cell format of the reference's ``to_dataframe``."""
    return str(array.tolist())


def index_list(mask_row: np.ndarray) -> str:
    """``str(mapping.tolist())`` for a jagged index mapping given as a boolean mask over the task slots."""
    return str(np.nonzero(mask_row)[0].tolist())
