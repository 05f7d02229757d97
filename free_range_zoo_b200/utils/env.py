"""AEC runtime shell of the B200 engine.

``BatchedAECEnv`` keeps the surface of the reference base class (free_range_zoo/utils/env.py:18-359): the same
constructor keywords, ``reset / reset_batches / step / observe / state / action_space / observation_space``, the
five domain hooks (``step_environment``, ``update_actions``, ``update_observations``, ``action_space``,
``observation_space``) and the public attributes (``agents``, ``rewards``, ``terminations``, ``truncations``,
``infos``, ``num_moves``, ``environment_task_count``, ``agent_task_count``, ``finished`` ...).

What changed underneath: the whole cycle the reference runs on the last agent of an AEC round --
``step_environment -> num_moves += 1 -> truncation -> accumulate rewards -> update_observations -> update_actions``
(utils/env.py:220-237) -- is ONE fused CUDA launch through the C ABI of ``include/frz.h``.  All per-step outputs
live in persistent device buffers that the kernel updates in place; the dictionaries handed to the caller are views
of those buffers, so there is no host round trip, no ``.tolist()``, no allocation on the step path.  The reference's
host-side "is every environment done?" test (utils/env.py:212, a device synchronisation per agent call) is evaluated
on the device from flags the previous launch published.

NOTE on aliasing: because buffers are updated in place, tensors returned by ``step`` are overwritten by the next
``step``.  Pass ``detach_outputs=True`` to get fresh copies (the reference's behaviour) at the cost of one copy each.
"""
from __future__ import annotations

import ctypes
import math
from abc import ABC, abstractmethod
from typing import Any, Dict, List, Optional, Tuple

import torch

from free_range_zoo_b200 import _lib
from free_range_zoo_b200.utils.configuration import Configuration
from free_range_zoo_b200.utils.containers import ObservationDict
from free_range_zoo_b200.utils.selector import AgentSelector


# most slices ``step_host`` cuts a batch into by default (measured on B200, profiles/README.md)
HOST_STEP_MAX_DEFAULT_CHUNKS = 5


class _DeviceBound:
    """libfrz entry points bound to one CUDA device: every call of the C ABI works on the CURRENT device
    (include/frz.h), so calls made while another device is current switch to the environment's device for their
    duration.  When it already is current (the normal case) the call goes straight through."""

    def __init__(self, library, device: torch.device):
        self._library, self._index = library, device.index

    def __getattr__(self, name):
        entry, index = getattr(self._library, name), self._index

        def call(*args):
            if torch.cuda.current_device() == index:
                return entry(*args)
            with torch.cuda.device(index):
                return entry(*args)

        self.__dict__[name] = call
        return call


def _seed_to_u64(seed) -> int:
    """Fold whatever ``reset(seed=...)`` received (None / int / list / tensor) into one 64-bit Philox key."""
    if seed is None:
        return int(torch.randint(0, 2**62, (1, )).item())
    if isinstance(seed, torch.Tensor):
        seed = seed.flatten().tolist()
    if isinstance(seed, (list, tuple)):
        folded = 0x9E3779B97F4A7C15
        for value in seed:
            folded = ((folded ^ (int(value) & 0xFFFFFFFFFFFFFFFF)) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
        return folded
    return int(seed) & 0xFFFFFFFFFFFFFFFF


class BatchedAECEnv(ABC):
    """Batched agent-environment-cycle environment whose step runs as one CUDA kernel."""

    metadata: Dict[str, Any] = {}

    def __init__(
        self,
        *args,
        configuration: Configuration = None,
        max_steps: int = 1,
        parallel_envs: int = 1,
        device: torch.device = torch.device('cuda'),
        render_mode: str | None = None,
        log_directory: str = None,
        single_seeding: bool = False,
        buffer_size: int = 0,
        override_initialization_check: bool = False,
        env_offset: int = 0,
        detach_outputs: bool = False,
        **kwargs,
    ):
        """
        Args (same meaning as the reference, utils/env.py:21-48):
            configuration: the domain's configuration structure
            max_steps: truncation horizon
            parallel_envs: number of environments stepped together
            device: must be a CUDA device -- there is no CPU path
            render_mode: accepted; rendering is a host-side subsystem outside this engine
            log_directory / override_initialization_check: a directory turns on the asynchronous logging tap
                (utils/logging_tap.py; same files as the reference's CSVLogger), a ``sqlite://`` URL its SQL form
                (utils/sql_tap.py; the reference's SQLLogger tables); other connection strings are refused
            single_seeding / buffer_size: accepted and ignored -- randomness is counter-based Philox generated inside
                the step kernel, keyed by (seed, global env index, step), so there are no generator states to juggle
            env_offset: global index of this shard's first environment (multi-GPU sharding keeps trajectories invariant)
            detach_outputs: return copies instead of views of the in-place updated device buffers
        """
        device = torch.device(device)
        if device.type != 'cuda':
            raise RuntimeError(f'free_range_zoo_b200 runs on CUDA devices only (got device={device}); '
                               'there is no CPU fallback -- use the reference implementation on CPU.')
        if device.index is None:
            device = torch.device('cuda', torch.cuda.current_device())
        if log_directory is not None and str(log_directory).startswith(('postgres', 'mysql')):
            raise NotImplementedError('the asynchronous logging tap writes CSV files or sqlite:// databases; '
                                      f'no driver for {log_directory!r} is available')
        self.parallel_envs = int(parallel_envs)
        self.max_steps = max_steps
        self.device = device
        self.render_mode = render_mode
        self.log_directory = log_directory
        self.override_initialization_check = override_initialization_check
        self.single_seeding = single_seeding
        self.log_description = None
        self.logger = None
        self.env_offset = int(env_offset)
        self.detach_outputs = detach_outputs

        if configuration is not None:
            self.config = configuration
            # hoist the sub-configurations as attributes, like the reference (utils/env.py:58-63)
            for key, value in vars(configuration).items():
                if hasattr(value, 'validate') and not isinstance(value, torch.Tensor):
                    setattr(self, key, value)

        self._lib = _DeviceBound(_lib.library(), self.device)
        self._tap = None  # asynchronous CSV logging (utils/logging_tap.py), created by the first reset
        self._control = torch.zeros(8, dtype=torch.int64, device=self.device)  # FrzControl, 64 bytes
        self._seed_value = None
        self._graph = None

    def __del__(self):
        handle = getattr(self, '_host_events', None)
        if handle is not None:
            try:
                self._lib.frz_host_pipeline_destroy(handle)
            except Exception:
                pass
            self._host_events = None

    # ------------------------------------------------------------------------------------------ random stream

    def generator_state_dict(self) -> Dict[str, int]:
        """The whole random state of the engine (reference: ``RandomGenerator.state_dict``, one pickled
        ``torch.Generator`` state per environment, utils/random_generator.py:148-162): every draw is
        Philox(seed; global environment index, step, event), so ``seed`` and the step counter say it all.
        Synchronises (reads the device control block)."""
        block = self.control_block()
        return {'seed': block['seed'], 'step': block['step'], 'env_offset': self.env_offset}

    def load_generator_state_dict(self, state: Dict[str, int]) -> None:
        """Continue the random stream of a checkpoint (reference: ``RandomGenerator.load_state_dict``,
        utils/random_generator.py:164-176).  Restore the state tensors (``env.state().load(...)`` /
        ``reset(options={'initial_state': ...})``) and call ``update_actions()`` to re-publish masks and flags."""
        self._seed_value = int(state['seed'])
        _lib.check(self._lib.frz_control_restore(self._control.data_ptr(), ctypes.c_uint64(int(state['seed'])),
                                                 ctypes.c_uint64(int(state['step'])), self._stream()),
                   'frz_control_restore')

    # ------------------------------------------------------------------------------------------ properties

    @property
    def num_agents(self) -> int:
        return len(self.agents)

    @property
    def max_num_agents(self) -> int:
        return len(self.possible_agents)

    @property
    def unwrapped(self) -> 'BatchedAECEnv':
        return self

    @property
    def terminated(self) -> torch.Tensor:
        """bool [B]; every agent of an environment terminates together (reference utils/env.py:341-349)."""
        return self._terminated.view(torch.bool)

    @property
    def truncated(self) -> torch.Tensor:
        return self._truncated.view(torch.bool)

    @property
    def finished(self) -> torch.Tensor:
        return torch.logical_or(self.terminated, self.truncated)

    @property
    def agent_task_count(self) -> torch.Tensor:
        """int32 [A, B] like the reference (utils/env.py:160); a transposed view of the kernel's [B, A] buffer."""
        return self._agent_task_count.t()

    @property
    def rewards(self) -> Dict[str, torch.Tensor]:
        """Per-agent rewards of the last environment step.  In the middle of an AEC round the reference has cleared
        them to zero (utils/env.py:215,251-254); that is mirrored without touching device memory."""
        if self._mid_cycle:
            return {agent: self._zero_rewards for agent in self.agents}
        return self._reward_views

    # ------------------------------------------------------------------------------------------ allocation

    def _allocate_runtime(self, num_agents: int) -> None:
        """Persistent AEC buffers (what the reference re-creates in reset, utils/env.py:129-160)."""
        B, A, dev = self.parallel_envs, num_agents, self.device
        self._actions = torch.zeros((B, A, 2), dtype=torch.int32, device=dev)
        self._rewards = torch.zeros((B, A), dtype=torch.float32, device=dev)
        self._cumulative = torch.zeros((B, A), dtype=torch.float32, device=dev)
        self._terminated = torch.zeros(B, dtype=torch.uint8, device=dev)
        self._truncated = torch.zeros(B, dtype=torch.uint8, device=dev)
        self.num_moves = torch.zeros(B, dtype=torch.int32, device=dev)
        self.environment_task_count = torch.zeros(B, dtype=torch.int32, device=dev)
        self._agent_task_count = torch.zeros((B, A), dtype=torch.int32, device=dev)
        self._zero_rewards = torch.zeros(B, dtype=torch.float32, device=dev)

    def _stream(self) -> ctypes.c_void_p:
        return _lib.stream_handle(self.device)

    def _horizon(self) -> int:
        return 2**31 - 1 if self.max_steps is None else int(self.max_steps)

    # ------------------------------------------------------------------------------------------ reset

    @torch.no_grad()
    def reset(self, seed=None, options: Optional[Dict[str, Any]] = None) -> None:
        """Reset every environment (reference utils/env.py:95-160). Subclasses extend this with their state."""
        options = options or {}
        if options.get('max_steps') is not None:
            self.max_steps = options['max_steps']
        self._log_label = options.get('log_label')
        self.log_description = options.get('log_description')

        if options.get('skip_seeding'):
            if self._seed_value is None:
                raise ValueError("Seed must be set before skipping seeding is possible")
        else:
            self._seed_value = _seed_to_u64(seed)
        self.seeds = torch.full((self.parallel_envs, ), self._seed_value & 0x7FFFFFFF, dtype=torch.int32,
                                device=self.device)
        _lib.check(self._lib.frz_control_init(self._control.data_ptr(), ctypes.c_uint64(self._seed_value), self._stream()),
                   'frz_control_init')

        self.agents = self.possible_agents
        self.infos = {agent: {} for agent in self.agents}
        self._reward_views = {a: self._rewards[:, i] for i, a in enumerate(self.agents)}
        self._cumulative_rewards = {a: self._cumulative[:, i] for i, a in enumerate(self.agents)}
        terminated, truncated = self.terminated, self.truncated
        self.terminations = {a: terminated for a in self.agents}
        self.truncations = {a: truncated for a in self.agents}
        self.actions = {a: self._actions[:, i] for i, a in enumerate(self.agents)}
        self._agent_slot = {a: i for i, a in enumerate(self.agents)}
        self._mid_cycle = False

        self.agent_selector = AgentSelector(self.agents)
        self.agent_selection = self.agent_selector.reset()

    @torch.no_grad()
    def reset_batches(self, batch_indices, seed=None, options: Optional[Dict[str, Any]] = None) -> None:
        """Reset only the environments ``batch_indices`` (reference utils/env.py:163-189; index list or tensor)."""
        mask = torch.zeros(self.parallel_envs, dtype=torch.uint8, device=self.device)
        mask[torch.as_tensor(batch_indices, device=self.device, dtype=torch.int64)] = 1
        self._reset_masked(mask)

    @abstractmethod
    def _reset_masked(self, mask: Optional[torch.Tensor]) -> None:
        """Launch the domain's reset kernel for the environments selected by ``mask`` (uint8 [B]; None = all)."""

    # ------------------------------------------------------------------------------------------ step

    @abstractmethod
    def step_environment(self) -> Tuple[Dict[str, torch.Tensor], Dict[str, torch.Tensor], Dict[str, Dict]]:
        """Launch the domain's fused step kernel; returns (rewards, terminations, infos) views."""

    @abstractmethod
    def update_actions(self) -> None:
        """Refresh task counts / action masks from the current state (already done inside the fused step)."""

    @abstractmethod
    def update_observations(self) -> None:
        """Refresh observations from the current state (already done inside the fused step)."""

    @abstractmethod
    def action_space(self, agent: str):
        """Per-environment action spaces of ``agent`` (a lazily materialised ``Space.Vector``)."""

    @abstractmethod
    def observation_space(self, agent: str):
        """Per-environment observation spaces of ``agent``."""

    @torch.no_grad()
    def step(self, actions: torch.Tensor) -> None:
        """AEC step of the currently selected agent (reference utils/env.py:203-242).

        The agent's ``[B, 2]`` actions are staged into the device action table; when the last agent of the round has
        acted, one fused launch advances every environment.
        """
        self._actions[:, self._agent_slot[self.agent_selection]].copy_(actions, non_blocking=True)
        self._mid_cycle = True
        if self.agent_selector.is_last():
            self._advance()
        self.agent_selection = self.agent_selector.next()

    @torch.no_grad()
    def step_all(self, actions: Optional[torch.Tensor] = None) -> None:
        """One environment step for all agents at once: ``actions`` int32 [B, A, 2] (None = already staged in
        ``env._actions``, e.g. by ``sample_actions``).  Equivalent to A consecutive ``step`` calls."""
        if actions is not None and actions.data_ptr() != self._actions.data_ptr():
            self._actions.copy_(actions, non_blocking=True)
        self._advance()

    # ------------------------------------------------------------------------------------------ host-buffer step

    def _host_entry(self):
        """The domain's ``frz_<domain>_step_host`` entry point."""
        raise NotImplementedError

    def _host_pipeline(self, chunks: Optional[int]):
        """Page-locked result buffers, per-slice control blocks and streams of ``step_host`` (built once per slicing)."""
        B, A = self.parallel_envs, len(self.possible_agents)
        if chunks is None:  # slices of >= 12 288 environments: below that the copies are too short to be worth hiding
            chunks = max(1, min(HOST_STEP_MAX_DEFAULT_CHUNKS, B // 12288))
        chunks = max(1, min(int(chunks), _lib.MAX_CHUNKS))
        cached = getattr(self, '_host_state', None)
        if cached is not None and cached['chunks'] == chunks:
            return cached
        state = dict(
            chunks=chunks,
            # (initialised from the device: a step that is skipped because every environment is done writes nothing)
            rewards=self._rewards.cpu().pin_memory(),
            done=torch.stack([self._terminated, self._truncated]).cpu().pin_memory(),
            controls=torch.zeros((chunks, 8), dtype=torch.int64, device=self.device),
            streams=[torch.cuda.Stream(self.device) for _ in range(chunks)],
            # the stream the pipeline is issued on: a capturable one (the legacy default stream is not), so that the
            # library can replay the whole pipeline as one CUDA graph from the second call on
            main=torch.cuda.Stream(self.device),
            packed=torch.zeros((B, A, 2), dtype=torch.int16, device=self.device),  # staging of int16 action uploads
        )
        if getattr(self, '_host_events', None) is None:  # this environment's own events (never shared, include/frz.h)
            handle = ctypes.c_void_p()
            _lib.check(self._lib.frz_host_pipeline_create(ctypes.byref(handle)), 'frz_host_pipeline_create')
            self._host_events = handle
        handles = (ctypes.c_void_p * chunks)(*[stream.cuda_stream for stream in state['streams']])
        block = _lib.HostStep()
        block.rewards = state['rewards'].data_ptr()
        block.terminated = state['done'][0].data_ptr()
        block.truncated = state['done'][1].data_ptr()
        block.chunk_controls = state['controls'].data_ptr()
        block.streams = ctypes.cast(handles, ctypes.POINTER(ctypes.c_void_p))
        block.chunks = chunks
        block.packed_actions = state['packed'].data_ptr()
        block.pipeline = self._host_events
        done = state['done'].view(torch.bool)
        state.update(handles=handles, block=block, terminated=done[0], truncated=done[1])
        self._host_state = state
        return state

    # ------------------------------------------------------------------------------------------ observations on the host

    def _observation_download(self) -> Tuple[Dict[str, torch.Tensor], Dict[str, Tuple[torch.Tensor, int]]]:
        """What a policy on the host reads after a step: ``(dense, ragged)``.  ``dense`` arrays leave the device whole;
        ``ragged`` maps a name to ``(padded array [B, groups, capacity, ...] or [B, capacity, ...], groups)`` of which
        only the first ``environment_task_count[b]`` rows per environment (and row block) are live.  Domain hook."""
        raise NotImplementedError

    def _gather_state(self):
        state = getattr(self, '_gather', None)
        if state is not None:
            return state
        B, dev = self.parallel_envs, self.device
        dense, ragged = self._observation_download()
        arrays = (_lib.GatherArray * max(1, len(ragged)))()
        packed, layout = {}, {}
        for at, (name, (tensor, groups)) in enumerate(ragged.items()):
            capacity = tensor.shape[2] if tensor.dim() >= 3 and groups > 1 else tensor.shape[1]
            row_shape = tuple(tensor.shape[3:] if groups > 1 else tensor.shape[2:])
            row_bytes = tensor.element_size() * math.prod(row_shape)
            packed[name] = torch.empty(tensor.numel(), dtype=tensor.dtype, device=dev)
            arrays[at].src, arrays[at].dst = tensor.data_ptr(), packed[name].data_ptr()
            arrays[at].row_bytes, arrays[at].capacity, arrays[at].groups = row_bytes, capacity, groups
            layout[name] = (groups, row_shape, row_bytes // tensor.element_size())
        state = dict(
            dense=dense, ragged=ragged, arrays=arrays, packed=packed, layout=layout,
            offsets=torch.zeros(B + 1, dtype=torch.int32, device=dev),
            scratch=torch.zeros((B + 1023) // 1024 + 1, dtype=torch.int32, device=dev),
            host_counts=torch.zeros(B, dtype=torch.int32).pin_memory(),
            host_offsets=torch.zeros(B + 1, dtype=torch.int32).pin_memory(),
            host_dense={name: torch.empty(tensor.shape, dtype=tensor.dtype).pin_memory() for name, tensor in dense.items()},
            host_packed={name: torch.empty(tensor.numel(), dtype=tensor.dtype).pin_memory() for name, tensor in packed.items()},
        )
        self._gather = state
        return state

    def _enqueue_gather(self, state) -> None:
        """Compaction kernels + the downloads whose size is known up front (counts, offsets, dense arrays), on the
        current stream of the device."""
        _lib.check(self._lib.frz_gather_live_rows(self.environment_task_count.data_ptr(), self.parallel_envs,
                                                  state['offsets'].data_ptr(), state['scratch'].data_ptr(), state['arrays'],
                                                  len(state['ragged']), self._stream()), 'frz_gather_live_rows')
        state['host_counts'].copy_(self.environment_task_count, non_blocking=True)
        state['host_offsets'].copy_(state['offsets'], non_blocking=True)
        for name, tensor in state['dense'].items():
            state['host_dense'][name].copy_(tensor, non_blocking=True)

    def _finish_gather(self, state) -> Dict[str, Any]:
        """Second half, after the stream has been synchronised once: the packed rows, exactly as many bytes as are live
        (one copy per array), then the host views."""
        B = self.parallel_envs
        total = int(state['host_offsets'][B])
        moved = state['host_counts'].numel() * 4 + state['host_offsets'].numel() * 4
        moved += sum(t.numel() * t.element_size() for t in state['host_dense'].values())
        out: Dict[str, Any] = {'counts': state['host_counts'], 'offsets': state['host_offsets'], 'total': total}
        out.update(state['host_dense'])
        for name, device_rows in state['packed'].items():
            groups, row_shape, row_elements = state['layout'][name]
            live = total * groups * row_elements
            host_rows = state['host_packed'][name][:live]
            if live:
                host_rows.copy_(device_rows[:live], non_blocking=True)
            moved += live * device_rows.element_size()
            # environment b's block g = rows [offsets[b] * groups + g * counts[b], ... + counts[b])
            out[name] = host_rows.view((total * groups, ) + row_shape) if row_shape else host_rows
        torch.cuda.current_stream(self.device).synchronize()
        out['bytes'] = moved
        return out

    @torch.no_grad()
    def gather_observations(self) -> Dict[str, Any]:
        """The observations of the current state in page-locked HOST memory, packed like the reference's jagged nested
        tensors (a value buffer + offsets; wildfire.py:669-717, rideshare.py:398-467): ``counts`` int32 [B] live tasks
        per environment, ``offsets`` int32 [B + 1] their exclusive prefix sum, the dense arrays (``self_obs``,
        ``agent_task_count``, ...) whole, and for every padded array its live rows only -- environment b's row block
        g of array ``name`` is ``out[name][offsets[b] * groups + g * counts[b] :][:counts[b]]`` (groups = 1 for task
        observations, = agents for the action / task masks).  The live rows are compacted on the device
        (``frz_gather_live_rows``) so that each array crosses PCIe in one copy of exactly its live bytes.  The
        returned tensors are overwritten by the next call.  ``out['bytes']`` = bytes copied device -> host."""
        with torch.cuda.device(self.device):
            state = self._gather_state()
            self._enqueue_gather(state)
            torch.cuda.current_stream(self.device).synchronize()
            return self._finish_gather(state)

    @torch.no_grad()
    def step_host(self, host_actions: torch.Tensor, chunks: Optional[int] = None, observations: bool = False
                  ) -> Tuple[torch.Tensor, ...]:
        """One environment step for callers that live on the host: ``host_actions`` is a page-locked int32 -- or int16 /
        int8 (at most 127 tasks per environment), which halve / quarter the upload and are widened on the device --
        ``[B, A, 2]`` CPU tensor (agent order = ``env.agents``);
        returns page-locked CPU tensors ``(rewards f32 [B, A],
        terminated bool [B], truncated bool [B])`` that are valid when the call returns and are overwritten by the next
        ``step_host``.  Equivalent to copying the actions to the device, ``step_all`` and copying the results back,
        but the batch is cut into ``chunks`` slices whose uploads, step kernels and downloads overlap
        (``frz_<domain>_step_host``, include/frz.h); the results are bit-identical.  Observations, masks and counts stay
        on the device as usual -- unless ``observations=True``: then the call also returns, as a fourth element, what
        ``gather_observations()`` returns (the next observations in host memory, packed), which is what a policy that
        runs on the CPU needs for its next decision."""
        B, A = self.parallel_envs, len(self.possible_agents)
        if (host_actions.device.type != 'cpu' or host_actions.dtype not in (torch.int32, torch.int16, torch.int8)
                or not host_actions.is_contiguous() or tuple(host_actions.shape) != (B, A, 2)
                or not host_actions.is_pinned()):
            raise ValueError(f'step_host expects a page-locked contiguous int32 / int16 / int8 CPU tensor of shape {(B, A, 2)} '
                             '(torch.empty(..., dtype=torch.int32).pin_memory())')
        state = self._host_pipeline(chunks)
        state['block'].actions = host_actions.data_ptr()
        state['block'].action_format = {torch.int32: _lib.HOST_ACTIONS_I32, torch.int16: _lib.HOST_ACTIONS_I16,
                                        torch.int8: _lib.HOST_ACTIONS_I8}[host_actions.dtype]
        current, main = torch.cuda.current_stream(self.device), state['main']
        main.wait_stream(current)
        _lib.check(self._host_entry()(ctypes.byref(self._params), ctypes.byref(self._io), B,
                                      ctypes.byref(state['block']), ctypes.c_void_p(main.cuda_stream)), 'step_host')
        current.wait_stream(main)  # later work on the caller's stream sees the stepped state
        # host-side bookkeeping while the device works; the views do not depend on the data
        self._mid_cycle = False
        self._rebind_outputs()
        if self.log_directory is not None:
            self._log_environment()
        if not observations:
            main.synchronize()
            return state['rewards'], state['terminated'], state['truncated']
        with torch.cuda.stream(main):  # compaction + fixed-size downloads ride on the same synchronisation
            gather = self._gather_state()
            self._enqueue_gather(gather)
            main.synchronize()
            packed = self._finish_gather(gather)
        return state['rewards'], state['terminated'], state['truncated'], packed

    def _advance(self) -> None:
        _, _, infos = self.step_environment()
        self.infos = infos
        self._mid_cycle = False
        self._rebind_outputs()
        if self.log_directory is not None:
            self._log_environment()

    # ------------------------------------------------------------------------------------------ logging tap

    def _log_environment(self, reset: bool = False) -> None:
        """Queue one CSV row per environment (reference utils/env.py:256-271 -> CSVLogger); nothing here waits for the
        device: see utils/logging_tap.py."""
        if self._tap is None:
            from free_range_zoo_b200.utils.logging_tap import LoggingTap
            if str(self.log_directory).startswith('sqlite://'):  # reference utils/env.py:65-86
                from free_range_zoo_b200.utils.sql_tap import SqliteSink
                sink = SqliteSink(self.log_directory, self.metadata['name'], self.parallel_envs)
                self._tap = LoggingTap(self.log_directory, self.parallel_envs, self.device, self._log_snapshot,
                                       self._log_records, sink=sink, agents=self.possible_agents)
            else:
                self._tap = LoggingTap(self.log_directory, self.parallel_envs, self.device, self._log_snapshot,
                                       self._log_columns,
                                       override_initialization_check=self.override_initialization_check)
        self._tap.capture(reset, self.log_description, getattr(self, '_log_label', None))

    def flush_logs(self) -> None:
        """Wait until every queued log row is on disk."""
        if self._tap is not None:
            self._tap.flush()

    def _log_snapshot(self) -> Dict[str, torch.Tensor]:
        """Live device tensors the log rows are built from (domain tensors are added by the subclasses)."""
        return dict(actions=self._actions, rewards=self._rewards, num_moves=self.num_moves, terminated=self._terminated,
                    truncated=self._truncated, env_task_count=self.environment_task_count,
                    agent_task_count=self._agent_task_count)

    def _log_columns(self, host: Dict[str, Any], reset: bool) -> Dict[str, Any]:
        """Ordered CSV columns from the host copy of a snapshot: state columns, the per-agent action / reward columns,
        step, complete, the per-agent mappings, then the domain's extra columns (logging_handlers.py:77-100)."""
        from free_range_zoo_b200.utils.logging_tap import nested
        B = self.parallel_envs
        columns = dict(self._log_state_columns(host))
        finished = (host['terminated'] != 0) | (host['truncated'] != 0)
        for index, agent in enumerate(self.possible_agents):
            columns[f'{agent}_action'] = [None] * B if reset else [nested(host['actions'][b, index]) for b in range(B)]
            columns[f'{agent}_rewards'] = [None] * B if reset else host['rewards'][:, index].astype(float).tolist()
        columns['step'] = [-1] * B if reset else host['num_moves'].tolist()
        columns['complete'] = [None] * B if reset else finished.tolist()
        for index, agent in enumerate(self.possible_agents):
            action_map, observation_map = self._log_mappings(host, index)
            columns[f'{agent}_action_map'] = action_map
            columns[f'{agent}_observation_map'] = observation_map
        columns.update(self._log_extra_columns(host, reset))
        return columns

    def _log_records(self, host: Dict[str, Any], reset: bool) -> Dict[str, Any]:
        """The SQL sink's record of one snapshot (utils/sql_tap.py::SqliteSink.write): the same state / mapping cells as
        the CSV columns, raw action and reward values per agent (logging_handlers.py:160-241)."""
        state = dict(self._log_state_columns(host))
        state.update({name: cells for name, cells in self._log_extra_columns(host, reset).items()
                      if name not in ('burnouts', 'putouts')})
        agents = {}
        for index, agent in enumerate(self.possible_agents):
            action_map, observation_map = self._log_mappings(host, index)
            agents[agent] = dict(reward=host['rewards'][:, index], action_field=host['actions'][:, index, 1],
                                 task_field=host['actions'][:, index, 0], action_map=action_map,
                                 observation_map=observation_map)
        return dict(timestep=host['num_moves'], state=state, agents=agents)

    def _log_state_columns(self, host) -> Dict[str, Any]:
        raise NotImplementedError

    def _log_mappings(self, host, agent_index: int):
        raise NotImplementedError

    def _log_extra_columns(self, host, reset: bool) -> Dict[str, Any]:
        return {}

    def _rebind_outputs(self) -> None:
        """Re-create the lazily evaluated observation views after the device buffers changed."""
        self.update_observation_views()
        if self.detach_outputs:
            self._reward_views = {a: self._rewards[:, i].clone() for i, a in enumerate(self.agents)}
            terminated, truncated = self.terminated.clone(), self.truncated.clone()
            self.terminations = {a: terminated for a in self.agents}
            self.truncations = {a: truncated for a in self.agents}

    @abstractmethod
    def update_observation_views(self) -> None:
        """Build ``self.observations`` (dict agent -> ObservationDict) as views of the device buffers."""

    @torch.no_grad()
    def observe(self, agent: str) -> ObservationDict:
        return self.observations[agent]

    @torch.no_grad()
    def state(self):
        return self._state

    def last(self, observe: bool = True):
        """pettingzoo AEC convenience: (observation, reward, termination, truncation, info) of the selected agent."""
        agent = self.agent_selection
        observation = self.observe(agent) if observe else None
        return (observation, self._cumulative_rewards[agent], self.terminations[agent], self.truncations[agent],
                self.infos[agent])

    # ------------------------------------------------------------------------------------------ diagnostics

    def control_block(self) -> Dict[str, int]:
        """Host copy of the device control block (synchronises; diagnostics / tests only)."""
        raw = self._control.cpu().numpy().tobytes()
        block = _lib.Control.from_buffer_copy(raw)
        return {name: getattr(block, name) for name, _ in _lib.Control._fields_ if name != 'reserved'}

    def check_errors(self) -> None:
        """Raise ``ValueError`` for data-dependent faults the kernels recorded (the reference raises these eagerly,
        at the price of a host sync per step: cybersecurity.py:341-363).  Synchronises; call it off the hot path."""
        word = self.control_block()['error_word']
        if word:
            self._control.view(torch.int32)[7] = 0
            reasons = [text for bit, text in _lib.FAULT_NAMES.items() if word & bit]
            raise ValueError('invalid actions were submitted: ' + '; '.join(reasons))

    # ------------------------------------------------------------------------------------------ CUDA graph

    def capture_graph(self, sample: bool = False, sampler_seed: int = 2026, steps: int = 1) -> None:
        """Capture ``steps`` x ``[sample_actions ->] step`` in a CUDA graph; ``replay()`` then costs one graph launch
        per ``steps`` environment steps (small batches are bound by launch latency: several steps per launch keep the
        kernels back to back).  Kernel arguments are pointer-stable and the step counter lives on the device, so the
        graph needs no updates.  Capturing leaves the environment where it was (the warm-up launch is undone).  Replays
        do not feed the logging tap, so a ``log_directory`` is refused."""
        if self.log_directory is not None:
            raise RuntimeError('graph replays bypass the logging tap; capture_graph needs log_directory=None')
        torch.cuda.synchronize(self.device)
        # the warm-up launch outside capture (lazy module load, occupancy query) steps for real: every tensor the
        # kernels write -- state, outputs, the control block with the Philox step counter -- is put back afterwards,
        # so capturing does not advance the environment
        saved = {name: tensor.clone() for name, tensor in self._bound.items()
                 if tensor is not None and not name.startswith('init_')}
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            if sample:
                self.sample_actions(sampler_seed)
            self.step_environment()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        for name, tensor in saved.items():
            self._bound[name].copy_(tensor)
        del saved
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for _ in range(max(1, int(steps))):
                if sample:
                    self.sample_actions(sampler_seed)
                self.step_environment()
        self._graph = graph

    def replay(self) -> None:
        """Launch the captured graph (``steps`` environment steps); outputs are re-bound like after ``step``."""
        self._graph.replay()
        self._mid_cycle = False
        if self.detach_outputs:
            self._rebind_outputs()

    def sample_actions(self, sampler_seed: int = 2026) -> torch.Tensor:
        """Uniform random legal actions for every agent, generated on the device into the staged action table
        (the caller side of the path: replaces ``action_space(agent).sample_nested()`` per agent per step)."""
        raise NotImplementedError
