"""Benchmark of the fused environment-step path (BASELINE.json metric: env-steps/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload wildfire_c4] [--impl engine|reference]

A "step" is one environment step of ALL ``parallel_envs`` environments with random legal actions.  At N=1 the default
workload is BASELINE.json configs[3] -- wildfire 10x10 / 10 agents, agent+task+frame openness, parallel_envs=65,536 --
the configuration the headline target (>= 50x the host-CPU path, >= 60 % of HBM roofline) is quoted on.  The other
named configs (``--workload wildfire_c1 | rideshare_c2 | cyber_c3``) can be measured the same way.  For N>1 (launched
by ``python -m torch.distributed.run``) every rank steps its own shard of ``parallel_envs`` environments (weak
scaling, no step-path collective; the global env index keys the RNG so trajectories do not depend on N); NCCL is used
only for the barrier, the max-over-ranks timing and the episode-statistics all-reduce.

One JSON line is printed by rank 0:
  value         whole-job env-steps/s, inputs resident in HBM, [sample_actions -> step] captured in one CUDA graph
  roofline      the step kernel alone: algorithmic bytes per launch / CUDA-event duration of the
                launch, against the measured HBM copy bandwidth of MEASURED_PEAKS.json
  e2e           the same metric through the public Parallel API with HOST buffers (``parallel_env.step_host``):
                page-locked actions H2D + step + rewards / dones D2H + stream sync, every step; upload, kernel and
                download are pipelined over slices of the batch inside frz_<domain>_step_host
  cpu_baseline  the CPU oracle port of the same workload on the box's host cores (bounded sample, rank 0, N=1)
``--impl reference`` times the CPU oracle port (the reference's algorithm restated in numpy -- the Python reference
and its uninstallable dependencies cannot travel to the GPU box) on all host cores and prints the same line shape.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# name: domain, preset, parallel_envs per GPU, env kwargs
WORKLOADS = {
    'wildfire_c4': dict(domain='wildfire', preset='wildfire_large', envs=65536, kwargs={}),
    'wildfire_c1': dict(domain='wildfire', preset='wildfire_3x3', envs=1024, kwargs={}),
    'rideshare_c2': dict(domain='rideshare', preset='rideshare_c2', envs=16384, kwargs={}),
    # extra wildfire geometries (kernel tuning; not named by BASELINE.json)
    'wildfire_5x6': dict(domain='wildfire', preset='wildfire_large', envs=262144, kwargs={},
                         preset_kwargs=dict(height=5, width=6, num_agents=6, seed=21)),
    'wildfire_7x8': dict(domain='wildfire', preset='wildfire_large', envs=262144, kwargs={},
                         preset_kwargs=dict(height=7, width=8, num_agents=5, seed=5)),
    'cyber_c3': dict(domain='cybersecurity', preset='cyber_c3', envs=16384,
                     kwargs=dict(show_bad_actions=False, partially_observable=True)),
}
KERNELS = {'wildfire': 'wildfire_step_kernel', 'rideshare': 'rideshare_step_kernel',
           'cybersecurity': 'cyber_step_tiled_kernel'}


def algorithmic_bytes(domain: str, raw, present: float = 0.0) -> float:
    """ALGORITHMIC bytes per env-step (restated in DESIGN.md section 4): every live tensor read
    once and written once in the reference's dtypes, outputs written once, padded int32 observations, u8 masks."""
    if domain == 'wildfire':
        HW, A = raw.max_y * raw.max_x, len(raw.agents)
        state = 12 * HW + 12 * A
        return 2 * state + 8 * A + 4 * A + 2 * A + 16 + 16 * A + 16 * HW + 4 + A * HW + 4 * A
    if domain == 'cybersecurity':
        N, att, dfd = raw._n_nodes, raw._n_att, raw._n_def
        n = att + dfd
        state = 4 * N + 4 * dfd + n
        return 2 * state + 8 * n + 4 * n + 2 * n + 4 + (8 * att + 12 * dfd) + 8 * N + dfd + n
    A, K = len(raw.agents), raw._capacity  # `present` = measured mean passengers per environment over the timed steps
    return 2 * (8 * A + 44 * present) + 8 * A + 4 * A + 2 + 4 + 16 * A + 32 * present + A * K + 4 * (A + 1)


def measured_peak_gbs():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        return float(json.load(open(path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def ncu_traffic(workload: str, parallel_envs: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of one step-kernel launch, from the committed `ncu --set full`
    capture of this workload (profiles/traffic.json, written by profiles/summarize.py); None when there is none."""
    path = os.path.join(ROOT, 'profiles', 'traffic.json')
    if not os.path.exists(path):
        return None
    entry = json.load(open(path)).get(f'{workload}@{parallel_envs}')
    return None if entry is None else entry['dram_bytes_per_launch']


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    QUERY = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, index: int):
        self.index, self.samples, self._stop, self._thread = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(['nvidia-smi', f'--query-gpu={self.QUERY}', '--format=csv,noheader,nounits', '-i',
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([field.strip() for field in out.split(',')])
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join()

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
        sm = sorted(float(s[0]) for s in self.samples)
        reasons = set()
        for s in self.samples:
            for name, flag in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), s[3:7]):
                if flag.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': float(self.samples[0][1]), 'samples': len(sm),
                'window': 'graph replays + kernel timing + end-to-end leg',
                'power_w_max': max(float(s[2]) for s in self.samples), 'reasons': sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU arm


def _oracle_rollout(job):
    """Worker: steps one oracle batch and returns the seconds spent inside ``oracle.step`` (action sampling and the
    generation of the injected uniforms are excluded, as SURVEY.md section 8d prescribes for both sides)."""
    workload, envs, steps, seed = job
    import numpy as np

    from free_range_zoo_b200 import presets
    spec = WORKLOADS[workload]
    config = getattr(presets, spec['preset'])(**spec.get('preset_kwargs', {}))
    rng = np.random.default_rng(seed)
    if spec['domain'] == 'wildfire':
        from oracle.wildfire import WildfireOracle
        oracle = WildfireOracle(config, envs, 1 << 30, **spec['kwargs'])
        oracle.reset()
        H, W, A = oracle.H, oracle.W, oracle.A

        def one_step():
            counts = oracle.agent_task_count
            k = np.minimum((rng.random(counts.shape) * (counts + 1)).astype(np.int64), counts)
            actions = np.stack([k, np.where(k == counts, -1, 0)], axis=2).astype(np.int32)
            uniforms = (rng.random((3, envs, H, W), dtype=np.float32), rng.random((5, envs, A), dtype=np.float32))
            start = time.perf_counter()
            oracle.step(actions, *uniforms)
            return time.perf_counter() - start
    elif spec['domain'] == 'cybersecurity':
        from oracle.cybersecurity import CybersecurityOracle
        oracle = CybersecurityOracle(config, envs, 1 << 30, **spec['kwargs'])
        oracle.reset()

        def one_step():
            n = oracle.n_agents
            count = oracle.agent_task_count
            k = np.minimum((rng.random(count.shape) * (count + 1)).astype(np.int64), count)
            actions = np.stack([k, np.where(k == count, -1, 0)], axis=2).astype(np.int32)  # attack / move or noop
            uniforms = (rng.random((1, envs, oracle.N), dtype=np.float32), rng.random((1, envs, n), dtype=np.float32))
            start = time.perf_counter()
            oracle.step(actions, *uniforms)
            return time.perf_counter() - start
    else:
        from oracle.rideshare import RideshareOracle
        oracle = RideshareOracle(config, envs, 1 << 30)
        oracle.reset()

        def one_step():
            actions = np.zeros((envs, oracle.A, 2), np.int32)
            for b in range(envs):
                for a in range(oracle.A):
                    mine = oracle._task_list(b, a)
                    k = min(int(rng.random() * (len(mine) + 1)), len(mine))
                    actions[b, a] = (k, -1) if k == len(mine) else (k, oracle.tables[b][mine[k]][6])
            start = time.perf_counter()
            oracle.step(actions)
            return time.perf_counter() - start

    one_step()
    return sum(one_step() for _ in range(steps))


def cpu_oracle_throughput(workload: str, processes: int, envs_per_process: int, steps: int) -> float:
    """env-steps/s of the CPU oracle: ``processes`` workers, each stepping its own batch (slowest worker counts)."""
    import multiprocessing as mp
    jobs = [(workload, envs_per_process, steps, 100 + i) for i in range(processes)]
    if processes == 1:
        seconds = [_oracle_rollout(jobs[0])]
    else:
        with mp.get_context('fork').Pool(processes) as pool:
            seconds = pool.map(_oracle_rollout, jobs)
    return processes * envs_per_process * steps / max(seconds)


def cpu_sample_size(workload: str):
    """(envs per process, steps): a bounded sample (tens of seconds at most) of the same workload."""
    return {'wildfire': (1024, 10), 'cybersecurity': (4096, 20), 'rideshare': (128, 10)}[WORKLOADS[workload]['domain']]


def run_reference(args):
    """--impl reference: the CPU oracle port on all host cores (rank 0 only; other ranks exit without work)."""
    if int(os.environ.get('RANK', '0')) != 0:
        return
    spec = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    envs_per_process, sample_steps = cpu_sample_size(args.workload)
    steps = max(1, min(args.steps, sample_steps))
    cpu_oracle_throughput(args.workload, cores, envs_per_process, 1)  # warm-up (page-in, numpy import per worker)
    start = time.perf_counter()
    value = cpu_oracle_throughput(args.workload, cores, envs_per_process, steps)
    wall = time.perf_counter() - start
    sample = (f'{cores} processes x {envs_per_process} envs x {steps} steps of {args.workload}; numpy oracle port '
              f'(oracle/{spec["domain"]}.py) of the reference CPU path; step() only, action sampling excluded')
    line = {
        'impl': 'reference', 'metric': 'env-steps/sec', 'value': value, 'unit': 'env-steps/s', 'n_gpus': args.gpus,
        'steps': steps, 'warmup': 1, 'ms_per_step': 1e3 * cores * envs_per_process / value, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'int32+f32', 'data': 'synthetic',
        'config': {'workload': args.workload, 'domain': spec['domain'], 'preset': spec['preset'],
                   'parallel_envs_sampled': cores * envs_per_process},
        'cpu_baseline': {'value': value, 'unit': 'env-steps/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'env-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0, 'wall_s': wall,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU arm


def run_engine(args):
    import importlib

    import torch
    import torch.distributed as dist

    from free_range_zoo_b200 import presets
    from free_range_zoo_b200.distributed import all_reduce_statistics, episode_statistics

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)

    spec = WORKLOADS[args.workload]
    domain = spec['domain']
    B = args.parallel_envs or spec['envs']
    config = getattr(presets, spec['preset'])(**spec.get('preset_kwargs', {}))
    module = importlib.import_module(f'free_range_zoo_b200.envs.{domain}_v0')
    env = module.parallel_env(parallel_envs=B, max_steps=1 << 30, configuration=config, device=device,
                              env_offset=rank * B, **spec['kwargs'])
    raw = env.unwrapped
    agents = len(raw.agents)
    K, W = args.steps, max(3, args.warmup)
    SEED, SAMPLER = 2026, 2026

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- (1) device-resident throughput: [sample_actions -> step] as one CUDA graph per step
    # (several steps per graph launch when the batch is small enough for the launch latency to matter)
    per_replay = max(d for d in (10, 5, 4, 3, 2, 1) if K % d == 0 and W % d == 0) if B <= 32768 else 1
    env.reset(seed=SEED)
    raw.capture_graph(sample=True, sampler_seed=SAMPLER, steps=per_replay)
    env.reset(seed=SEED)
    for _ in range(W // per_replay):
        raw.replay()
    barrier()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # clocks / throttle reasons are sampled from here to the end of the end-to-end leg: the device-timed region alone
    # lasts a few milliseconds, less than one nvidia-smi query
    clocks = ClockSampler(local_rank)
    clocks.__enter__()
    start.record()
    for _ in range(K // per_replay):
        raw.replay()
    stop.record()
    barrier()
    graph_ms = max_over_ranks(start.elapsed_time(stop))
    value = world * B * K / (graph_ms * 1e-3)

    # ---- (2) the step kernel alone, CUDA events on the launching stream.  Two measurements:
    #   eager:    an event pair around each frz_<domain>_step launch (includes the host's launch latency when the kernel
    #             is shorter than a launch, i.e. for the small named batches)
    #   in graph: R x [sample, step] captured in one CUDA graph minus R x [sample] -- back-to-back kernels, no host in
    #             between; this is the duration the roofline fraction uses
    # Every measurement covers the same window of the same seeded rollout: reset, W warm-up steps, K timed steps (the
    # cost of a step follows the number of live tasks, which changes along a rollout).
    def restart():
        env.reset(seed=SEED)
        for _ in range(W):
            raw.sample_actions(SAMPLER)
            raw.step_environment()
        barrier()

    pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    restart()
    tasks_seen = torch.zeros((), dtype=torch.float64, device=device)  # live tasks per environment, summed over steps
    for before, after in pairs:
        raw.sample_actions(SAMPLER)
        before.record()
        raw.step_environment()
        after.record()
        tasks_seen += raw.environment_task_count.sum()
    barrier()
    eager_ms = sum(b.elapsed_time(a) for b, a in pairs) / K
    mean_tasks = float(tasks_seen.item()) / (K * B)
    raw.check_errors()

    def in_graph_ms(with_step: bool) -> float:
        repeats = max(d for d in (10, 5, 4, 3, 2, 1) if K % d == 0)
        torch.cuda.synchronize(device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for _ in range(repeats):
                raw.sample_actions(SAMPLER)
                if with_step:
                    raw.step_environment()
        graph.replay()  # first replay uploads the graph
        restart()
        start.record()
        for _ in range(K // repeats):
            graph.replay()
        stop.record()
        torch.cuda.synchronize(device)
        return start.elapsed_time(stop) / K

    kernel_ms = max(in_graph_ms(True) - in_graph_ms(False), 1e-6)
    bytes_per_env = algorithmic_bytes(domain, raw, mean_tasks)
    peak, peak_kind = measured_peak_gbs()
    achieved = bytes_per_env * B / (kernel_ms * 1e-3) / 1e9
    raw.check_errors()

    # ---- (3) end to end through the public Parallel API with host buffers.  Legal actions for every step come from
    # a recorded dry run of the same seeded rollout (the engine is deterministic, so they stay legal on replay).
    host_actions = torch.empty((W + K, B, agents, 2), dtype=torch.int32).pin_memory()
    env.reset(seed=SEED)
    for t in range(W + K):
        raw.sample_actions(SAMPLER)
        host_actions[t].copy_(raw._actions, non_blocking=True)
        raw.step_all()
    torch.cuda.synchronize(device)
    env.reset(seed=SEED)

    def e2e_step(t):
        # public API with HOST buffers: page-locked actions in, page-locked rewards / terminated / truncated out; the
        # upload, the fused step and the download are pipelined over slices of the batch (frz_<domain>_step_host);
        # the call returns after the stream synchronised, i.e. when the caller can read the results
        return env.step_host(host_actions[t], args.host_chunks or None)

    for t in range(W):
        e2e_step(t)
    barrier()
    start.record()
    for t in range(W, W + K):
        e2e_step(t)
    stop.record()
    barrier()
    e2e_ms = max_over_ranks(start.elapsed_time(stop))
    e2e_value = world * B * K / (e2e_ms * 1e-3)
    clocks.__exit__()
    host_rewards, host_done = raw._host_state['rewards'], raw._host_state['done']  # what the last step_host returned
    host_terminated, host_truncated = host_done[0], host_done[1]
    h2d = host_actions[0].numel() * 4
    d2h = host_rewards.numel() * 4 + host_terminated.numel() + host_truncated.numel()
    # the host copies are the device's results (checked outside the timed region)
    assert torch.equal(host_rewards, raw._rewards.cpu()) and torch.equal(host_terminated, raw._terminated.cpu())

    # ---- episode statistics: the only collective, off the step path
    stats = all_reduce_statistics(episode_statistics(raw._cumulative, raw.terminated, raw.truncated, raw.num_moves))

    if rank == 0:
        line = {
            'metric': 'env-steps/sec', 'value': value, 'unit': 'env-steps/s', 'n_gpus': world, 'steps': K, 'warmup': W,
            'ms_per_step': graph_ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'int32+f32', 'data': 'synthetic',
            'config': {
                'workload': args.workload, 'domain': domain, 'preset': spec['preset'], 'parallel_envs_per_gpu': B,
                'parallel_envs_total': world * B, 'agents': agents, 'agent_steps_per_s': value * agents,
                'actions': 'uniform random legal actions sampled on device (Philox), inside the timed region',
                'steps_per_graph_launch': per_replay,
                'l2': (f'state+outputs per step = {bytes_per_env * B / 1e6:.0f} MB > 126 MB L2 (inputs larger than L2)'
                       if bytes_per_env * B > 126e6 else
                       f'working set {bytes_per_env * B / 1e6:.1f} MB fits in L2: launch-latency-bound configuration'),
                'parallelism': f'dp{world} (env-batch sharding, no step-path collective)',
            },
            'roofline': {
                'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': ncu_traffic(args.workload, B), 'peak_kind': peak_kind, 'kernel': KERNELS[domain],
                'kernel_ms': kernel_ms, 'kernel_ms_eager_launch': eager_ms,
                'algorithmic_bytes_per_launch': bytes_per_env * B, 'algorithmic_bytes_per_env_step': bytes_per_env,
                'mean_tasks_per_env': mean_tasks,
            },
            'e2e': {'value': e2e_value, 'unit': 'env-steps/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'ms_per_step': e2e_ms / K, 'api': 'parallel_env.step_host (frz_<domain>_step_host)',
                    'slices': raw._host_state['chunks']},
            'gpu_launches': 2 * K,  # K x [sample kernel, step kernel] in the device-timed region
            'clocks': clocks.summary(),
            'stats': {'env_steps_executed': float(stats[0]), 'terminated_envs': float(stats[1]),
                      'truncated_envs': float(stats[2]), 'cumulative_reward_sum': float(stats[3:].sum())},
        }
        if world == 1:
            envs_per_process, cpu_steps = cpu_sample_size(args.workload)
            cpu_value = cpu_oracle_throughput(args.workload, 1, envs_per_process, cpu_steps)
            line['cpu_baseline'] = {
                'value': cpu_value, 'unit': 'env-steps/s', 'cores': 1, 'kind': 'port',
                'sample': f'{envs_per_process} envs x {cpu_steps} steps of {args.workload}, CPU oracle port '
                          f'(oracle/{domain}.py), 1 process; step() only, action sampling excluded'}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--gpus', type=int, default=1)
    parser.add_argument('--steps', type=int, default=50)
    parser.add_argument('--warmup', type=int, default=5)
    parser.add_argument('--impl', default='engine', choices=['engine', 'reference'])
    parser.add_argument('--workload', default='wildfire_c4', choices=sorted(WORKLOADS))
    parser.add_argument('--parallel-envs', type=int, default=0, help='override parallel_envs per GPU')
    parser.add_argument('--host-chunks', type=int, default=0,
                        help='slices of the pipelined host-buffer step of the e2e leg (0 = the engine\'s default)')
    args = parser.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_engine(args)


if __name__ == '__main__':
    main()
