"""Kernel-tuning helper: time the fused step kernel of one workload with CUDA events, for one or more builds of
libfrz.so, each in its own process (FRZ_LIBRARY selects the build).

    python profiles/time_kernel.py --workload wildfire_c4 [--parallel-envs N] [--steps 30] lib_a.so lib_b.so ...

Prints one line per library: mean / min step-kernel time and the roofline fraction bench.py would report.  This is a
development tool; the numbers that are reported come from bench.py.
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(args):
    sys.path.insert(0, ROOT)
    import importlib

    import torch

    import bench
    from free_range_zoo_b200 import presets
    spec = bench.WORKLOADS[args.workload]
    B = args.parallel_envs or spec['envs']
    device = torch.device('cuda', 0)
    module = importlib.import_module(f'free_range_zoo_b200.envs.{spec["domain"]}_v0')
    env = module.parallel_env(parallel_envs=B, max_steps=1 << 30, configuration=getattr(presets, spec['preset'])(**spec.get('preset_kwargs', {})),
                              device=device, **spec['kwargs'], **spec.get('engine_kwargs', {}))
    raw = env.unwrapped
    env.reset(seed=2026)
    for _ in range(args.warmup):
        raw.sample_actions(2026)
        raw.step_environment()
    torch.cuda.synchronize()
    pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for before, after in pairs:
        raw.sample_actions(2026)
        before.record()
        raw.step_environment()
        after.record()
    torch.cuda.synchronize()
    times = [b.elapsed_time(a) for b, a in pairs]
    raw.check_errors()

    # launch-overhead-free estimate for small batches: K x [sample, step] in one CUDA graph minus K x [sample]
    # (same window of the same seeded rollout as the eager measurement: reset, warm-up steps, `steps` timed steps)
    def graph_time(with_step: bool, repeats: int = 10) -> float:
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for _ in range(repeats):
                raw.sample_actions(2026)
                if with_step:
                    raw.step_environment()
        graph.replay()
        env.reset(seed=2026)
        for _ in range(args.warmup):
            raw.sample_actions(2026)
            raw.step_environment()
        torch.cuda.synchronize()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        replays = max(1, args.steps // repeats)
        start.record()
        for _ in range(replays):
            graph.replay()
        stop.record()
        torch.cuda.synchronize()
        return start.elapsed_time(stop) / (replays * repeats)

    graph_us = 1e3 * (graph_time(True) - graph_time(False))
    bytes_per_env = bench.algorithmic_bytes(spec['domain'], raw, float(raw.environment_task_count.float().mean().item()))
    peak, _ = bench.measured_peak_gbs()
    mean = sum(times) / len(times)
    checksum = float(raw._cumulative.double().sum().item())
    print(json.dumps({'library': os.environ.get('FRZ_LIBRARY', 'default'), 'workload': args.workload, 'B': B,
                      'kernel_us_mean': 1e3 * mean, 'kernel_us_min': 1e3 * min(times), 'kernel_us_in_graph': graph_us,
                      'frac': bytes_per_env * B / (mean * 1e-3) / 1e9 / peak, 'reward_checksum': checksum}))


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--workload', default='wildfire_c4')
    parser.add_argument('--parallel-envs', type=int, default=0)
    parser.add_argument('--steps', type=int, default=30)
    parser.add_argument('--warmup', type=int, default=5)
    parser.add_argument('--child', action='store_true')
    parser.add_argument('libraries', nargs='*')
    args = parser.parse_args()
    if args.child:
        return child(args)
    for library in args.libraries or ['']:
        env = dict(os.environ)
        if library:
            env['FRZ_LIBRARY'] = os.path.abspath(library)
        command = [sys.executable, os.path.abspath(__file__), '--child', '--workload', args.workload, '--parallel-envs',
                   str(args.parallel_envs), '--steps', str(args.steps), '--warmup', str(args.warmup)]
        result = subprocess.run(command, env=env, capture_output=True, text=True)
        print(result.stdout.strip() or result.stderr.strip()[-800:], flush=True)


if __name__ == '__main__':
    main()
