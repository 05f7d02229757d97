"""Development helper: where the time of ``env.step_host`` goes (wall clock per step, 65 536-env wildfire C4).

    python profiles/time_host_step.py [--workload wildfire_c4] [--steps 40]

Prints, per slice count: the bare C-ABI call + stream sync, and the full Python ``step_host``.
"""
import argparse
import ctypes
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import bench
    from free_range_zoo_b200 import _lib, presets
    parser = argparse.ArgumentParser()
    parser.add_argument('--workload', default='wildfire_c4')
    parser.add_argument('--parallel-envs', type=int, default=0)
    parser.add_argument('--steps', type=int, default=40)
    parser.add_argument('--slices', default='1,2,3,4,5,6,8')
    args = parser.parse_args()
    spec = bench.WORKLOADS[args.workload]
    B = args.parallel_envs or spec['envs']
    device = torch.device('cuda', 0)
    module = importlib.import_module(f'free_range_zoo_b200.envs.{spec["domain"]}_v0')
    env = module.parallel_env(parallel_envs=B, max_steps=1 << 30, device=device,
                              configuration=getattr(presets, spec['preset'])(**spec.get('preset_kwargs', {})), **spec['kwargs'])
    raw = env.unwrapped
    A = len(raw.agents)
    steps = args.steps
    host_actions = torch.empty((steps + 5, B, A, 2), dtype=torch.int32).pin_memory()
    env.reset(seed=1)
    for t in range(steps + 5):
        raw.sample_actions(3)
        host_actions[t].copy_(raw._actions, non_blocking=True)
        raw.step_all()
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream(device)
    side = torch.cuda.Stream(device)
    packed = host_actions.to(torch.int16).pin_memory()
    for chunks in [int(c) for c in args.slices.split(',')]:
        row = {'chunks': chunks}
        # direct: every stream call issued per step (no pipeline handle); graph: the library replays the captured
        # pipeline (handle + capturable stream); *_i16: int16 actions
        for mode in ('direct', 'graph', 'graph_i16', 'step_host', 'step_host_i16'):
            env.reset(seed=1)
            state = raw._host_pipeline(chunks)
            block = state['block']
            saved = block.pipeline
            times, issue = [], []
            for t in range(steps + 5):
                torch.cuda.synchronize()
                begin = time.perf_counter()
                if mode.startswith('step_host'):
                    env.step_host(packed[t] if mode.endswith('i16') else host_actions[t], chunks)
                    issued = begin
                else:
                    source = packed if mode.endswith('i16') else host_actions
                    block.actions = source[t].data_ptr()
                    block.action_format = 1 if mode.endswith('i16') else 0
                    block.pipeline = None if mode == 'direct' else saved
                    on = stream if mode == 'direct' else side
                    _lib.check(raw._host_entry()(ctypes.byref(raw._params), ctypes.byref(raw._io), B, ctypes.byref(block),
                                                 ctypes.c_void_p(on.cuda_stream)))
                    issued = time.perf_counter()
                    on.synchronize()
                times.append(time.perf_counter() - begin)
                issue.append(issued - begin)
            block.pipeline = saved
            block.action_format = 0
            times, issue = sorted(times[5:]), sorted(issue[5:])
            row[mode + '_us'] = round(1e6 * times[len(times) // 2], 1)
            if not mode.startswith('step_host'):
                row[mode + '_issue_us'] = round(1e6 * issue[len(issue) // 2], 1)
        print(json.dumps(row), flush=True)
    # the unpipelined sequence on one stream, for comparison
    env.reset(seed=1)
    rewards = torch.empty((B, A), dtype=torch.float32).pin_memory()
    done = torch.empty((2, B), dtype=torch.uint8).pin_memory()
    times = []
    for t in range(steps + 5):
        torch.cuda.synchronize()
        begin = time.perf_counter()
        raw._actions.copy_(host_actions[t], non_blocking=True)
        raw.step_environment()
        rewards.copy_(raw._rewards, non_blocking=True)
        done[0].copy_(raw._terminated, non_blocking=True)
        done[1].copy_(raw._truncated, non_blocking=True)
        stream.synchronize()
        times.append(time.perf_counter() - begin)
    times = sorted(times[5:])
    print(json.dumps({'serial_copy_step_copy_us': round(1e6 * times[len(times) // 2], 1)}))


if __name__ == '__main__':
    main()
