"""Turn the raw ncu artefacts a gpurun call brought back (gpurun_out/, scratch) into the small text summaries that are
committed under profiles/.

    python profiles/summarize.py launches gpurun_out/<launches>.csv            > profiles/<name>_launches.txt
    python profiles/summarize.py kernel   gpurun_out/<report>.ncu-rep [top_n]  > profiles/<name>_ncu.txt

``launches`` groups the per-launch ``gpu__time_duration.sum`` list by kernel (count, total, mean, share of GPU time).
``kernel`` prints the roofline-relevant raw metrics of every profiled launch, the warp-stall breakdown and the hottest
source lines (needs -lineinfo, which csrc/Makefile passes).
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

RAW_METRICS = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
    'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
    'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
    'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
    'smsp__issue_active.avg.per_cycle_active', 'smsp__warps_eligible.avg.per_cycle_active',
    'smsp__warps_active.avg.per_cycle_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active',
]


def launches(path: str) -> None:
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    header = rows[0]
    name_at, value_at = header.index('Kernel Name'), header.index('Metric Value')
    per_kernel = defaultdict(list)
    for row in rows[1:]:
        per_kernel[row[name_at]].append(float(row[value_at].replace(',', '')))
    total = sum(sum(v) for v in per_kernel.values())
    print(f'# {path}: {sum(len(v) for v in per_kernel.values())} launches, {total / 1e3:.1f} us of GPU time '
          '(ncu-serialised, cold cache: compare shares, not absolutes)')
    print(f'{"share":>6s} {"count":>6s} {"total_us":>10s} {"mean_us":>9s}  kernel')
    for name, values in sorted(per_kernel.items(), key=lambda kv: -sum(kv[1])):
        print(f'{sum(values) / total:6.3f} {len(values):6d} {sum(values) / 1e3:10.1f} {sum(values) / len(values) / 1e3:9.2f}  {name[:150]}')


def kernel(path: str, top: int) -> None:
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units = rows[0], rows[1]
    print(f'# {path}')
    for number, row in enumerate(rows[2:]):
        print(f'## launch {number}: {row[header.index("Kernel Name")][:120]}')
        for metric in RAW_METRICS:
            if metric in header:
                print(f'  {metric:72s} {row[header.index(metric)]:>16s} {units[header.index(metric)]}')
        traffic = sum(float(row[header.index(m)].replace(',', '')) * {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}[units[header.index(m)]]
                      for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum') if m in header)
        print(f'  {"traffic = dram read + write":72s} {traffic / 1e6:16.1f} MB')
        stalls = []
        for at, name in enumerate(header):
            if name.startswith('smsp__average_warp_latency_issue_stalled_') and name.endswith('.ratio') is False:
                continue
            if name.startswith('smsp__average_warps_issue_stalled_') and name.endswith('_per_issue_active.ratio'):
                try:
                    stalls.append((float(row[at]), name[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]))
                except ValueError:
                    pass
        if stalls:
            print('  warp stall reasons (warps stalled per issue-active cycle): ' +
                  ', '.join(f'{name} {value:.2f}' for value, name in sorted(stalls, reverse=True)[:8]))
    source = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--launch-count', '1'],
                            capture_output=True, text=True).stdout
    lines, current, head = [], None, None
    for record in csv.reader(io.StringIO(source)):
        if not record:
            continue
        if record[0] == 'File Path':
            current = record[1].split('/')[-1]
        elif record[0] == 'Line No':
            head = record
        elif head and record[0].isdigit():
            def get(name):
                value = record[head.index(name)] if name in head else ''
                return int(value) if value.isdigit() else 0
            lines.append((current, int(record[0]), record[1].strip(), get('Instructions Executed'), get('# Samples')))
    instructions = sum(r[3] for r in lines) or 1
    samples = sum(r[4] for r in lines) or 1
    print(f'## hottest source lines of launch 0 ({instructions} warp instructions, {samples} stall samples)')
    for name, line, text, inst, smp in sorted(lines, key=lambda r: -r[3])[:top]:
        print(f'  {100 * inst / instructions:5.1f}% inst {100 * smp / samples:5.1f}% smp  {name}:{line:<4d} {text[:100]}')


def traffic(path: str, key: str) -> None:
    """Record the DRAM bytes of the first profiled launch under profiles/traffic.json[key] (key = workload@envs)."""
    import json
    import os
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units, row = rows[0], rows[1], rows[2]
    scale = {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}
    total = sum(float(row[header.index(m)].replace(',', '')) * scale[units[header.index(m)]]
                for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
    store = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'traffic.json')
    data = json.load(open(store)) if os.path.exists(store) else {}
    data[key] = {'dram_bytes_per_launch': total, 'report': os.path.basename(path),
                 'kernel': row[header.index('Kernel Name')][:80]}
    json.dump(data, open(store, 'w'), indent=1, sort_keys=True)
    print(key, total)


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2])
    elif sys.argv[1] == 'traffic':
        traffic(sys.argv[2], sys.argv[3])
    else:
        kernel(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 30)
