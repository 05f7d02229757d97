"""Dynamic instruction counts per source line / phase: joins the per-instruction counts of an ncu report (SASS page)
with the line table of the same build (nvdisasm -g), by instruction offset.

    python profiles/sass_dynamic.py report.ncu-rep all.sass 'kernel mangled-name pattern' envs_per_count [--lines N]
"""
import collections
import csv
import io
import re
import subprocess
import sys


def main():
    report, sass, pattern, units = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
    raw = subprocess.run(['ncu', '-i', report, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header = next(r for r in rows if r and r[0] == 'Address')
    counts = []
    for r in rows:
        if len(r) == len(header) and r[0].startswith('0x'):
            counts.append((int(r[0], 16), r[header.index('Source')].strip(), int(r[header.index('Instructions Executed')] or 0),
                           int(r[header.index('# Samples')] or 0)))
    base = counts[0][0]
    lines = open(sass).read().split('\n')
    start = next(i for i, l in enumerate(lines) if l.lstrip().startswith('.section') and '.text.' in l and pattern in l)
    table, current = {}, None
    for l in lines[start + 1:]:
        if l.lstrip().startswith('.section'):
            break
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            current = (m.group(1).split('/')[-1], int(m.group(2)))
            continue
        m = re.match(r'\s*/\*([0-9a-f]+)\*/\s+(.*?);', l)
        if m:
            table[int(m.group(1), 16)] = (current, m.group(2).strip())
    by_line, by_line_samples = collections.Counter(), collections.Counter()
    total = samples = 0
    mismatched = 0
    for address, text, n, smp in counts:
        entry = table.get(address - base)
        if entry is None or entry[1].split()[0].lstrip('@!UP0123456789 ') [:3] != text.split()[0].lstrip('@!UP0123456789 ')[:3]:
            mismatched += 1
        line = entry[0] if entry else None
        by_line[line] += n
        by_line_samples[line] += smp
        total += n
        samples += smp
    print(f'total {total} warp instructions = {total / units:.1f} per unit; {samples} samples; {mismatched} instructions did not line up')
    top = int(sys.argv[sys.argv.index('--lines') + 1]) if '--lines' in sys.argv else 60
    source = {}
    for (line, n) in by_line.most_common(top):
        if line is None:
            continue
        if line[0] not in source:
            try:
                path = subprocess.run(['find', '/root/repo/free_range_zoo_b200', '/usr/local/cuda/include', '-name', line[0]], capture_output=True, text=True).stdout.split('\n')[0]
                source[line[0]] = open(path).read().split('\n')
            except Exception:
                source[line[0]] = []
        text = source[line[0]][line[1] - 1].strip()[:90] if len(source[line[0]]) >= line[1] else ''
        print(f'{n / units:7.1f} {100 * n / total:5.1f}% inst {100 * by_line_samples[line] / max(samples, 1):5.1f}% smp  {line[0]}:{line[1]:<4d} {text}')
    if '--phases' in sys.argv:
        spec = sys.argv[sys.argv.index('--phases') + 1]  # name:first-last,name:first-last (lines of the main .cu file)
        for item in spec.split(','):
            name, span = item.split(':')
            first, last = (int(v) for v in span.split('-'))
            n = sum(v for (k, v) in by_line.items() if k and k[0].endswith('.cu') and first <= k[1] <= last)
            s = sum(v for (k, v) in by_line_samples.items() if k and k[0].endswith('.cu') and first <= k[1] <= last)
            print(f'{name:24s} {n / units:7.1f} {100 * n / total:5.1f}% inst {100 * s / max(samples, 1):5.1f}% smp')


main()
