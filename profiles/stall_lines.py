"""Per-source-line stall breakdown of one `ncu --set full --import-source on` report (needs -lineinfo):

    python profiles/stall_lines.py gpurun_out/<report>.ncu-rep [top_n] [file-substring]

Lists the lines with the most warp-stall samples and, for each, the dominant stall reasons; sorted by samples (where
warps WAIT), unlike summarize.py's list, which is sorted by executed instructions (what warps ISSUE)."""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def main(path, top=30, only=''):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--launch-count', '1'],
                         capture_output=True, text=True).stdout
    current, head = None, None
    per_line = defaultdict(lambda: defaultdict(float))
    text = {}
    for record in csv.reader(io.StringIO(out)):
        if not record:
            continue
        if record[0] == 'File Path':
            current = record[1].split('/')[-1]
        elif record[0] == 'Line No':
            head = record
        elif head and record[0].isdigit() and current:
            key = (current, int(record[0]))
            text[key] = record[1].strip()
            for at, name in enumerate(head):
                if at < len(record) and (name in ('# Samples', 'Instructions Executed') or (name.startswith('stall_') and 'Not Issued' not in name)):
                    try:
                        per_line[key][name] += float(record[at])
                    except ValueError:
                        pass
    total = sum(v['# Samples'] for v in per_line.values()) or 1
    reasons = defaultdict(float)
    for v in per_line.values():
        for name, value in v.items():
            if name.startswith('stall_'):
                reasons[name] += value
    print('# all lines:', ', '.join(f'{n[6:]} {100 * x / total:.1f}%' for n, x in sorted(reasons.items(), key=lambda kv: -kv[1])[:10]))
    for key, v in sorted(per_line.items(), key=lambda kv: -kv[1]['# Samples'])[:top]:
        if only and only not in key[0]:
            continue
        why = ', '.join(f'{n[6:]} {100 * x / max(v["# Samples"], 1):.0f}%' for n, x in
                        sorted(((n, x) for n, x in v.items() if n.startswith('stall_')), key=lambda kv: -kv[1])[:3] if x > 0)
        print(f'{100 * v["# Samples"] / total:5.1f}% smp {int(v["Instructions Executed"]):>9d} inst  {key[0]}:{key[1]:<4d} [{why}]  {text[key][:90]}')


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30, sys.argv[3] if len(sys.argv) > 3 else '')
