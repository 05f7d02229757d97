"""Summarise an ncu report per CUDA source line: instructions executed and stall samples.

    python profiles/ncu_lines.py gpurun_out/<report>.ncu-rep [top_n]
"""
import csv
import io
import subprocess
import sys


def main():
    report = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(['ncu', '-i', report, '--page', 'source', '--csv', '--print-source', 'cuda,sass',
                          '--launch-count', '1'], capture_output=True, text=True).stdout
    rows, current_file = [], None
    header = None
    for record in csv.reader(io.StringIO(out)):
        if not record:
            continue
        if record[0] == 'File Path':
            current_file = record[1].split('/')[-1]
        elif record[0] == 'Line No':
            header = record
        elif header and record[0].isdigit():
            def get(name):
                value = record[header.index(name)]
                return int(value) if value.isdigit() else 0

            rows.append((current_file, int(record[0]), record[1].strip(), get('Instructions Executed'), get('# Samples')))
    total_inst = sum(r[3] for r in rows) or 1
    total_samples = sum(r[4] for r in rows) or 1
    print(f'total warp instructions {total_inst}, stall samples {total_samples}')
    for name, line, text, inst, samples in sorted(rows, key=lambda r: -r[3])[:top]:
        print(f'{100 * inst / total_inst:5.1f}% inst {100 * samples / total_samples:5.1f}% smp  {name}:{line:<4d} {text[:110]}')


if __name__ == '__main__':
    main()
