"""Static SASS statistics of one kernel: instructions per source line and per opcode.

    cuobjdump -xelf all libfrz.so; nvdisasm -g frz_wildfire.sm_100a.cubin > all.sass
    python profiles/sass_lines.py all.sass 'wildfire_step_kernelILi16ELi7ELi0ELb0' [top]
"""
import collections
import re
import sys


def main():
    path, pattern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    lines = open(path).read().split('\n')
    start = next(i for i, l in enumerate(lines) if l.lstrip().startswith('.section') and '.text.' in l and pattern in l)
    by_line, by_op = collections.Counter(), collections.Counter()
    current, total = None, 0
    for l in lines[start + 1:]:
        if l.lstrip().startswith('.section'):
            break
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            current = (m.group(1).split('/')[-1], int(m.group(2)))
            continue
        m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)', l)
        if m:
            by_line[current] += 1
            by_op[m.group(2).split('.')[0]] += 1
            total += 1
    print('total static instructions', total)
    print('--- by opcode')
    for op, n in by_op.most_common(30):
        print(f'{n:6d} {100 * n / total:5.1f}%  {op}')
    print('--- by source line')
    for (f, ln), n in by_line.most_common(top):
        print(f'{n:6d} {100 * n / total:5.1f}%  {f}:{ln}')


main()
