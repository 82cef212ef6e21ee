#!/usr/bin/env python
"""SASS evidence for profiles/: from `cuobjdump -sass` of a kernel object, the mnemonic histogram of
one kernel, every bulk-copy / mbarrier instruction (UBLKCP, SYNCS) with context, and the window
of N consecutive instructions with the most FP64 instructions (the hottest fast-op handler).
    python scripts/extract_hot_sass.py <object.o> <kernel symbol substring> > profiles/rNN_sass_*.md"""
import collections
import re
import subprocess
import sys


def main():
    obj, sym = sys.argv[1:3]
    win = int(sys.argv[3]) if len(sys.argv) > 3 else 110
    txt = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout.split('\n')
    start = [i for i, l in enumerate(txt) if 'Function : ' in l and sym in l][0]
    end = next((i for i in range(start + 1, len(txt)) if 'Function : ' in txt[i]), len(txt))
    ins = []
    for l in txt[start:end]:
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
        if m:
            ins.append((m.group(1), m.group(2).strip()))
    def mnem(s):
        ff = s.split()
        if ff[0].startswith('@'):
            ff = ff[1:]
        return ff[0].split('.')[0]
    hist = collections.Counter(mnem(s) for _, s in ins)
    print(f'# SASS of `{txt[start].split("Function : ")[1].strip()}`\n')
    print(f'`cuobjdump -sass {obj.split("/")[-1]}`: {len(ins)} instructions (device functions '
          'that are not inlined are part of the listing).\n')
    print('| mnemonic | count |\n|---|---|')
    for k, v in hist.most_common(28):
        print(f'| {k} | {v} |')
    print('\n## Bulk async copies and mbarrier operations (tile pipeline)\n\n```')
    for i, (a, s) in enumerate(ins):
        if mnem(s) in ('UBLKCP', 'SYNCS'):
            print(f'/*{a}*/  {s}')
    print('```\n')
    fp = [1 if mnem(s) in ('DADD', 'DMUL', 'DFMA') else 0 for _, s in ins]
    acc = sum(fp[:win])
    best, bi = acc, 0
    for i in range(win, len(fp)):
        acc += fp[i] - fp[i - win]
        if acc > best:
            best, bi = acc, i - win + 1
    print(f'## Densest FP64 window: {best} DADD/DMUL/DFMA in {win} instructions '
          f'(a drift-prefixed fast-op handler of xtb_run_fast, three particles per thread)\n\n```')
    for a, s in ins[bi:bi + win]:
        print(f'/*{a}*/  {s}')
    print('```')


if __name__ == '__main__':
    main()
