#!/usr/bin/env python
"""Warp-stall samples and executed instructions of one kernel of an ncu report, aggregated by
the device FUNCTION (inlined or not) each SASS instruction came from.

    python scripts/ncu_by_function.py <report.ncu-rep> <object.o> <kernel symbol substring>

The report must have been taken with `--import-source on` on a library built with -lineinfo
from the SAME sources as <object.o> (the SASS of the two is matched instruction by
instruction).  Line info comes from `nvdisasm -g`.  Test / tuning tool, not product code."""
import bisect
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, obj, sym = sys.argv[1:4]
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr, body = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
    dis = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cubin)], capture_output=True,
                         text=True).stdout.split('\n')
    start = [i for i, l in enumerate(dis) if l.startswith('.text.') and sym in l][0]
    loc_re = re.compile(r'//## File "([^"]+)", line (\d+)')
    cur, locs = None, []
    for l in dis[start + 1:]:
        if l.startswith('//---------------------'):
            break
        m = loc_re.search(l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
        elif re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+', l):
            locs.append(cur)
    assert len(locs) == len(body), (len(locs), len(body), 'report and object differ')
    marks = {}
    csrc = os.path.join(ROOT, 'xtrack_b200', 'csrc')
    for f in os.listdir(csrc):
        if not f.endswith(('.cuh', '.cu')):
            continue
        mm = []
        text = open(os.path.join(csrc, f)).read().split('\n')
        for i, l in enumerate(text, 1):
            if ('__device__' in l or '__global__' in l) and not l.lstrip().startswith('//'):
                m = re.search(r'\b(\w+)\s*\([^;]*$', l)
                if m:
                    mm.append((i, m.group(1)))
                elif i < len(text):
                    m = re.match(r'\s*(\w+)\s*\(', text[i])
                    if m:
                        mm.append((i + 1, m.group(1)))
        marks[f] = mm

    def fn(loc):
        if loc is None:
            return '?'
        f, l = loc
        mm = marks.get(f)
        if not mm:
            return f
        k = bisect.bisect_right([m[0] for m in mm], l) - 1
        return f + ':' + (mm[k][1] if k >= 0 else '?')
    def category(sass):
        ff = sass.split()
        if ff and ff[0].startswith('@'):
            ff = ff[1:]
        m = ff[0].split('.')[0] if ff else '?'
        if m in ('DADD', 'DMUL', 'DFMA', 'DSETP'):
            return 'fp64'
        if m in ('MUFU',):
            return 'mufu'
        if m in ('MOV', 'IMAD', 'IADD3', 'LOP3', 'SHF', 'SEL', 'ISETP', 'LEA', 'PRMT', 'IADD', 'PLOP3', 'FSEL',
                 'UMOV', 'ULOP3', 'UIADD3', 'UISETP', 'USEL', 'USHF', 'ULEA', 'UIMAD', 'R2UR', 'S2R', 'CS2R',
                 'FSETP', 'FADD', 'FMUL', 'FFMA', 'I2F', 'F2I', 'F2F', 'IABS', 'IMNMX', 'VIADD', 'VIMNMX', 'P2R', 'R2P'):
            return 'int/mov'
        if m.startswith(('LD', 'ST', 'ATOM', 'RED', 'UBLKCP', 'SYNCS', 'ULD')):
            return 'mem'
        if m in ('BRA', 'BSSY', 'BSYNC', 'CALL', 'RET', 'EXIT', 'WARPSYNC', 'BAR', 'JMP', 'BRX', 'NOP', 'VOTE', 'VOTEU', 'SHFL', 'REDUX'):
            return 'ctrl'
        return 'other'
    mix = collections.defaultdict(collections.Counter)
    samples, execd = collections.Counter(), collections.Counter()
    stalls = collections.defaultdict(collections.Counter)
    scols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    for loc, r in zip(locs, body):
        k = fn(loc)
        samples[k] += int(r[ix['# Samples']])
        execd[k] += int(r[ix['Instructions Executed']])
        mix[k][category(r[ix['Source']])] += int(r[ix['Instructions Executed']])
        mix['(kernel)'][category(r[ix['Source']])] += int(r[ix['Instructions Executed']])
        for c in scols:
            stalls[k][c[6:]] += int(r[ix[c]])
    ts, te = sum(samples.values()), sum(execd.values())
    print(f'kernel {rows[0][1]}: {len(body)} SASS instructions, {ts} samples, {te:.3e} warp instructions')
    print('| function | samples | executed | top stalls |\n|---|---|---|---|')
    for k, v in samples.most_common(24):
        top = ', '.join(f'{n} {100 * c / max(v, 1):.0f} %' for n, c in stalls[k].most_common(3))
        print(f'| {k} | {100 * v / ts:.1f} % | {100 * execd[k] / te:.1f} % | {top} |')
    print('\nexecuted warp instructions by kind (share of the function\'s own):\n')
    cats = ['fp64', 'mufu', 'int/mov', 'mem', 'ctrl', 'other']
    print('| function | ' + ' | '.join(cats) + ' |\n|---|' + '---|' * len(cats))
    for k in ['(kernel)'] + [k for k, _ in samples.most_common(16)]:
        tot = max(sum(mix[k].values()), 1)
        print(f'| {k} | ' + ' | '.join(f'{100 * mix[k][c] / tot:.0f} %' for c in cats) + ' |')


if __name__ == '__main__':
    main()
