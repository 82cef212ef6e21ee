#!/usr/bin/env python
"""Summarises `ncu --set full --import-source on` captures of the tracking kernel into the
tracked files under profiles/ (run HERE, after `gpurun` merged the .ncu-rep files back):

    python scripts/summarize_ncu.py gpurun_out/<tag> [--round r01]

Reads <tag>/prof_track.ncu-rep (EXACT), <tag>/prof_track--fma.ncu-rep (FMA), optionally
<tag>/prof_lep.ncu-rep (thick kernel), through `ncu -i ... --page raw|source --csv`, and writes
profiles/<round>_ncu_track.md, profiles/<round>_ncu_summary.json, profiles/<round>_ncu_lep.md.
"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

RAW_KEYS = [
    ('gpu__time_duration.sum', 'gpu__time_duration [ms]'),
    ('launch__registers_per_thread', 'registers / thread'),
    ('launch__block_size', 'block size'),
    ('launch__grid_size', 'grid size'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__warps_active [% of 64]'),
    ('smsp__inst_executed.sum', 'smsp__inst_executed (warp instructions)'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active [%]'),
    ('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
     '**sm__inst_executed_pipe_fp64 (FP64 pipe utilisation) [%]**'),
    ('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'pipe alu [%]'),
    ('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'pipe fma [%]'),
    ('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'pipe lsu [%]'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'shared-memory bank conflicts'),
    ('derived__memory_l1_wavefronts_shared_excessive', 'excessive shared wavefronts'),
    ('smsp__inst_executed_op_local_ld.sum', 'local loads (spills + thread-local state)'),
    ('smsp__inst_executed_op_local_st.sum', 'local stores'),
    ('dram__bytes_read.sum', 'dram__bytes_read [MB]'),
    ('dram__bytes_write.sum', 'dram__bytes_write [MB]'),
    ('l1tex__t_sector_hit_rate.pct', 'L1 hit rate [%]'),
]


def ncu_csv(rep, page):
    out = subprocess.run(['ncu', '-i', rep, '--page', page, '--csv'] +
                         (['--print-source', 'sass'] if page == 'source' else []),
                         capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def raw_metrics(rep):
    rows = ncu_csv(rep, 'raw')
    hdr, units, vals = rows[0], rows[1], rows[-1]
    return dict(zip(hdr, vals)), dict(zip(hdr, units))


def stall_profile(rep):
    rows = ncu_csv(rep, 'source')
    kernel = rows[0][1] if rows and len(rows[0]) > 1 else ''
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    agg = collections.Counter()
    byop = collections.Counter()
    byop_inst = collections.Counter()
    tot = 0
    for r in rows[2:]:
        try:
            n = int(r[idx['# Samples']])
        except (ValueError, IndexError):
            continue
        tot += n
        for h in stalls:
            agg[h[6:]] += int(r[idx[h]] or 0)
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)', r[idx['Source']])
        op = m.group(2) if m else '?'
        byop[op] += n
        byop_inst[op] += int(r[idx['Instructions Executed']] or 0)
    return kernel, tot, agg, byop, byop_inst


def section(title, rep, pet_per_launch=None):
    d, units = raw_metrics(rep)
    kernel, tot, agg, byop, byop_inst = stall_profile(rep)
    L = [f'## {title}', '', f'kernel `{kernel}`', '', '| metric | value |', '|---|---|']
    for key, label in RAW_KEYS:
        if key in d:
            L.append(f'| {label} | {d[key]} {units.get(key, "")} |')
    fp64 = sum(byop_inst[o] for o in ('DADD', 'DMUL', 'DFMA', 'DSETP'))
    L.append(f'| FP64 warp instructions (DADD+DMUL+DFMA+DSETP, from the source view) | {fp64:.4e} |')
    summary = {'kernel': kernel, 'gpu_time_ms': float(d.get('gpu__time_duration.sum', 'nan')),
               'fp64_pipe_pct': float(d.get(
                   'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'nan')),
               'registers': int(float(d.get('launch__registers_per_thread', 0))),
               'warp_instructions': float(d.get('smsp__inst_executed.sum', 'nan')),
               'fp64_warp_instructions': float(fp64)}
    try:
        mb = float(d['dram__bytes_read.sum']) + float(d['dram__bytes_write.sum'])
        unit = units.get('dram__bytes_read.sum', 'Mbyte').lower()
        scale = {'byte': 1.0, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}.get(unit, 1e6)
        summary['dram_bytes_per_launch'] = mb * scale
    except (KeyError, ValueError):
        pass
    if pet_per_launch:
        warp_pet = pet_per_launch / 32.0
        summary['fp64_inst_per_pet'] = fp64 / warp_pet
        summary['inst_per_pet'] = summary['warp_instructions'] / warp_pet
        summary['pet_per_s_under_ncu'] = pet_per_launch / (summary['gpu_time_ms'] * 1e-3)
        L.append(f'| per particle-element-turn | {summary["inst_per_pet"]:.1f} instructions, '
                 f'{summary["fp64_inst_per_pet"]:.1f} of them FP64; '
                 f'{summary["pet_per_s_under_ncu"]:.3e} PET/s (ncu: cold, serialised) |')
    L += ['', f'warp stall samples ({tot} in total), share of all samples:', '',
          '| ' + ' | '.join(k for k, _ in agg.most_common(9)) + ' |',
          '|' + '---|' * min(9, len(agg)),
          '| ' + ' | '.join('%.1f %%' % (100.0 * v / max(tot, 1)) for _, v in agg.most_common(9)) + ' |',
          '', 'samples by opcode: ' + ', '.join(
              '%s %.1f %%' % (o, 100.0 * v / max(tot, 1)) for o, v in byop.most_common(8)), '']
    return L, summary


def main():
    tag = sys.argv[1]
    rnd = sys.argv[sys.argv.index('--round') + 1] if '--round' in sys.argv else 'r01'
    n_el = 11843
    particles = 1_000_000
    out = [f'# Round {rnd[1:]} — `ncu --set full` summary of the tracking kernel (B200, sm_100a)', '',
           'Command (under gpurun, one GPU): `ncu --set full --clock-control none --import-source on '
           '-k regex:xtb_track_kernel -s 2 -c 1 python bench.py --quick --steps 1 --warmup 1 '
           '--turns 3 [--fma]` (`Line.track(num_turns=3)` issues a 1-turn and a 2-turn launch, '
           'tracker.py:1372-1413; the captured launch is ONE turn of hllhc_14 over 10^6 particles; '
           f'source: `{tag}/prof_track*.ncu-rep`, summarised by `scripts/summarize_ncu.py`).', '']
    summ = {}
    for label, fn, key in (('EXACT (-fmad=false, default)', 'prof_track.ncu-rep', 'exact'),
                           ('FMA (-fmad=true)', 'prof_track--fma.ncu-rep', 'fma')):
        rep = os.path.join(tag, fn)
        if not os.path.exists(rep):
            continue
        L, s = section(label, rep, pet_per_launch=1.0 * n_el * particles)
        s['source'] = f'{rep} (ncu --set full, one turn of hllhc_14, 1e6 particles)'
        out += L
        summ[key] = s
    if summ:
        top = dict(summ.get('exact', next(iter(summ.values()))))
        top['variants'] = summ
        with open(os.path.join(ROOT, 'profiles', f'{rnd}_ncu_summary.json'), 'w') as fid:
            json.dump(top, fid, indent=1)
        with open(os.path.join(ROOT, 'profiles', f'{rnd}_ncu_track.md'), 'w') as fid:
            fid.write('\n'.join(out) + '\n')
        print('wrote', f'profiles/{rnd}_ncu_track.md', f'profiles/{rnd}_ncu_summary.json')
    rep = os.path.join(tag, 'prof_lep.ncu-rep')
    if os.path.exists(rep):
        L, s = section('Thick kernel (LEP stand-in, 300 000 particles, one turn, EXACT)', rep,
                       pet_per_launch=1.0 * 9230 * 300000)
        with open(os.path.join(ROOT, 'profiles', f'{rnd}_ncu_lep.md'), 'w') as fid:
            fid.write(f'# Round {rnd[1:]} — `ncu --set full` summary of the thick-magnet kernel\n\n'
                      '`ncu --set full --clock-control none --import-source on -k '
                      'regex:xtb_track_kernel -s 1 -c 1 python bench.py --workload lep_thick '
                      f'--particles 300000 --quick --steps 1 --warmup 1 --turns 2` ({rep}).\n\n'
                      + '\n'.join(L) + '\n')
        print('wrote', f'profiles/{rnd}_ncu_lep.md')


if __name__ == '__main__':
    main()
