"""ORACLE (test infrastructure) -- generator of the xobjects-side C API.

The reference's element physics (`/root/reference/xtrack/**/*.h`) is written
against an accessor API that the external `xobjects` package generates at JIT
time and that therefore does not exist in the reference tree:

  * `LocalParticle` struct + `LocalParticle_{get,set,add_to,scale}_<field>`,
    `LocalParticle_exchange`, `Particles_to_LocalParticle`,
    `LocalParticle_to_Particles`       (spec: xtrack/particles/particles.py:86-408)
  * `ParticlesData_*` accessors        (call sites particles.py:143-181)
  * `<Element>Data` structs/accessors  (call sites in elements_src/*.h)
  * `ParticlesMonitorData_*`, `LastTurnsMonitorData_*`, `LastTurnsData_*`
                                       (monitors/*.h)
  * per-class `*_track_local_particle_with_transformations`
                                       (base_element.py:83-127 -> the template
                                        headers/track_local_particle_with_transformations.h)

This script writes that API as plain C over plain structs into
`oracle/_ref/gen/xt_generated.h`.  It is *our* restatement of generated glue;
the physics headers it includes are compiled unmodified from /root/reference.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

from element_specs import SPECS, CLASS_ORDER, TYPE_ID, all_fields  # noqa: E402

SIZE_VARS = [('int64_t', '_capacity'), ('int64_t', '_num_active_particles'),
             ('int64_t', '_num_lost_particles'), ('int64_t', 'start_tracking_at_element')]
SCALAR_VARS = [('double', 'q0'), ('double', 'mass0'), ('double', 't_sim')]
PER_PARTICLE = (
    [('double', nn) for nn in (
        'p0c', 'gamma0', 'beta0', 's', 'zeta', 'x', 'y', 'px', 'py', 'ptau', 'delta',
        'rpp', 'rvv', 'chi', 'charge_ratio', 'weight', 'ax', 'ay', 'spin_x',
        'spin_y', 'spin_z', 'anomalous_magnetic_moment')]
    + [('int64_t', nn) for nn in ('pdg_id', 'particle_id', 'at_element', 'at_turn',
                                  'state', 'parent_particle_id')]
    + [('uint32_t', nn) for nn in ('_rng_s1', '_rng_s2', '_rng_s3', '_rng_s4')])

CTYPE = {'f64': 'double', 'i64': 'int64_t'}


def gen_particles_api():
    L = []
    L.append('/* ---- ParticlesData: SoA of pointers (layout chosen by the oracle) ---- */')
    L.append('typedef struct ParticlesData_s {')
    for tt, vv in SIZE_VARS + SCALAR_VARS:
        L.append(f'    {tt} {vv};')
    for tt, vv in PER_PARTICLE:
        L.append(f'    {tt}* {vv};')
    L.append('} *ParticlesData;')
    for tt, vv in SIZE_VARS + SCALAR_VARS:
        L.append(f'GPUFUN {tt} ParticlesData_get_{vv}(ParticlesData p){{ return p->{vv}; }}')
        L.append(f'GPUFUN void ParticlesData_set_{vv}(ParticlesData p, {tt} v){{ p->{vv} = v; }}')
    for tt, vv in PER_PARTICLE:
        L.append(f'GPUFUN {tt}* ParticlesData_getp1_{vv}(ParticlesData p, int64_t i){{ return p->{vv} + i; }}')
        L.append(f'GPUFUN {tt} ParticlesData_get_{vv}(ParticlesData p, int64_t i){{ return p->{vv}[i]; }}')
        L.append(f'GPUFUN void ParticlesData_set_{vv}(ParticlesData p, int64_t i, {tt} v){{ p->{vv}[i] = v; }}')

    L.append('/* ---- LocalParticle (particles.py:96-109) ---- */')
    L.append('typedef struct {')
    for tt, vv in SIZE_VARS + SCALAR_VARS:
        L.append(f'    {tt} {vv};')
    for tt, vv in PER_PARTICLE:
        L.append(f'    {tt}* {vv};')
    L += ['    int64_t ipart;', '    int64_t endpart;', '    uint64_t track_flags;',
          '    double line_length;', '    int8_t* io_buffer;', '} LocalParticle;']
    L.append('GPUFUN int8_t* LocalParticle_get_io_buffer(LocalParticle* part){ return part->io_buffer; }')
    L.append('GPUFUN uint64_t LocalParticle_check_track_flag(LocalParticle* part, uint8_t index){'
             ' return (part->track_flags >> index) & 1; }')
    for tt, vv in SIZE_VARS + SCALAR_VARS:
        L.append(f'GPUFUN {tt} LocalParticle_get_{vv}(LocalParticle* part){{ return part->{vv}; }}')
    for tt, vv in PER_PARTICLE:
        L.append(f'GPUFUN {tt} LocalParticle_get_{vv}(LocalParticle* part){{ return part->{vv}[part->ipart]; }}')
        for op, sym in (('set', '='), ('add_to', '+='), ('scale', '*=')):
            L.append(f'GPUFUN void LocalParticle_{op}_{vv}(LocalParticle* part, {tt} value){{')
            L.append(f'#ifndef FREEZE_VAR_{vv}')
            L.append(f'    part->{vv}[part->ipart] {sym} value;')
            L.append('#endif')
            L.append('}')
    L.append('GPUFUN void LocalParticle_exchange(LocalParticle* part, int64_t i1, int64_t i2){')
    for tt, vv in PER_PARTICLE:
        L.append(f'    {{ {tt} temp = part->{vv}[i2]; part->{vv}[i2] = part->{vv}[i1]; part->{vv}[i1] = temp; }}')
    L.append('}')
    L.append('GPUFUN void Particles_to_LocalParticle(ParticlesData source, LocalParticle* dest, int64_t id, int64_t eid){')
    for tt, vv in SIZE_VARS + SCALAR_VARS:
        L.append(f'    dest->{vv} = ParticlesData_get_{vv}(source);')
    for tt, vv in PER_PARTICLE:
        L.append(f'    dest->{vv} = ParticlesData_getp1_{vv}(source, 0);')
    L += ['    dest->ipart = id;', '    dest->endpart = eid;', '}']
    L.append('GPUFUN void LocalParticle_to_Particles(LocalParticle* source, ParticlesData dest, int64_t id, int64_t set_scalar){')
    L.append('    if (set_scalar){')
    for tt, vv in SIZE_VARS + SCALAR_VARS:
        L.append(f'        ParticlesData_set_{vv}(dest, LocalParticle_get_{vv}(source));')
    L.append('    }')
    for tt, vv in PER_PARTICLE:
        L.append(f'    ParticlesData_set_{vv}(dest, id, LocalParticle_get_{vv}(source));')
    L.append('}')
    return '\n'.join(L)


def gen_slice_struct(name):
    """Slice: `_parent` reference + own fields; `XData_get__parent_<f>` forwards to the parent's
    accessor (what xobjects generates for an `xo.Ref` field)."""
    spec = SPECS[name]
    parent = spec['parent']
    L = [f'/* ---- {name}Data (slice of {parent}) ---- */', f'typedef struct {name}Data_s {{',
         f'    {parent}Data _parent;']
    for fn, kind in spec['fields']:
        L.append(f'    {CTYPE[kind]} {fn};')
    L.append(f'}} *{name}Data;')
    for fn, kind in spec['fields']:
        L.append(f'GPUFUN {CTYPE[kind]} {name}Data_get_{fn}({name}Data el){{ return el->{fn}; }}')
    for fn, kind in all_fields(parent):
        if kind == 'arr':
            L.append(f'GPUFUN double {name}Data_get__parent_{fn}({name}Data el, int64_t i){{ return el->_parent->{fn}[i]; }}')
            L.append(f'GPUFUN double* {name}Data_getp1__parent_{fn}({name}Data el, int64_t i){{ return el->_parent->{fn} + i; }}')
            L.append(f'GPUFUN int64_t {name}Data_len__parent_{fn}({name}Data el){{ return el->_parent->{fn}__len; }}')
        else:
            L.append(f'GPUFUN {CTYPE[kind]} {name}Data_get__parent_{fn}({name}Data el){{ return el->_parent->{fn}; }}')
    return '\n'.join(L)


def gen_element_struct(name):
    if 'parent' in SPECS[name]:
        return gen_slice_struct(name)
    L = [f'/* ---- {name}Data ---- */', f'typedef struct {name}Data_s {{']
    ff = all_fields(name)
    for fn, kind in ff:
        if kind == 'arr':
            L.append(f'    double* {fn}; int64_t {fn}__len;')
        else:
            L.append(f'    {CTYPE[kind]} {fn};')
    L.append(f'}} *{name}Data;')
    for fn, kind in ff:
        if kind == 'arr':
            L.append(f'GPUFUN double {name}Data_get_{fn}({name}Data el, int64_t i){{ return el->{fn}[i]; }}')
            L.append(f'GPUFUN double* {name}Data_getp1_{fn}({name}Data el, int64_t i){{ return el->{fn} + i; }}')
            L.append(f'GPUFUN int64_t {name}Data_len_{fn}({name}Data el){{ return el->{fn}__len; }}')
        else:
            L.append(f'GPUFUN {CTYPE[kind]} {name}Data_get_{fn}({name}Data el){{ return el->{fn}; }}')
    if SPECS[name].get('internal_record'):
        L.append(f'#define {name}Data_getp_internal_record(el, part) (NULL)')
    return '\n'.join(L)


def gen_with_transformations(name):
    """Mirror of base_element.py:83-127."""
    spec = SPECS[name]
    opts = [('ELEMENT_NAME', name)]
    if spec.get('rot_shift') or spec.get('rot_shift_from_parent'):
        opts.append(('ALLOW_ROT_AND_SHIFT', 1))
    if spec.get('rot_shift_from_parent'):
        opts.append(('IS_SLICE', 1))
    if spec.get('curved') and (spec.get('rot_shift_from_parent') or 'parent' not in spec):
        opts.append(('CURVED', 1))
    if spec.get('thin_slice') and spec.get('curved'):
        opts.append(('THIN_SLICE_OF_CURVED_ELEMENT', 1))
    if spec.get('dyn_thick'):
        opts.append(('IS_THICK_DYNAMIC', 1))
    elif spec.get('isthick'):
        opts.append(('IS_THICK', 1))
    L = [f'#define {k} {v}' for k, v in opts]
    L.append('#include "xtrack/headers/track_local_particle_with_transformations.h"')
    L += [f'#undef {k}' for k, _ in opts]
    return '\n'.join(L)


MONITOR_API = r'''
/* ---- ParticlesMonitorData (monitors/particles_monitor.py:180-192) ---- */
typedef struct ParticlesMonitorData_s {
    int64_t start_at_turn, stop_at_turn, part_id_start, part_id_end, ebe_mode,
            n_records, n_repetitions, repetition_period, flag_auto_to_numpy;
    ParticlesData data;
} *ParticlesMonitorData;
#define XTB_MON_GET(f) GPUFUN int64_t ParticlesMonitorData_get_##f(ParticlesMonitorData el){ return el->f; }
XTB_MON_GET(start_at_turn) XTB_MON_GET(stop_at_turn) XTB_MON_GET(part_id_start)
XTB_MON_GET(part_id_end) XTB_MON_GET(ebe_mode) XTB_MON_GET(n_repetitions)
XTB_MON_GET(repetition_period)
GPUFUN ParticlesData ParticlesMonitorData_getp_data(ParticlesMonitorData el){ return el->data; }

/* ---- LastTurnsMonitorData (monitors/last_turns_monitor.py:18-44) ---- */
typedef struct LastTurnsData_s {
    uint32_t* lost_at_offset; uint32_t* particle_id; uint32_t* at_turn;
    float *x, *px, *y, *py, *delta, *zeta;
} *LastTurnsData;
typedef struct LastTurnsMonitorData_s {
    int64_t particle_id_start, num_particles, n_last_turns, every_n_turns;
    LastTurnsData data;
} *LastTurnsMonitorData;
#define XTB_LTM_GET(f) GPUFUN int64_t LastTurnsMonitorData_get_##f(LastTurnsMonitorData el){ return el->f; }
XTB_LTM_GET(particle_id_start) XTB_LTM_GET(num_particles) XTB_LTM_GET(n_last_turns)
XTB_LTM_GET(every_n_turns)
GPUFUN LastTurnsData LastTurnsMonitorData_getp_data(LastTurnsMonitorData el){ return el->data; }
#define XTB_LTD_SET(f, T) GPUFUN void LastTurnsData_set_##f(LastTurnsData d, int64_t i, T v){ d->f[i] = v; }
XTB_LTD_SET(lost_at_offset, uint32_t) XTB_LTD_SET(particle_id, uint32_t) XTB_LTD_SET(at_turn, uint32_t)
XTB_LTD_SET(x, float) XTB_LTD_SET(px, float) XTB_LTD_SET(y, float) XTB_LTD_SET(py, float)
XTB_LTD_SET(delta, float) XTB_LTD_SET(zeta, float)
'''

BEAM_MONITOR_API = r'''
/* ---- BeamPositionMonitorData / BeamSizeMonitorData (monitors/beam_position_monitor.py:18-36,
 *      beam_size_monitor.py:18-38): same layout, the size monitor has two more sums ---- */
typedef struct XtbBeamRecord_s { int64_t n; double *count, *x_sum, *y_sum, *x2_sum, *y2_sum; } *XtbBeamRecord;
typedef struct XtbBeamMonitorData_s {
    int64_t particle_id_start, num_particles, start_at_turn, stop_at_turn;
    double frev, sampling_frequency;
    XtbBeamRecord data;
} *XtbBeamMonitorData;
#define XTB_BEAMMON_API(CLS) \
typedef XtbBeamMonitorData CLS##Data; typedef XtbBeamRecord CLS##Record; \
GPUFUN int64_t CLS##Data_get_start_at_turn(CLS##Data el){ return el->start_at_turn; } \
GPUFUN int64_t CLS##Data_get_particle_id_start(CLS##Data el){ return el->particle_id_start; } \
GPUFUN int64_t CLS##Data_get_num_particles(CLS##Data el){ return el->num_particles; } \
GPUFUN double CLS##Data_get_frev(CLS##Data el){ return el->frev; } \
GPUFUN double CLS##Data_get_sampling_frequency(CLS##Data el){ return el->sampling_frequency; } \
GPUFUN CLS##Record CLS##Data_getp_data(CLS##Data el){ return el->data; } \
GPUFUN int64_t CLS##Record_len_count(CLS##Record r){ return r->n; } \
GPUFUN double* CLS##Record_getp1_count(CLS##Record r, int64_t i){ return r->count + i; } \
GPUFUN double* CLS##Record_getp1_x_sum(CLS##Record r, int64_t i){ return r->x_sum + i; } \
GPUFUN double* CLS##Record_getp1_y_sum(CLS##Record r, int64_t i){ return r->y_sum + i; } \
GPUFUN double* CLS##Record_getp1_x2_sum(CLS##Record r, int64_t i){ return r->x2_sum + i; } \
GPUFUN double* CLS##Record_getp1_y2_sum(CLS##Record r, int64_t i){ return r->y2_sum + i; }
XTB_BEAMMON_API(BeamPositionMonitor)
XTB_BEAMMON_API(BeamSizeMonitor)

/* ---- BeamProfileMonitorData (monitors/beam_profile_monitor.py:20-37) ---- */
typedef struct BeamProfileMonitorRecord_s { int64_t n_x, n_y; double *counts_x, *counts_y; } *BeamProfileMonitorRecord;
typedef struct BeamProfileMonitorData_s {
    int64_t particle_id_start, num_particles, start_at_turn, stop_at_turn, nx, ny, sample_size;
    double frev, sampling_frequency, x_min, dx, y_min, dy;
    BeamProfileMonitorRecord data;
} *BeamProfileMonitorData;
#define XTB_BPROF_I(f) GPUFUN int64_t BeamProfileMonitorData_get_##f(BeamProfileMonitorData el){ return el->f; }
#define XTB_BPROF_D(f) GPUFUN double BeamProfileMonitorData_get_##f(BeamProfileMonitorData el){ return el->f; }
XTB_BPROF_I(particle_id_start) XTB_BPROF_I(num_particles) XTB_BPROF_I(start_at_turn)
XTB_BPROF_I(nx) XTB_BPROF_I(ny) XTB_BPROF_I(sample_size)
XTB_BPROF_D(frev) XTB_BPROF_D(sampling_frequency) XTB_BPROF_D(x_min) XTB_BPROF_D(dx)
XTB_BPROF_D(y_min) XTB_BPROF_D(dy)
GPUFUN BeamProfileMonitorRecord BeamProfileMonitorData_getp_data(BeamProfileMonitorData el){ return el->data; }
GPUFUN int64_t BeamProfileMonitorRecord_len_counts_x(BeamProfileMonitorRecord r){ return r->n_x; }
GPUFUN int64_t BeamProfileMonitorRecord_len_counts_y(BeamProfileMonitorRecord r){ return r->n_y; }
GPUFUN double* BeamProfileMonitorRecord_getp1_counts_x(BeamProfileMonitorRecord r, int64_t i){ return r->counts_x + i; }
GPUFUN double* BeamProfileMonitorRecord_getp1_counts_y(BeamProfileMonitorRecord r, int64_t i){ return r->counts_y + i; }

/* ---- BeamStatsMonitorData (monitors/beam_stats_monitor/beam_stats_monitor.py:60-118, 240-256):
 *      the scalars, the slot tables, the record of 38 moment arrays (length 0: not kept), the
 *      touched-record flags and the profile record, as one flat struct; the nested xobjects
 *      the reference's header asks for are views of it ---- */
typedef struct BeamStatsMonitorData_s {
    int64_t start_at_turn, stop_at_turn, every_n_turns, _mode, _num_records, _num_selected_slots,
            _num_slices, _particle_id_start, _particle_id_stop;
    double _z_min_edge, _dzeta, _bunch_spacing_zeta;
    int64_t n_slot_to_selected;
    int64_t *_slot_to_selected, *_selected_slots;
    double* field[38];
    int64_t len_field[38];
    int64_t* touched;
    int64_t n_profiles, len_counts;
    double* counts;
    int64_t *offsets, *num_bins, *coord_id;
    double *pmin, *bin_width;
} *BeamStatsMonitorData;
typedef BeamStatsMonitorData BeamStatsMonitorRecordData;
typedef BeamStatsMonitorData BeamStatsMonitorTouchedRecordsData;
typedef BeamStatsMonitorData BeamStatsMonitorProfileRecordData;
#define XTB_BSM_I(f) GPUFUN int64_t BeamStatsMonitorData_get_##f(BeamStatsMonitorData el){ return el->f; }
#define XTB_BSM_D(f) GPUFUN double BeamStatsMonitorData_get_##f(BeamStatsMonitorData el){ return el->f; }
XTB_BSM_I(start_at_turn) XTB_BSM_I(stop_at_turn) XTB_BSM_I(every_n_turns) XTB_BSM_I(_mode)
XTB_BSM_I(_num_records) XTB_BSM_I(_num_selected_slots) XTB_BSM_I(_num_slices)
XTB_BSM_I(_particle_id_start) XTB_BSM_I(_particle_id_stop)
XTB_BSM_D(_z_min_edge) XTB_BSM_D(_dzeta) XTB_BSM_D(_bunch_spacing_zeta)
GPUFUN int64_t BeamStatsMonitorData_len__slot_to_selected(BeamStatsMonitorData el){ return el->n_slot_to_selected; }
GPUFUN int64_t BeamStatsMonitorData_get__slot_to_selected(BeamStatsMonitorData el, int64_t i){ return el->_slot_to_selected[i]; }
GPUFUN int64_t BeamStatsMonitorData_get__selected_slots(BeamStatsMonitorData el, int64_t i){ return el->_selected_slots[i]; }
GPUFUN BeamStatsMonitorRecordData BeamStatsMonitorData_getp_data(BeamStatsMonitorData el){ return el; }
GPUFUN BeamStatsMonitorTouchedRecordsData BeamStatsMonitorData_getp_touched_records(BeamStatsMonitorData el){ return el; }
GPUFUN BeamStatsMonitorProfileRecordData BeamStatsMonitorData_getp__profile_data(BeamStatsMonitorData el){ return el; }
#define XTB_BSM_FIELD(f, i) \
GPUFUN int64_t BeamStatsMonitorRecordData_len_##f(BeamStatsMonitorRecordData r){ return r->len_field[i]; } \
GPUFUN double* BeamStatsMonitorRecordData_getp1_##f(BeamStatsMonitorRecordData r, int64_t k){ return r->field[i] + k; }
XTB_BSM_FIELD(num_particles, 0) XTB_BSM_FIELD(sum_beta0_gamma0, 1) XTB_BSM_FIELD(sum_x, 2) 
XTB_BSM_FIELD(sum_px, 3) XTB_BSM_FIELD(sum_y, 4) XTB_BSM_FIELD(sum_py, 5) 
XTB_BSM_FIELD(sum_zeta, 6) XTB_BSM_FIELD(sum_delta, 7) XTB_BSM_FIELD(sum_pzeta, 8) 
XTB_BSM_FIELD(sum_charge_ratio, 9) XTB_BSM_FIELD(sum_mass_ratio, 10) XTB_BSM_FIELD(sum_x_x, 11) 
XTB_BSM_FIELD(sum_x_px, 12) XTB_BSM_FIELD(sum_x_y, 13) XTB_BSM_FIELD(sum_x_py, 14) 
XTB_BSM_FIELD(sum_x_zeta, 15) XTB_BSM_FIELD(sum_x_delta, 16) XTB_BSM_FIELD(sum_x_pzeta, 17) 
XTB_BSM_FIELD(sum_px_px, 18) XTB_BSM_FIELD(sum_px_y, 19) XTB_BSM_FIELD(sum_px_py, 20) 
XTB_BSM_FIELD(sum_px_zeta, 21) XTB_BSM_FIELD(sum_px_delta, 22) XTB_BSM_FIELD(sum_px_pzeta, 23) 
XTB_BSM_FIELD(sum_y_y, 24) XTB_BSM_FIELD(sum_y_py, 25) XTB_BSM_FIELD(sum_y_zeta, 26) 
XTB_BSM_FIELD(sum_y_delta, 27) XTB_BSM_FIELD(sum_y_pzeta, 28) XTB_BSM_FIELD(sum_py_py, 29) 
XTB_BSM_FIELD(sum_py_zeta, 30) XTB_BSM_FIELD(sum_py_delta, 31) XTB_BSM_FIELD(sum_py_pzeta, 32) 
XTB_BSM_FIELD(sum_zeta_zeta, 33) XTB_BSM_FIELD(sum_zeta_delta, 34) XTB_BSM_FIELD(sum_zeta_pzeta, 35) 
XTB_BSM_FIELD(sum_delta_delta, 36) XTB_BSM_FIELD(sum_pzeta_pzeta, 37) 
GPUFUN int64_t* BeamStatsMonitorTouchedRecordsData_getp1_value(BeamStatsMonitorTouchedRecordsData r, int64_t k){ return r->touched + k; }
GPUFUN int64_t BeamStatsMonitorProfileRecordData_len_counts(BeamStatsMonitorProfileRecordData r){ return r->len_counts; }
GPUFUN int64_t BeamStatsMonitorProfileRecordData_len_num_bins(BeamStatsMonitorProfileRecordData r){ return r->n_profiles; }
GPUFUN double* BeamStatsMonitorProfileRecordData_getp1_counts(BeamStatsMonitorProfileRecordData r, int64_t k){ return r->counts + k; }
GPUFUN int64_t* BeamStatsMonitorProfileRecordData_getp1_offsets(BeamStatsMonitorProfileRecordData r, int64_t k){ return r->offsets + k; }
GPUFUN int64_t* BeamStatsMonitorProfileRecordData_getp1_num_bins(BeamStatsMonitorProfileRecordData r, int64_t k){ return r->num_bins + k; }
GPUFUN int64_t* BeamStatsMonitorProfileRecordData_getp1_coord_id(BeamStatsMonitorProfileRecordData r, int64_t k){ return r->coord_id + k; }
GPUFUN double* BeamStatsMonitorProfileRecordData_getp1_min(BeamStatsMonitorProfileRecordData r, int64_t k){ return r->pmin + k; }
GPUFUN double* BeamStatsMonitorProfileRecordData_getp1_bin_width(BeamStatsMonitorProfileRecordData r, int64_t k){ return r->bin_width + k; }
'''


RECORD_STUBS = r'''
/* In-kernel photon logging is outside the contract: the record handle is NULL
   everywhere (synrad_spectrum.h:505 checks it), these only satisfy the compiler. */
typedef struct RecordIndex_s { int64_t dummy; } *RecordIndex;
typedef struct SynchrotronRadiationRecordData_s { int64_t dummy; } *SynchrotronRadiationRecordData;
GPUFUN RecordIndex SynchrotronRadiationRecordData_getp__index(SynchrotronRadiationRecordData r){ (void)r; return NULL; }
GPUFUN int64_t RecordIndex_get_slot(RecordIndex r){ (void)r; return -1; }
#define XTB_REC_SET(f, T) GPUFUN void SynchrotronRadiationRecordData_set_##f(SynchrotronRadiationRecordData r, int64_t i, T v){ (void)r; (void)i; (void)v; }
XTB_REC_SET(photon_energy, double) XTB_REC_SET(at_element, int64_t) XTB_REC_SET(at_turn, int64_t)
XTB_REC_SET(particle_id, int64_t) XTB_REC_SET(particle_delta, double)
'''


def generate():
    out = []
    out.append('/* GENERATED by oracle/gen_shim.py -- ORACLE test infrastructure, do not edit. */')
    out.append('#ifndef XTB_ORACLE_GENERATED_H\n#define XTB_ORACLE_GENERATED_H')
    out.append('#include "xobjects/headers/common.h"')
    out.append('#define XS_FLAG_BACKTRACK (0)\n#define XS_FLAG_KILL_CAVITY_KICK (2)\n'
               '#define XS_FLAG_IGNORE_GLOBAL_APERTURE (3)\n#define XS_FLAG_IGNORE_LOCAL_APERTURE (4)\n'
               '#define XS_FLAG_SR_TAPER (5)\n#define XS_FLAG_SR_KICK_SAME_AS_FIRST (6)')
    out.append('#include "xtrack/particles/rng_src/base_rng.h"')
    out.append(gen_particles_api())
    out.append('#include "xtrack/particles/rng_src/particles_rng.h"')
    out.append('#include "xtrack/particles/local_particle_custom_api.h"')
    out.append('#include "xtrack/headers/constants.h"')
    out.append('#include "xtrack/headers/checks.h"')
    out.append('#include "xtrack/headers/particle_states.h"')
    out.append(RECORD_STUBS)
    out.append('#ifndef XTRACK_MULTIPOLE_NO_SYNRAD')
    # handles of the (field-less) random generator elements, random/random_generators.py
    out.append('typedef void* RandomUniformData;\ntypedef void* RandomUniformAccurateData;\n'
               'typedef void* RandomExponentialData;')
    out.append('#include "xtrack/random/random_src/uniform.h"')
    out.append('#include "xtrack/random/random_src/uniform_accurate.h"')
    out.append('#include "xtrack/random/random_src/exponential.h"')
    out.append('#endif')
    out.append('#include "xtrack/beam_elements/elements_src/track_srotation.h"')
    out.append('#include "xtrack/beam_elements/elements_src/track_drift.h"')
    out.append(MONITOR_API)
    out.append('#include "xtrack/monitors/particles_monitor.h"')
    out.append('#include "xtrack/monitors/last_turns_monitor.h"')
    out.append(BEAM_MONITOR_API)
    out.append('#include "xtrack/monitors/beam_position_monitor.h"')
    out.append('#include "xtrack/monitors/beam_size_monitor.h"')
    out.append('#include "xtrack/monitors/beam_profile_monitor.h"')
    out.append('#include "xtrack/monitors/beam_stats_monitor.h"')
    for name in CLASS_ORDER:
        out.append(gen_element_struct(name))
        out.append(f'#include "xtrack/beam_elements/elements_src/{SPECS[name]["header"]}"')
        out.append(gen_with_transformations(name))
    # dispatch (tracker.py:660-697)
    out.append('/* ---- element dispatch: tracker.py:660-697 ---- */')
    out.append('GPUFUN void xtb_oracle_dispatch(int64_t elem_type, void* el, LocalParticle* lpart){')
    out.append('    switch(elem_type){')
    for name in CLASS_ORDER:
        out.append(f'        case {TYPE_ID[name]}:')
        out.append(f'            {name}_track_local_particle_with_transformations(({name}Data) el, lpart);')
        if SPECS[name].get('isthick') is True:
            out.append('            #ifdef XTRACK_GLOBAL_XY_LIMIT')
            out.append('            global_aperture_check(lpart);')
            out.append('            #endif')
        out.append('            break;')
    out.append('        case 1000: ParticlesMonitor_track_local_particle((ParticlesMonitorData) el, lpart); break;')
    out.append('        case 1001: LastTurnsMonitor_track_local_particle((LastTurnsMonitorData) el, lpart); break;')
    out.append('        case 1002: BeamPositionMonitor_track_local_particle((BeamPositionMonitorData) el, lpart); break;')
    out.append('        case 1003: BeamSizeMonitor_track_local_particle((BeamSizeMonitorData) el, lpart); break;')
    out.append('        case 1004: BeamProfileMonitor_track_local_particle((BeamProfileMonitorData) el, lpart); break;')
    out.append('        case 1005: BeamStatsMonitor_track_local_particle((BeamStatsMonitorData) el, lpart); break;')
    out.append('    }\n}')
    out.append('#endif')
    return '\n'.join(out) + '\n'


if __name__ == '__main__':
    dest = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, '_ref', 'gen')
    os.makedirs(dest, exist_ok=True)
    with open(os.path.join(dest, 'xt_generated.h'), 'w') as fid:
        fid.write(generate())
    print('wrote', os.path.join(dest, 'xt_generated.h'))
