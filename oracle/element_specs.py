"""ORACLE (test infrastructure) -- element struct specifications.

One entry per reference element class that the hot path covers.  Each spec
restates the reference's `_xofields` table (file cited per class) so that
`gen_shim.py` can synthesise what xobjects would have generated: a C struct
`<Class>Data` plus the `<Class>Data_get_*/getp1_*/len_*` accessors that the
reference's per-element headers call.

kind: 'f64' | 'i64' | 'arr' (double[:]).
`attr` maps a struct field to the attribute of the host element object
(xtrack_b200.elements) when the names differ.
"""

MISALIGN = [('shift_x', 'f64'), ('shift_y', 'f64'), ('shift_s', 'f64'),
            ('rot_s_rad', 'f64'), ('rot_x_rad', 'f64'), ('rot_y_rad', 'f64'),
            ('rot_s_rad_no_frame', 'f64'), ('rot_shift_anchor', 'f64')]

_KNL = [('order', 'i64'), ('inv_factorial_order', 'f64'), ('knl', 'arr'),
        ('ksl', 'arr'), ('knl_rel', 'arr'), ('ksl_rel', 'arr')]

_STRAIGHT_COMMON = [('length', 'f64'), ('num_multipole_kicks', 'i64'), *_KNL,
                    ('main_is_skew', 'i64'), ('edge_entry_active', 'i64'),
                    ('edge_exit_active', 'i64'), ('model', 'i64'),
                    ('integrator', 'i64'), ('radiation_flag', 'i64'),
                    ('delta_taper', 'f64')]

# _common.py:566-597
_BEND_COMMON = [('k0', 'f64'), ('k1', 'f64'), ('k2', 'f64'), ('h', 'f64'),
                ('angle', 'f64'), ('length', 'f64'), ('model', 'i64'),
                ('integrator', 'i64'), ('radiation_flag', 'i64'),
                ('delta_taper', 'f64'), ('edge_entry_active', 'i64'),
                ('edge_exit_active', 'i64'), ('edge_entry_model', 'i64'),
                ('edge_exit_model', 'i64'), ('edge_entry_angle', 'f64'),
                ('edge_exit_angle', 'f64'), ('edge_entry_angle_fdown', 'f64'),
                ('edge_exit_angle_fdown', 'f64'), ('edge_entry_fint', 'f64'),
                ('edge_exit_fint', 'f64'), ('edge_entry_hgap', 'f64'),
                ('edge_exit_hgap', 'f64'), ('num_multipole_kicks', 'i64'),
                *_KNL, ('k0_from_h', 'i64')]

# name -> spec.  `isthick`: literal class-level thickness (drives the
# global-aperture check, tracker.py:681-689); `dyn_thick`: has an `isthick`
# *field* (Multipole); `rot_shift`: allow_rot_and_shift.
SPECS = {
    'Marker': dict(header='marker.h', fields=[('_dummy', 'i64')],
                   isthick=False, rot_shift=False),
    # drift.py
    'Drift': dict(header='drift.h', fields=[('length', 'f64'), ('model', 'i64')],
                  isthick=True, rot_shift=False),
    'DriftExact': dict(header='drift_exact.h', fields=[('length', 'f64')],
                       isthick=True, rot_shift=False),
    # multipole.py:80-97
    'Multipole': dict(header='multipole.h',
                      fields=[('order', 'i64'), ('inv_factorial_order', 'f64'),
                              ('length', 'f64'), ('hxl', 'f64'),
                              ('radiation_flag', 'i64'), ('delta_taper', 'f64'),
                              ('knl', 'arr'), ('ksl', 'arr'), ('knl_rel', 'arr'),
                              ('ksl_rel', 'arr'), ('main_order', 'i64'),
                              ('main_is_skew', 'i64'), ('isthick', 'i64'),
                              ('num_multipole_kicks', 'i64'), ('model', 'i64'),
                              ('integrator', 'i64')],
                      attr={'isthick': '_isthick_field'},
                      isthick=False, dyn_thick=True, rot_shift=True,
                      internal_record=True),
    # quadrupole.py:65-83, sextupole.py, octupole.py
    'Quadrupole': dict(header='quadrupole.h',
                       fields=[('k1', 'f64'), ('k1s', 'f64'), *_STRAIGHT_COMMON],
                       isthick=True, rot_shift=True),
    'Sextupole': dict(header='sextupole.h',
                      fields=[('k2', 'f64'), ('k2s', 'f64'), *_STRAIGHT_COMMON],
                      isthick=True, rot_shift=True),
    'Octupole': dict(header='octupole.h',
                     fields=[('k3', 'f64'), ('k3s', 'f64'), *_STRAIGHT_COMMON],
                     isthick=True, rot_shift=True),
    'Bend': dict(header='bend.h', fields=list(_BEND_COMMON),
                 attr={'k0': '_k0', 'k0_from_h': '_k0_from_h'},
                 isthick=True, rot_shift=True, curved=True),
    # rbend.py:91-98
    'RBend': dict(header='rbend.h',
                  fields=[*_BEND_COMMON, ('length_straight', 'f64'),
                          ('rbend_model', 'i64'), ('rbend_compensate_sagitta', 'i64'),
                          ('rbend_shift', 'f64'), ('rbend_angle_diff', 'f64')],
                  attr={'k0': '_k0', 'k0_from_h': '_k0_from_h'},
                  isthick=True, rot_shift=True, curved=True),
    # cavity.py:63-76
    'Cavity': dict(header='cavity.h',
                   fields=[('length', 'f64'), ('voltage', 'f64'), ('frequency', 'f64'),
                           ('lag', 'f64'), ('phase', 'f64'), ('harmonic', 'f64'),
                           ('lag_taper', 'f64'), ('phase_taper', 'f64'),
                           ('absolute_time', 'i64'), ('num_kicks', 'i64'),
                           ('model', 'i64'), ('integrator', 'i64')],
                   isthick=True, rot_shift=True),
    # crab_cavity.py:51-63
    'CrabCavity': dict(header='crab_cavity.h',
                       fields=[('length', 'f64'), ('crab_voltage', 'f64'), ('frequency', 'f64'),
                               ('lag', 'f64'), ('phase', 'f64'), ('lag_taper', 'f64'),
                               ('phase_taper', 'f64'), ('absolute_time', 'i64'),
                               ('num_kicks', 'i64'), ('model', 'i64'), ('integrator', 'i64')],
                       isthick=True, rot_shift=True),
    # rf_multipole.py:51-65
    'RFMultipole': dict(header='rfmultipole.h',
                        fields=[('voltage', 'f64'), ('frequency', 'f64'), ('lag', 'f64'),
                                ('phase', 'f64'), ('order', 'i64'),
                                ('inv_factorial_order', 'f64'), ('knl', 'arr'),
                                ('ksl', 'arr'), ('pn', 'arr'), ('ps', 'arr'),
                                ('phase_n', 'arr'), ('phase_s', 'arr'),
                                ('absolute_time', 'i64')],
                        isthick=False, rot_shift=True),
    # dipole_edge.py:40-51
    'DipoleEdge': dict(header='dipoleedge.h',
                       fields=[('r21', 'f64'), ('r43', 'f64'), ('hgap', 'f64'),
                               ('k', 'f64'), ('e1', 'f64'), ('e1_fd', 'f64'),
                               ('fint', 'f64'), ('model', 'i64'), ('side', 'i64'),
                               ('delta_taper', 'f64')],
                       isthick=False, rot_shift=True),
    'SRotation': dict(header='srotation.h',
                      fields=[('cos_z', 'f64'), ('sin_z', 'f64')],
                      isthick=False, rot_shift=False),
    'XYShift': dict(header='xyshift.h', fields=[('dx', 'f64'), ('dy', 'f64')],
                    isthick=False, rot_shift=False),
    # rotation.py:29-36, translation.py:27-30
    'Rotation': dict(header='rotation.h',
                     fields=[('rot_s_rad', 'f64'), ('rot_x_rad', 'f64'), ('rot_y_rad', 'f64'),
                             ('_first_rot', 'i64'), ('_second_rot', 'i64'), ('_third_rot', 'i64')],
                     isthick=False, rot_shift=False),
    'Translation': dict(header='translation.h',
                        fields=[('shift_x', 'f64'), ('shift_y', 'f64')],
                        isthick=False, rot_shift=False),
    'LimitRect': dict(header='limitrect.h',
                      fields=[('min_x', 'f64'), ('max_x', 'f64'), ('min_y', 'f64'),
                              ('max_y', 'f64')],
                      isthick=False, rot_shift=True),
    'LimitEllipse': dict(header='limitellipse.h',
                         fields=[('a_squ', 'f64'), ('b_squ', 'f64'), ('a_b_squ', 'f64')],
                         isthick=False, rot_shift=True),
    'LimitPolygon': dict(header='limitpolygon.h',
                         fields=[('x_vertices', 'arr'), ('y_vertices', 'arr'),
                                 ('x_normal', 'arr'), ('y_normal', 'arr'),
                                 ('resc_fac', 'f64')],
                         defaults={'x_normal': [], 'y_normal': [], 'resc_fac': 1.0},
                         isthick=False, rot_shift=True),
}

# ---- slices (slice_base.py:9-14; slice_elements_{thin,thick,drift,edge}.py): a reference to
# the parent element + four fields of their own; the generated wrappers read everything else
# through `XData_get__parent_<field>`
SLICE_FIELDS = [('radiation_flag', 'i64'), ('delta_taper', 'f64'), ('weight', 'f64'),
                ('slice_offset', 'f64')]


def _snake(parent):
    return {'RBend': 'rbend', 'CrabCavity': 'crab_cavity'}.get(parent, parent.lower())


def _add_slices():
    for parent in ('Bend', 'RBend', 'Quadrupole', 'Sextupole', 'Octupole', 'Multipole', 'Cavity',
                   'CrabCavity'):
        kinds = [('ThinSlice' + parent, f'thin_slice_{_snake(parent)}.h', False, True, True),
                 ('ThickSlice' + parent, f'thick_slice_{_snake(parent)}.h', True, True, False),
                 ('DriftSlice' + parent, f'drift_slice_{_snake(parent)}.h', True, False, False)]
        if parent not in ('Multipole', 'Cavity', 'CrabCavity'):
            kinds += [('ThinSlice' + parent + 'Entry', f'thin_slice_{_snake(parent)}_entry.h',
                       False, True, False),
                      ('ThinSlice' + parent + 'Exit', f'thin_slice_{_snake(parent)}_exit.h',
                       False, True, False)]
        for cname, header, thick, from_parent, is_thin in kinds:
            SPECS[cname] = dict(header=header, fields=list(SLICE_FIELDS), parent=parent,
                                isthick=thick, rot_shift=False, rot_shift_from_parent=from_parent,
                                thin_slice=is_thin, curved=SPECS[parent].get('curved', False))
    SPECS['DriftSlice'] = dict(header='drift_slice.h', fields=list(SLICE_FIELDS), parent='Drift',
                               isthick=True, rot_shift=False, rot_shift_from_parent=False)
    SPECS['DriftExactSlice'] = dict(header='drift_exact_slice.h', fields=list(SLICE_FIELDS),
                                    parent='DriftExact', isthick=True, rot_shift=False,
                                    rot_shift_from_parent=False)


_add_slices()

# parents before their slices in the generated header; ids as tracker.py:517-519 (sorted names)
CLASS_ORDER = sorted(SPECS, key=lambda nn: ('parent' in SPECS[nn], nn))
TYPE_ID = {name: ii for ii, name in enumerate(CLASS_ORDER)}


def all_fields(name):
    spec = SPECS[name]
    ff = list(spec['fields'])
    if spec.get('rot_shift'):
        ff += MISALIGN
    return ff
