"""Loss-location refinement, the reference's own test restated:
tests/test_collimation.py:105-223 `test_aperture_refinement` -- same line, same beam, same
parameters, same assertions and tolerance -- on the host build of the device code and, marked
`gpu`, on the B200.  (The reference runs this on its CPU context only.)
"""
import numpy as np
import pytest

import xtrack_b200 as xb
from xtrack_b200 import loss_location_refinement as llr
import common
from test_rows_both_tiers import BACKENDS, _build


def _example_line(sandwitch_aper, shift_x, shift_y):
    aper_0 = xb.LimitEllipse(a=2e-2, b=2e-2)
    aper_1 = xb.LimitEllipse(a=1e-2, b=1e-2)
    rot_deg = 10.
    if sandwitch_aper:
        def sandwich(aper):
            return [xb.Translation(shift_x=shift_x, shift_y=shift_y),
                    xb.Rotation(rot_s_rad=np.deg2rad(rot_deg)), aper, xb.Multipole(knl=[0.00]),
                    xb.Rotation(rot_s_rad=np.deg2rad(-rot_deg)),
                    xb.Translation(shift_x=-shift_x, shift_y=-shift_y)]
        els_0, els_1 = sandwich(aper_0), sandwich(aper_1)
    else:
        for aper in (aper_0, aper_1):
            aper.shift_x = shift_x
            aper.shift_y = shift_y
            aper.rot_s_rad = np.deg2rad(rot_deg)
        els_0 = [aper_0, xb.Multipole(knl=[0.0])]
        els_1 = [aper_1, xb.Multipole(knl=[0.00])]
    els = ([xb.Drift(length=0.5)] + els_0
           + [xb.Drift(length=1), xb.Multipole(knl=[0.]), xb.Quadrupole(length=1),
              xb.Cavity(voltage=3e6, frequency=400e6),
              xb.ParticlesMonitor(start_at_turn=0, stop_at_turn=10, num_particles=3),
              xb.Drift(length=1.), xb.Marker()]
           + els_1)
    return xb.Line(elements=els), aper_0, aper_1


@pytest.mark.parametrize('on_gpu', BACKENDS)
@pytest.mark.parametrize('sandwitch_aper', [True, False], ids=['sandwich_aper', 'no_sandwich_aper'])
def test_aperture_refinement(sandwitch_aper, on_gpu):
    n_part = 10000
    shift_x, shift_y = 0.3e-2, 0.5e-2
    line, aper_0, aper_1 = _example_line(sandwitch_aper, shift_x, shift_y)
    dev = _build(line, on_gpu)

    r = np.linspace(0, 0.018, n_part)
    theta = np.linspace(0, 8 * np.pi, n_part)
    particles = xb.Particles(p0c=6500e9, x=r * np.cos(theta) + shift_x,
                             y=r * np.sin(theta) + shift_y, _device=dev)
    line.track(particles)

    refinement = xb.LossLocationRefinement(
        line, n_theta=360, r_max=0.5, dr=50e-6, ds=0.1, save_refine_lines=True,
        allowed_backtrack_types=[xb.Multipole, xb.Cavity])
    refinement.refine_loss_location(particles)

    state = particles.get('state')
    mask_lost = state == 0
    x, y, s = particles.get('x'), particles.get('y'), particles.get('s')
    r_calc = np.sqrt((x - shift_x) ** 2 + (y - shift_y) ** 2)
    assert np.all(r_calc[~mask_lost] < 1e-2)
    assert np.all(r_calc[mask_lost] > 1e-2)
    i_aper_1 = line.elements.index(aper_1)
    at_element = particles.get('at_element')
    assert np.all(at_element[mask_lost] == i_aper_1)
    assert np.all(at_element[~mask_lost] == 0)
    s_el = llr._element_s_locations(line)
    s0, s1 = s_el[line.elements.index(aper_0)], s_el[i_aper_1]
    r0, r1 = np.sqrt(aper_0.a_squ), np.sqrt(aper_1.a_squ)
    s_expected = s0 + (r_calc - r0) / (r1 - r0) * (s1 - s0)
    np.testing.assert_allclose(s[mask_lost], s_expected[mask_lost], atol=0.11)   # (the reference's bar)
    # the refined stretch: the apertures every ds, some of them inside the quadrupole
    interp = refinement.refine_lines[i_aper_1]
    classes = [type(ee).__name__ for ee in interp.elements]
    assert classes.count('LimitPolygon') == 31
    assert classes.count('ThickSliceQuadrupole') == 10
    assert abs(interp.get_length() - 3.0) < 1e-12


def test_replicate_mode_and_refusals():
    """Identical apertures without transformations in between are copied, not characterised
    (:161-172); an element that cannot be backtracked stops the refinement (:334-347); the
    first aperture of a line cannot be refined (:136-140)."""
    import hostsim
    aper = dict(min_x=-1e-2, max_x=1e-2, min_y=-5e-3, max_y=5e-3)
    els = [xb.LimitRect(**aper), xb.Drift(length=2.0), xb.Sextupole(length=0.5, k2=0.1),
           xb.Drift(length=1.5), xb.LimitRect(**aper)]
    line = xb.Line(elements=els)
    line.build_tracker(_device='cpu', _tracker_class=hostsim.HostSimTracker)
    n = 2000
    rng = np.random.default_rng(3)
    particles = xb.Particles(p0c=6500e9, x=rng.uniform(-8e-3, 8e-3, n), px=rng.uniform(-3e-3, 3e-3, n),
                             y=rng.uniform(-4e-3, 4e-3, n), py=rng.uniform(-2e-3, 2e-3, n))
    start = {nn: particles.get(nn).copy() for nn in ('x', 'px', 'y', 'py')}
    line.track(particles)
    lost = particles.get('state') == 0
    assert 100 < lost.sum() < n - 100 and np.all(particles.get('at_element')[lost] == 4)
    assert np.allclose(particles.get('s')[lost], 4.0)
    ref = xb.LossLocationRefinement(line, ds=0.05, save_refine_lines=True)
    ref.refine_loss_location(particles)
    interp = ref.refine_lines[4]
    assert all(type(ee).__name__ != 'LimitPolygon' for ee in interp.elements)
    s = particles.get('s')[lost]
    assert np.all(s > 0.0) and np.all(s <= 4.0 + 1e-9) and s.std() > 0.5
    # the sextupole is weak: straight lines from the start cross the chamber wall where
    # the refined s says, to within one aperture spacing
    t_x = np.where(start['px'] > 0, (1e-2 - start['x']) / start['px'], (-1e-2 - start['x']) / start['px'])
    t_y = np.where(start['py'] > 0, (5e-3 - start['y']) / start['py'], (-5e-3 - start['y']) / start['py'])
    s_wall = np.minimum(t_x, t_y)[lost]
    assert np.all(s >= s_wall - 5e-3) and np.all(s <= s_wall + 0.05 + 5e-3)

    els = [xb.LimitRect(**aper), xb.Drift(length=2.0), xb.Multipole(knl=[0, 0.1]),
           xb.Drift(length=1.5), xb.LimitRect(**aper)]
    line = xb.Line(elements=els)
    line.build_tracker(_device='cpu', _tracker_class=hostsim.HostSimTracker)
    particles = xb.Particles(p0c=6500e9, x=[0., 5e-3], px=[0., 3e-3])
    line.track(particles)
    with pytest.raises(TypeError, match='Cannot backtrack through element'):
        xb.LossLocationRefinement(line, ds=0.05).refine_loss_location(particles)
    xb.LossLocationRefinement(line, ds=0.05, allowed_backtrack_types=[xb.Multipole]
                              ).refine_loss_location(particles)
    assert 0.0 < particles.get('s')[1] < 3.5


def test_polygon_impact_from_origin():
    xv, yv = llr.polygon_impact_from_origin([1, -1, -1, 1], [1, 1, -1, -1],
                                            np.deg2rad([0., 45., 90., 200.]))
    assert np.allclose(xv, [1, 1, 0, -1]) and np.allclose(yv, [0, 1, 1, -np.tan(np.deg2rad(20.))])
