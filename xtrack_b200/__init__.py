"""xtrack_b200 -- B200-native multi-turn single-particle tracking behind the
xtrack `Line.track(particles, num_turns, turn_by_turn_monitor=...)` API.

Only the tracking hot path of xsuite/xtrack is provided; the
compute path is hand-written sm_100a CUDA behind a C-ABI (`include/xtb200.h`,
`xtrack_b200/csrc/`).  There is no CPU fallback: tracking raises if the CUDA
library or a GPU is missing.
"""
from .particles import Particles, LAST_INVALID_STATE, PROTON_MASS_EV, ELECTRON_MASS_EV
from .elements import (Marker, Drift, DriftExact, Multipole, Quadrupole, Sextupole,
                       Octupole, Bend, RBend, Cavity, CrabCavity, RFMultipole, DipoleEdge, SRotation, XYShift, Rotation, Translation,
                       LimitRect, LimitEllipse, LimitPolygon)
from .monitors import (ParticlesMonitor, LastTurnsMonitor, BeamPositionMonitor,
                       BeamSizeMonitor, BeamProfileMonitor, BeamStatsMonitor)
from .line import Line
from .loss_location_refinement import LossLocationRefinement

__version__ = '0.1.0'
