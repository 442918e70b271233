"""Load records of the drop-in API.

Mirrors the reference's BeamFE2/Loads.py: NodalLoad / NodalMass (:17-63) and the five beam-internal
load types with their LOCAL fixed-end vectors (`reactions_asvector`, :202-208, 264-270, 333-342,
409-416, 474-484), their rotation to the global system (`reactions`, :100-112) and the
simply-supported add-on curves used by internal-action queries (:214-225, 277-289, 347-365, 421-432,
489-502).  These are scalar host-side formulas that build the right-hand side; the solve itself
runs on the GPU.  One table-driven class replaces the reference's five subclasses.
"""
import math

import numpy as np

BEAM_LOAD_TYPES = ('concentrated perpendicular force', 'concentrated moment', 'concentrated axial force',
                   'uniform axial force', 'uniform perpendicular force')
NODAL_LOAD_TYPES = ['force', 'moment']


class NodalMass(object):
    def __init__(self, node=None, mass=None, *args, **kwargs):
        self.node, self.mass = node, mass

    @property
    def asvector(self):
        return np.matrix([self.mass['mx'], self.mass['my']])


class NodalLoad(object):
    def __init__(self, node=None, dynam=None, *args, **kwargs):
        self.node, self.dynam = node, dynam

    @property
    def asvector(self):
        return np.matrix([self.dynam['FX'], self.dynam['FY'], self.dynam['MZ']])


def fixed_end_vector(loadtype, value, L, position=0.0):
    """Local 6-vector [FX_i, FY_i, MZ_i, FX_j, FY_j, MZ_j] equivalent to the load (a = position ratio)."""
    a, b = position, 1 - position
    if loadtype == 'uniform axial force':
        return [value * L / 2, 0, 0, value * L / 2, 0, 0]
    if loadtype == 'uniform perpendicular force':
        return [0, value * L / 2, value * (L ** 2) / 12, 0, value * L / 2, -1 * value * (L ** 2) / 12.]
    if loadtype == 'concentrated perpendicular force':
        return [0, (3 - 2 * b) * (b ** 2) * value, a * (b ** 2) * value * L,
                0, (3 - 2 * a) * (a ** 2) * value, -1 * b * (a ** 2) * value * L]
    if loadtype == 'concentrated axial force':
        return [value / 2., 0, 0, value / 2., 0, 0]
    if loadtype == 'concentrated moment':
        MR = 6 * a * b * value / L
        return [0, -MR, 1 * (3 * b - 2) * b * value, 0, MR, 1 * (3 * a - 2) * a * value]
    raise NotImplementedError(loadtype)


class BeamLoad(object):
    """A load acting inside a beam, defined in the beam's local system (value, optional position)."""

    def __init__(self, loadtype=None, beam=None, value=1, position=0, *args, **kwargs):
        assert loadtype in BEAM_LOAD_TYPES
        self.loadtype, self.beam, self.value, self.position = loadtype, beam, value, position
        self.nr_points = beam.number_internal_points
        self.EPS = 1e-15

    alpha = property(lambda self: self.position)
    beta = property(lambda self: 1 - self.position)

    @property
    def reactions_asvector(self):
        return np.matrix([fixed_end_vector(self.loadtype, self.value, self.beam.l, self.position)])

    @property
    def reactions(self):
        """Equivalent nodal loads in the GLOBAL system, {node: {FX, FY, MZ}} for both beam nodes."""
        r = fixed_end_vector(self.loadtype, self.value, self.beam.l, self.position)
        c, s = math.cos(self.beam.direction), math.sin(self.beam.direction)
        g = [c * r[0] - s * r[1], s * r[0] + c * r[1], r[2], c * r[3] - s * r[4], s * r[3] + c * r[4], r[5]]
        return {self.beam.i: {'FX': g[0], 'FY': g[1], 'MZ': g[2]},
                self.beam.j: {'FX': g[3], 'FY': g[4], 'MZ': g[5]}}

    @property
    def internal_points(self):
        if self.loadtype == 'uniform axial force':
            return []
        if self.loadtype == 'uniform perpendicular force':
            return [x / self.nr_points for x in range(self.nr_points + 1)]
        return [0, self.position - self.beam.l * self.EPS, self.position + self.beam.l * self.EPS, 1]

    # ---- simply-supported add-on curves, xi in [0, 1] from node i
    def axial_at_position(self, xi):
        if self.loadtype == 'uniform axial force':
            return -xi * self.value * self.beam.l
        if self.loadtype == 'concentrated axial force':
            return 0 if xi < self.position else -self.value
        return 0

    def shear_at_position(self, xi):
        if self.loadtype == 'uniform perpendicular force':
            return xi * self.beam.l * -self.value
        if self.loadtype == 'concentrated perpendicular force':
            return 0 if xi < self.position else -self.value
        return 0

    def moment_at_position(self, xi):
        L = self.beam.l
        if self.loadtype == 'uniform perpendicular force':
            return ((xi * (1 - xi)) * self.value * L ** 2) / 2.
        if self.loadtype == 'concentrated perpendicular force':
            m = self.beta * self.value * xi
            if xi > self.position:
                m -= abs(xi - self.position) * self.value
            return m * L
        if self.loadtype == 'concentrated moment':
            m = -(self.value / L) * xi * L
            if xi > self.position:
                m += self.value
            return m
        return 0

    def deflection_at_position(self, xi):
        if self.loadtype == 'uniform perpendicular force':
            return ((1 + xi * (1 - xi)) / (24 * self.beam.EI)) * (xi * (1 - xi)) * self.value * self.beam.l
        return 0

    moments = property(lambda self: (self.moment_at_position(x) for x in self.internal_points))
    shears = property(lambda self: (self.shear_at_position(x) for x in self.internal_points))
    deflections = property(lambda self: (self.deflection_at_position(x) for x in self.internal_points))


def _typed(name, loadtype):
    """Subclass with the reference's class name and constructor order (Loads.py:172, 236, 300, 376, 443) whose load
    type is fixed; all behaviour lives in BeamLoad."""
    def __init__(self, loadtype=loadtype, value=None, position=0, beam=None, **kw):
        BeamLoad.__init__(self, loadtype=loadtype, beam=beam, value=value, position=position)
    return type(name, (BeamLoad,), {'__init__': __init__, '__doc__': '%s acting inside a beam (local system).' % loadtype})


UniformAxialForce = _typed('UniformAxialForce', 'uniform axial force')
UniformPerpendicularForce = _typed('UniformPerpendicularForce', 'uniform perpendicular force')
ConcentratedPerpendicularForce = _typed('ConcentratedPerpendicularForce', 'concentrated perpendicular force')
ConcentratedAxialForce = _typed('ConcentratedAxialForce', 'concentrated axial force')
ConcentratedMoment = _typed('ConcentratedMoment', 'concentrated moment')
LOAD_CLASSES = {'uniform axial force': UniformAxialForce, 'uniform perpendicular force': UniformPerpendicularForce,
                'concentrated perpendicular force': ConcentratedPerpendicularForce,
                'concentrated axial force': ConcentratedAxialForce, 'concentrated moment': ConcentratedMoment}
