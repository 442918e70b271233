"""2-node / 3-DOF Hermitian (Bernoulli-Euler) beam element of the drop-in API.

Same public surface as the reference's BeamFE2/HermitianBeam_2D.py (constructor :42-67, from_dict
:76-92, l / EA / EI / direction :98-120, Ke / Me :233-301, transfer_matrix / matrix_in_global :168-175,
add_internal_load / reduce_internal_load :122-147, nodal_forces / nodal_reactions(_asvector) :336-358,
internal_action :398-431).  The 6x6 matrices come from the CUDA element kernel (K1, libbfe2); nothing
here assembles or solves on the CPU.
"""
import math

import numpy as np

from . import Loads as BL
# the reference does `from BeamFE2.helpers import *` (HermitianBeam_2D.py:4): user code reaches these through HB.*
from .helpers import compile_global_matrix, np_matrix_tolist, transfer_matrix  # noqa: F401


class HermitianBeam2D(object):
    dof = 3
    dofnames = 'ux', 'uy', 'rotz'
    dynamnames = 'FX', 'FY', 'MZ'
    internal_actions = 'axial', 'shear', 'moment'
    number_internal_points = 10

    def __init__(self, ID=None, i=None, j=None, E=None, crosssection=None, I=None, A=None, rho=None):
        self.ID, self.i, self.j, self.E, self.rho = ID, i, j, E, rho
        if crosssection is not None:
            self.A, self.I = crosssection.A, crosssection.I.x
        else:
            self.A, self.I = A, I
        self.internal_loads = []
        self.mass_matrix = 'consistent'
        self.results = {'linear static': None, 'modal': None, 'buckling': None}

    @classmethod
    def from_dict(cls, adict):
        need = ('ID', 'i', 'j', 'E', 'I', 'A', 'rho')
        for k in need:
            if k not in adict:
                raise Exception('Missing key from dict when creating %s: %r' % (cls.__name__, k))
        assert all(k in need for k in adict)
        return cls(**{k: adict[k] for k in need})

    def __repr__(self):
        return 'HermitianBeam2D(ID=%d, i=%r, j=%r, E=%d, I=%.2f, A=%.2f' % (self.ID, self.i, self.j, self.E, self.I,
                                                                            self.A)

    def __str__(self):
        return 'HB2D_%d' % self.ID

    # ---- scalar geometry / section products
    nodes = property(lambda self: (self.i, self.j))
    l = property(lambda self: math.sqrt((self.i.x - self.j.x) ** 2 + (self.i.y - self.j.y) ** 2))
    EA = property(lambda self: self.E * self.A)
    EI = property(lambda self: self.E * self.I)
    direction = property(lambda self: np.arctan2(self.j.y - self.i.y, self.j.x - self.i.x))
    mass = property(lambda self: self.A / 1e6 * self.rho / 1e3 * self.l / 1e3)

    # ---- matrices: CUDA element kernel (K1)
    def _k1(self, local):
        import torch
        from . import batch
        if local:
            xy = (0.0, 0.0, self.l, 0.0)
        else:
            xy = (self.i.x, self.i.y, self.j.x, self.j.y)

        def t(v):
            return torch.tensor([float(v)], dtype=torch.float64, device='cuda')
        lump = torch.tensor([1 if self.mass_matrix == 'lumped' else 0], dtype=torch.uint8, device='cuda')
        Kg, Mg = batch.element_km(*(t(v) for v in xy), t(self.E), t(self.A), t(self.I),
                                    t(0.0 if self.rho is None else self.rho), lump)   # rho is optional for static-only models
        return Kg[0].cpu().numpy(), Mg[0].cpu().numpy()

    def _Ke_geom(self, N=-1):
        """Geometric stiffness of the element for the axial force N (tension positive), local axes -- the consistent
        matrix of the cubic Hermite shape functions (H.-P. Gavin, CEE 421L, the source HermitianBeam_2D.py:261-282
        cites).  The reference's own transcription has sign slips in four entries ([1,4], [1,5], [5,5] and the missing
        [1,5]) and its buckling solver is unfinished (README.md:27); this is the matrix it names.  The GPU path
        assembles the same numbers in bfe2_assemble_geometric."""
        L = self.l
        a, b, c, d = 6. / 5., L / 10., 2. * L ** 2 / 15., -L ** 2 / 30.
        k = np.zeros((6, 6))
        k[np.ix_([1, 2, 4, 5], [1, 2, 4, 5])] = [[a, b, -a, b], [b, c, -b, d], [-a, -b, a, -b], [b, d, -b, c]]
        return np.multiply(N / L, k)

    @property
    def Ke_geom(self):
        return self._Ke_geom()

    @property
    def Ke(self):
        return self._k1(local=True)[0]

    def _Ke(self):
        return self._k1(local=True)[0]

    def _Me(self):
        return self._k1(local=True)[1]

    def _Me_lumped(self):
        """Lumped mass whatever the beam's mass_matrix flag says (HermitianBeam_2D.py:225-231)."""
        keep, self.mass_matrix = self.mass_matrix, 'lumped'
        try:
            return self._k1(local=True)[1]
        finally:
            self.mass_matrix = keep

    # ---- cubic Hermite shape functions at the normalised position x in [0, 1] (HermitianBeam_2D.py:176-222); host
    #      side, used for deflected shapes -- displacement u(x) = N(x) [ux_i, uy_i, rz_i, ux_j, uy_j, rz_j]
    def N1(self, x, L=1.):
        return 1 - x

    def N2(self, x, L=1.):
        return 1 - x * x * (3 - 2 * x)

    def N3(self, x, L=1.):
        return L * x * (1 - x) ** 2

    def N4(self, x, L=1.):
        return x

    def N5(self, x, L=1.):
        return x * x * (3 - 2 * x)

    def N6(self, x, L=1.):
        return L * x * x * (x - 1)

    def N(self, x, L=1.):
        out = np.zeros((2, 6))
        out[0, 0], out[0, 3] = self.N1(x, L), self.N4(x, L)
        out[1, 1], out[1, 2], out[1, 4], out[1, 5] = self.N2(x, L), self.N3(x, L), self.N5(x, L), self.N6(x, L)
        return out

    def deflected_shape(self, local=True, scale=1., disps=None):
        """Points of the deflected axis from the LOCAL end displacements `disps` (6x1), magnified by `scale`, at
        number_internal_points + 1 stations (HermitianBeam_2D.py:307-334; beam-internal loads not considered).
        local=True: 2x1 matrices in beam axes; local=False: (x, y) tuples in global axes, starting at node i."""
        pts = []
        n = self.number_internal_points
        R = transfer_matrix(alpha=0 if local else self.direction, asdegree=False, blocks=1, blocksize=2)
        for k in range(n + 1):
            v = np.multiply(scale, np.matrix(self.N(x=k / n, L=self.l)) * disps)
            v[0, 0] += (k / n) * self.l
            v = R * v
            if not local:
                v = np_matrix_tolist(v)
                v = (v[0] + self.i.x, v[1] + self.i.y)
            pts.append(v)
        return pts

    @property
    def Me(self):
        assert self.mass_matrix in ['consistent', 'lumped']
        return self._k1(local=True)[1]

    @property
    def transfer_matrix(self):
        c, s = math.cos(self.direction), math.sin(self.direction)
        T = np.zeros((6, 6))
        for o in (0, 3):
            T[o:o + 3, o:o + 3] = [[c, -s, 0], [s, c, 0], [0, 0, 1]]
        return np.matrix(T)

    def matrix_in_global(self, mtrx=None):
        return self.transfer_matrix * np.matrix(mtrx) * self.transfer_matrix.T

    # ---- beam-internal loads
    def add_internal_load(self, loadtype=None, **kwargs):
        assert loadtype in BL.BEAM_LOAD_TYPES
        self.internal_loads.append(BL.LOAD_CLASSES[loadtype](beam=self, **kwargs))

    def reduce_internal_load(self, load):
        r = load.reactions
        return {n: {k: 0 + r[n][k] for k in self.dynamnames} for n in self.nodes}

    @property
    def reduced_internal_loads_asvector(self):
        tot = np.zeros([1, 2 * self.dof])
        for x in self.internal_loads:
            tot += x.reactions_asvector
        return tot.T

    # ---- end forces and internal-action queries (host-side, fed by solver output)
    def nodal_forces(self, disps):
        return np.matrix(self.Ke) * disps

    def nodal_reactions_asvector(self, disps):
        return self.nodal_forces(disps) - self.reduced_internal_loads_asvector

    def nodal_reactions(self, disps):
        v = self.nodal_reactions_asvector(disps)
        return {self.i: {'FX': v[0, 0], 'FY': v[1, 0], 'MZ': v[2, 0]},
                self.j: {'FX': v[3, 0], 'FY': v[4, 0], 'MZ': v[5, 0]}}

    @property
    def internal_action_points(self):
        pts = {0, 1}
        for load in self.internal_loads:
            pts |= set(load.internal_points)
        return sorted(pts)

    def internal_action_at_position(self, action=None, pos=0):
        assert action in self.internal_actions
        return sum(getattr(load, '%s_at_position' % action)(xi=pos) for load in self.internal_loads)

    def internal_action_distribution(self, action=None):
        assert action in self.internal_actions
        return [self.internal_action_at_position(action=action, pos=p) for p in self.internal_action_points]

    def internal_action(self, action=None, disp=None, pos=0, pos_normed=True):
        """End-force baseline (moment: straight line between the end values; axial/shear: -v1) plus the
        add-on curves of the beam's internal loads.  `disp` = local end displacements (6x1)."""
        assert 0 <= pos <= min(1, self.l)
        assert action in self.internal_actions
        if not pos_normed:
            pos /= self.l
        k = self.internal_actions.index(action)
        q = self.nodal_reactions_asvector(disps=disp)
        v1, v2 = q[k, 0], -q[k + self.dof, 0]
        base = v1 + (v2 - v1) * pos if action == 'moment' else -v1
        return self.internal_action_at_position(action=action, pos=pos) + base
