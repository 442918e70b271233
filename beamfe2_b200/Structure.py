"""Structure container of the drop-in API -- same public surface as the reference's
BeamFE2/Structure.py (constructor :23-37, add_nodal_load / add_nodal_mass / add_internal_loads /
clear_loads :53-97, 244-313, nodes / sumdof / positions_to_eliminate / condense :154-191,
K / M / K_with_BC / M_with_masses :193-231, position_in_matrix :379-393, checks :315-354).

Host side = bookkeeping only (load and mass vectors, packing); every matrix and every solve comes
from libbfe2 kernels through the topology Plan.
"""
import copy
import itertools

import numpy as np

from . import Loads as BL
from . import Results, Solver
from . import plan as planmod
# the reference does `from BeamFE2.Loads import *` and `from BeamFE2.helpers import *` (Structure.py:3-8): user code
# reaches these names through the Structure module
from .Loads import (BEAM_LOAD_TYPES, NODAL_LOAD_TYPES, BeamLoad, ConcentratedAxialForce, ConcentratedMoment,  # noqa: F401
                    ConcentratedPerpendicularForce, NodalLoad, NodalMass, UniformAxialForce, UniformPerpendicularForce)
from .helpers import compile_global_matrix, np_matrix_tolist, transfer_matrix  # noqa: F401

VERY_LARGE_NUMBER = 10e40


class Structure(object):
    dof = 3
    loadnames = 'FX', 'FY', 'MZ'
    dofnames = 'ux', 'uy', 'rotz'
    massnames = 'mx', 'my'

    def __init__(self, beams=None, supports=None):
        self.beams = beams
        self.supports = supports
        self.nodal_loads, self.nodal_masses = [], []
        self._load_vector = None
        self._mass_vector = None
        self.results = {'linear static': Results.LinearStaticResult(), 'modal': Results.ModalResult(),
                        'buckling': Results.BucklingResult()}
        self.solver = {'linear static': Solver.LinearStaticSolver(self), 'modal': Solver.ModalSolver(self),
                       'buckling': Solver.BucklingSolver(self)}

    # ---- topology
    @property
    def nodes(self):
        return set(itertools.chain.from_iterable(b.nodes for b in self.beams))

    @property
    def sumdof(self):
        return self.dof * len(self.nodes)

    @property
    def supported_nodes(self):
        return (n for n in self.nodes if n.ID in self.supports.keys())

    def node_by_ID(self, id=None):
        hit = [n for n in self.nodes if n.ID == id]
        assert len(hit) == 1
        return hit[0]

    def position_in_matrix(self, nodeID=None, DOF=None, dynam=None):
        return planmod.position_in_matrix(nodeID=nodeID, DOF=DOF, dynam=dynam)

    def nodenr_dof_from_position(self, position=None):
        node, k = divmod(position, self.dof)
        return node + 1, self.loadnames[k]

    @property
    def positions_to_eliminate(self):
        return planmod.positions_to_eliminate(self.supports)

    def condense(self, mtrx=None):
        keep = [p for p in range(mtrx.shape[0]) if p not in set(self.positions_to_eliminate)]
        return mtrx[np.ix_(keep, keep)]

    def set_mass_matrix_type(self, matrixtype='consistent', beam_IDs='all'):
        if beam_IDs != 'all':
            assert all(x in [b.ID for b in self.beams] for x in beam_IDs)
        for b in self.beams:
            if beam_IDs == 'all' or b.ID in beam_IDs:
                b.mass_matrix = matrixtype

    # ---- loads and masses (host bookkeeping)
    def _node_or_raise(self, nodeID):
        hit = [n for n in self.nodes if n.ID == nodeID]
        if len(hit) != 1:
            raise Exception('There is no node with ID %d' % nodeID)
        return hit[0]

    def add_nodal_mass(self, nodeID=None, mass=None, clear=False):
        node = self._node_or_raise(nodeID)
        for name in self.massnames:
            mass.setdefault(name, 0)
        self.add_mass_to_node(nodeID=nodeID, mass=mass, clear=clear)
        self.nodal_masses.append(BL.NodalMass(node=node, mass=mass))

    def add_nodal_load(self, nodeID=None, dynam=None, clear=False):
        node = self._node_or_raise(nodeID)
        for name in self.loadnames:
            dynam.setdefault(name, 0)
        self.add_single_dynam_to_node(nodeID=nodeID, dynam=dynam, clear=clear)
        self.nodal_loads.append(BL.NodalLoad(node=node, dynam=dynam))

    def add_internal_loads(self, beam=None, loadtype=None, **kwargs):
        beam.add_internal_load(loadtype=loadtype, **kwargs)
        per_node = beam.reduce_internal_load(load=beam.internal_loads[-1])
        for node in beam.nodes:
            self.add_single_dynam_to_node(nodeID=node.ID, dynam=per_node[node])

    def clear_loads(self):
        for b in self.beams:
            b.internal_loads = []
        self.nodal_loads = []
        self._load_vector = None

    def add_single_dynam_to_node(self, nodeID=None, dynam=None, clear=False):
        assert nodeID in [n.ID for n in self.nodes]
        assert all(k in self.loadnames for k in dynam.keys())
        if clear:
            self.clear_loads()
        if self._load_vector is None:
            self._load_vector = np.matrix(np.zeros(self.sumdof))
        for name, value in dynam.items():
            self._load_vector[0, self.position_in_matrix(nodeID=nodeID, dynam=name)] += value

    def add_mass_to_node(self, nodeID=None, mass=None, clear=False):
        assert nodeID in [n.ID for n in self.nodes]
        if clear or self._mass_vector is None:
            self._mass_vector = np.matrix(np.zeros(self.sumdof))
        for name, value in mass.items():
            dofname = self.dofnames[self.massnames.index(name)]
            self._mass_vector[0, self.position_in_matrix(nodeID=nodeID, DOF=dofname)] += value

    @property
    def load_vector(self):
        return self._load_vector.T

    @property
    def mass_vector(self):
        return self._mass_vector

    # ---- packing for the GPU path
    def _sorted_nodes(self):
        nodes = sorted(self.nodes, key=lambda n: n.ID)
        if [n.ID for n in nodes] != list(range(1, len(nodes) + 1)):
            raise Exception('Node numbering is not ok: IDs must be 1..N without gaps')
        return nodes

    def pack(self, fixed=None):
        """(Plan, packed arrays) of the current state of the model: one load case made of the nodal
        loads plus the beams' internal loads, nodal masses summed per node."""
        nodes = self._sorted_nodes()
        nN, nE = len(nodes), len(self.beams)
        conn = np.array([[b.i.ID - 1, b.j.ID - 1] for b in self.beams], dtype=np.int32)
        lumped = np.array([b.mass_matrix == 'lumped' for b in self.beams], dtype=np.uint8)
        fixed = self.positions_to_eliminate if fixed is None else fixed
        p = planmod.Plan(nN, conn, fixed, lumped)
        if p.hbw > 8 and fixed is not None and len(fixed) > 0:
            # the node numbering of the model gives a wide band: let the plan renumber the free DOFs internally
            # (reverse Cuthill-McKee); inputs and outputs stay in the model's numbering
            q = planmod.Plan(nN, conn, fixed, lumped, reorder=True)
            if q.hbw < p.hbw:
                p = q
        coords = np.array([[float(n.x), float(n.y)] for n in nodes])
        props = np.array([[float(b.E), float(b.A), float(b.I), float(b.rho if b.rho is not None else 0.0)] for b in self.beams])
        mass = np.zeros((nN, 2))
        if self._mass_vector is not None:
            mv = np.asarray(self._mass_vector).ravel()
            mass[:, 0], mass[:, 1] = mv[0::3], mv[1::3]
        r_local = np.zeros((1, nE, 6))
        for e, b in enumerate(self.beams):
            for ld in b.internal_loads:
                r_local[0, e] += np.asarray(ld.reactions_asvector).ravel()
        # The reference solves with _load_vector (Solver.py:83), whatever call filled it (add_nodal_load,
        # add_internal_loads or add_single_dynam_to_node directly).  The kernel takes the nodal part and the beam
        # part separately (it needs r_local again for the end forces) and adds T(alpha) r_local itself, so the nodal
        # part is _load_vector minus that: the right-hand side is the reference's in every case.
        f_nodal = np.zeros((1, 3 * nN))
        if self._load_vector is not None:
            f_nodal[0] = np.asarray(self._load_vector, dtype=np.float64).ravel()
            for e, b in enumerate(self.beams):
                if not np.any(r_local[0, e]):
                    continue
                c, s = np.cos(b.direction), np.sin(b.direction)
                R = np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
                f_nodal[0, 3 * conn[e, 0]:3 * conn[e, 0] + 3] -= R @ r_local[0, e, :3]
                f_nodal[0, 3 * conn[e, 1]:3 * conn[e, 1] + 3] -= R @ r_local[0, e, 3:]
        return p, dict(coords=coords, props=props, nodal_mass=mass, f_nodal=f_nodal, r_local=r_local)

    # ---- matrices (CUDA assembly kernel K2 on the uncondensed plan)
    def _assemble_dense(self, with_masses):
        import torch
        from . import batch
        p, pk = self.pack(fixed=[])

        def dev(a):
            return torch.as_tensor(np.ascontiguousarray(a[None]), dtype=torch.float64, device='cuda')
        Kb, Mb = batch.assemble_banded(p, dev(pk['coords']), dev(pk['props']),
                                       dev(pk['nodal_mass']) if with_masses else None)
        torch.cuda.synchronize()
        return np.matrix(p.band_to_dense(Kb[0].cpu().numpy())), np.matrix(p.band_to_dense(Mb[0].cpu().numpy()))

    @property
    def K(self):
        return self._assemble_dense(False)[0]

    @property
    def M(self):
        return self._assemble_dense(False)[1]

    @property
    def K_geom(self):
        """compiled geometric stiffness with the reference's default N = -1 per beam (Structure.py:211-214)"""
        from . import helpers
        return helpers.compile_global_matrix(self.beams, geometrical=True)

    @property
    def M_with_masses(self):
        if self._mass_vector is None:
            self._mass_vector = np.matrix(np.zeros(self.sumdof))
        return self._assemble_dense(True)[1]

    @property
    def K_with_BC(self):
        K = copy.deepcopy(self.K)
        for pos in self.positions_to_eliminate:
            K[pos, pos] += VERY_LARGE_NUMBER
        return K

    # ---- sanity checks
    @property
    def stiffness_matrix_is_symmetric(self):
        K = self.K
        ok = np.allclose(K.T - K, 0)
        if not ok:
            print('The stiffness matrix is not symmetric')
        return ok

    @property
    def stiffness_matrix_is_nonsingular(self):
        """Positive definiteness of the supported stiffness matrix = the GPU band Cholesky succeeds."""
        import torch
        from . import batch
        p, pk = self.pack()
        if p.n_free == 0:
            return True

        def dev(a):
            return torch.as_tensor(np.ascontiguousarray(a[None]), dtype=torch.float64, device='cuda')
        f = np.zeros((1, 1, self.sumdof))
        out = batch.solve_fused(p, dev(pk['coords']), dev(pk['props']), None, dev(f[0]), None, modal=False,
                                want_endforces=False, want_reactions=False)
        return int(out['info'][0].item()) == 0

    @property
    def node_numbering_is_ok(self):
        ok = set(n.ID for n in self.nodes) == set(range(1, len(self.nodes) + 1))
        if not ok:
            print('Node numbering is not ok')
        return ok

    @property
    def mass_matrix_is_ok(self):
        return True

    @property
    def stiffness_matrix_is_ok(self):
        ok = self.stiffness_matrix_is_nonsingular
        if not ok:
            print('The stiffness matrix is singular')
        return ok

    def compile_global_stiffness_matrix(self):
        return self.K

    def compile_global_geometrical_stiffness_matrix(self):
        return self.K_geom

    def draw(self, show=True, analysistype=None, mode=0, internal_action=None):
        if self.results[analysistype].solved:                          # Structure.py:110-114
            print('drawing is outside the accelerated path')
        else:
            print('no results available, no printing')
