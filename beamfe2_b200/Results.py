"""Result containers of the drop-in API -- same attributes and methods as the reference's
BeamFE2/Results.py (AnalysisResult :10-84, LinearStaticResult :87-143, ModalResult :146-173).
The numbers inside come from the GPU solvers (beamfe2_b200/Solver.py)."""
import math

import numpy as np

REACTION_SOLUTION_ROUNDOFF = 5


class AnalysisResult(object):
    def __init__(self, structure=None):
        self.structure = structure
        self.displacement_results = []
        self.solved = False

    def global_displacement_vector(self, mode=None):
        return self.displacement_results[mode]

    def global_displacements(self, mode=0, asvector=False):
        d = self.global_displacement_vector(mode)
        if asvector:
            return d
        return {name: d[k::self.structure.dof] for k, name in enumerate(self.structure.dofnames)}

    def element_displacements(self, local=True, mode=0, beam=None, asvector=False):
        """End displacements of one beam, local (rotated by -direction) or global; 6x1 matrix or a dict
        partitioned by DOF name."""
        assert beam in self.structure.beams
        d = self.global_displacement_vector(mode)
        a = -beam.direction if local else 0.0
        c, s = math.cos(a), math.sin(a)
        R = np.matrix([[c, -s, 0], [s, c, 0], [0, 0, 1]])
        pi = self.structure.position_in_matrix(nodeID=beam.i.ID, DOF='ux')
        pj = self.structure.position_in_matrix(nodeID=beam.j.ID, DOF='ux')
        out = np.concatenate([R * np.matrix(d[pi:pi + 3]), R * np.matrix(d[pj:pj + 3])], axis=0)
        if asvector:
            return out
        return {name: out[k::beam.dof] for k, name in enumerate(beam.dofnames)}


class LinearStaticResult(AnalysisResult):
    def __init__(self, structure=None, displacements=None, endforces=None, reactions=None, f_nodal=None):
        super(LinearStaticResult, self).__init__(structure=structure)
        self.displacement_results = displacements
        self.element_end_forces = endforces        # [nE, 6] local end forces from the GPU (K3)
        self._reactions = reactions                # [sumdof] from the GPU (K3): rotated end forces minus f_nodal
        self._f_nodal = f_nodal                    # [sumdof] the nodal right-hand side the kernel subtracted

    @property
    def reaction_forces_asvector(self):
        """Global reaction vector (sum of rotated beam end forces minus the applied nodal loads).

        The reference subtracts the entries of structure.nodal_loads (Results.py:119-123), i.e. what add_nodal_load
        recorded -- a load put straight into the load vector with add_single_dynam_to_node is solved for but NOT
        subtracted.  The kernel subtracts the whole nodal right-hand side; the difference between the two (zero unless
        add_single_dynam_to_node was called directly) is put back here, at access time like the reference."""
        if not self.solved:
            print('no results available, solve the model first')
            return None
        out = np.array(self._reactions, dtype=np.float64).reshape(-1, 1)
        if self._f_nodal is not None:
            recorded = np.zeros_like(out)
            for x in self.structure.nodal_loads:
                pos = self.structure.position_in_matrix(nodeID=x.node.ID, DOF='ux')
                recorded[pos:pos + 3] += np.asarray(x.asvector, dtype=np.float64).reshape(3, 1)
            extra = np.asarray(self._f_nodal, dtype=np.float64).reshape(-1, 1) - recorded
            if np.any(extra):
                out = out + extra
        return out

    @property
    def reaction_forces(self):
        vec = self.reaction_forces_asvector
        out = {}
        st = self.structure
        for node in st.supported_nodes:
            out[node.ID] = {}
            for dofname in st.supports[node.ID]:
                pos = st.position_in_matrix(nodeID=node.ID, DOF=dofname)
                out[node.ID][st.loadnames[st.dofnames.index(dofname)]] = round(vec[pos, 0], REACTION_SOLUTION_ROUNDOFF)
        return out


class ModalResult(AnalysisResult):
    def __init__(self, structure=None, circular_freq=None, modalshapes=None):
        super(ModalResult, self).__init__(structure=structure)
        self.circular_frequencies = circular_freq
        self.displacement_results = modalshapes

    def _need(self):
        if not self.solved:
            print('no results available, solve the model first')
        return self.solved

    @property
    def frequencies(self):
        return [w / (2. * math.pi) for w in self.circular_frequencies] if self._need() else None

    @property
    def periods(self):
        return [1. / f for f in self.frequencies] if self._need() else None

    def modeshape(self, mode=None):
        return self.displacement_results[mode] if self._need() else None


class BucklingResult(AnalysisResult):
    def __init__(self, structure=None, criticals=None, bucklingshapes=None):
        super(BucklingResult, self).__init__(structure=structure)
        self.criticals = criticals
        self.displacement_results = bucklingshapes
