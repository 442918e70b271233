"""The two solve() bodies a maintainer would put into BeamFE2/Solver.py in place of :29-69 and :76-87
(option B), written as a patch that can be applied to the imported, otherwise UNMODIFIED
reference package:

    import BeamFE2.Solver, BeamFE2.Results
    from integration import solver_patch
    solver_patch.apply(BeamFE2.Solver, BeamFE2.Results)

tests/test_reference_suite.py applies it to baseline/_ref and runs the reference's own Tests/ on top."""
import math

import numpy as np

from . import _bfe2


def apply(SolverModule, ResultsModule):
    Results = ResultsModule

    def static_solve(self):
        structure = self.structure
        if np.count_nonzero(structure._load_vector) == 0:          # Solver.py:79-80
            return False
        out = _bfe2.solve_structure(structure, static=True)
        if out['info'] > 0:
            raise np.linalg.LinAlgError('Singular matrix')         # what scipy.linalg.inv raises at Solver.py:83
        disps = np.matrix(out['u'][0]).T                           # was: sp.inv(K_with_BC) * load_vector
        structure.results['linear static'] = Results.LinearStaticResult(structure=structure, displacements=[disps])
        structure.results['linear static'].solved = True
        return True

    def modal_solve(self):
        structure = self.structure
        if np.count_nonzero(structure.M_with_masses) == 0:         # Solver.py:34-35
            return False
        out = _bfe2.solve_structure(structure, static=False, modal=True)
        if out['info'] == -1:                                      # mass matrix not positive definite
            return False
        if out['info'] != 0:
            raise np.linalg.LinAlgError('modal solve failed (info %d)' % out['info'])
        circfreq = [math.sqrt(x) for x in out['lam'] if x > 0]     # was: sp.eigh(K, M); Solver.py:42
        shapes = [np.matrix(out['phi'][k]).T for k in range(out['phi'].shape[0])]
        structure.results['modal'] = Results.ModalResult(structure=structure, circular_freq=circfreq, modalshapes=shapes)
        structure.results['modal'].solved = True
        return True

    SolverModule.LinearStaticSolver.solve = static_solve
    SolverModule.ModalSolver.solve = modal_solve
    return SolverModule
