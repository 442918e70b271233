"""Child process of tests/test_reference_examples.py: runs ONE of the reference's Examples/*.py UNCHANGED, with only
Structure.draw neutralised (it needs matplotlib; the 'no results' message of an unsolved analysis is kept), and
prints what the script left behind as JSON.

    python tests/run_example_child.py real|alias|patch <path to example> [--fix]

--fix : four of the seven examples index the section's moment of inertia as I['x'] where BeamSections returns a
        namedtuple (the reference's own example bug); with --fix the text is run with I['x'] -> I.x,
        replaced in memory.

real  : the unmodified reference (CPU)          alias : BeamFE2.* = the beamfe2_b200 drop-in modules (GPU)
patch : the unmodified reference with the two solve() bodies replaced by the ctypes stub of INTEGRATION.md (GPU)"""
import contextlib
import io
import json
import os
import sys

import numpy as np

mode, path = sys.argv[1], sys.argv[2]
fix = '--fix' in sys.argv[3:]
if mode == 'patch':
    import BeamFE2.Results
    import BeamFE2.Solver
    from integration import solver_patch
    solver_patch.apply(BeamFE2.Solver, BeamFE2.Results)
from BeamFE2 import Structure                      # noqa: E402  (resolved through PYTHONPATH: shim or the real package)


def _draw(self, show=True, analysistype=None, mode=0, internal_action=None):
    # Structure.draw (Structure.py:110-114) draws when the analysis has been solved and says so when not
    if not self.results[analysistype].solved:
        print('no results available, no printing')


Structure.Structure.draw = _draw
out = {'mode': mode, 'example': os.path.basename(path)}
buf = io.StringIO()
g = {'__name__': '__main__', '__file__': path}
try:
    with open(path) as fh:
        src = fh.read()
    if fix:
        src = src.replace("I['x']", 'I.x')
    with contextlib.redirect_stdout(buf):
        exec(compile(src, path, 'exec'), g)
except SystemExit:                                                       # cantilever_beam.py ends with exit()
    pass
except Exception as ex:                                                  # some examples are broken in the reference itself
    out['error'] = type(ex).__name__ + ': ' + str(ex)[:200]
st = g.get('structure')
if st is not None:                                                       # whatever was solved before the script ended
    ls, mo = st.results.get('linear static'), st.results.get('modal')
    if ls is not None and getattr(ls, 'displacement_results', None):
        out['u'] = np.asarray(ls.displacement_results[0]).ravel().tolist()
        out['reactions'] = np.asarray(ls.reaction_forces_asvector).ravel().tolist()
    if mo is not None and getattr(mo, 'displacement_results', None):
        out['frequencies'] = [float(f) for f in mo.frequencies]
        M = np.asarray(st.M_with_masses)
        shapes = [np.asarray(s).ravel() for s in mo.displacement_results]
        out['modal_mass'] = [float(s @ M @ s) for s in shapes]          # 1 for M-normalised shapes
        out['shapes'] = [s.tolist() for s in shapes]
out['stdout'] = buf.getvalue()
out['stdout_lines'] = len(buf.getvalue().splitlines())
print('RESULT ' + json.dumps(out))
