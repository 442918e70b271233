"""Config 1 (the reference's Examples/cantilever_beam.py with I['x'] -> I.x): one steel cantilever element
with a concentrated axial force and a uniform perpendicular load; static + modal on the GPU."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beamfe2_b200 import HermitianBeam_2D as HB, Structure, Node      # noqa: E402

n1, n2 = Node.Node(ID=1, coords=(0, 0)), Node.Node(ID=2, coords=(200.0, 0))
beam = HB.HermitianBeam2D(ID=1, i=n1, j=n2, E=2.1e5, A=30.0, I=22.5, rho=7.85e-9)
structure = Structure.Structure(beams=[beam], supports={1: ['ux', 'uy', 'rotz']})
structure.add_internal_loads(beam=beam, loadtype='concentrated axial force', value=200, position=0.3)
structure.add_internal_loads(beam=beam, loadtype='uniform perpendicular force', value=-10)
assert structure.solver['linear static'].solve()
res = structure.results['linear static']
print('tip displacements', res.displacement_results[0][3:].T)
print('reactions', res.reaction_forces)
disp = res.element_displacements(local=True, beam=beam, asvector=True)
for pos in (0, 0.25, 0.5, 1):
    print('moment at %.2f: %.6g' % (pos, beam.internal_action(action='moment', disp=disp, pos=pos)))
assert structure.solver['modal'].solve()
print('circular frequencies', structure.results['modal'].circular_frequencies)
