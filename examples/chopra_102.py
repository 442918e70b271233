"""Config 2 (Anil K. Chopra, Example 10.2, as in the reference's Examples/Chopra_102.py): cantilever of two
elements with lumped nodal masses, modal analysis through the drop-in API on the GPU.
    python examples/chopra_102.py        (needs a CUDA device)"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beamfe2_b200 import HermitianBeam_2D as HB, Structure, Node      # noqa: E402

L, mass, rho, EE, I, A = 10, 100000 / 386., 0.1, 29000, 320, 63.41
nodes = [Node.Node.from_dict({'ID': k + 1, 'coords': (k * L / 2., 0)}) for k in range(3)]
beams = [HB.HermitianBeam2D.from_dict({'ID': k + 1, 'E': EE, 'I': I, 'A': A, 'rho': rho, 'i': nodes[k], 'j': nodes[k + 1]})
         for k in range(2)]
structure = Structure.Structure(beams=beams, supports={1: ['ux', 'uy', 'rotz']})
structure.add_nodal_mass(nodeID=2, mass={'mx': mass * L / 2., 'my': mass * L / 2.})
structure.add_nodal_mass(nodeID=3, mass={'mx': mass * L / 4., 'my': mass * L / 4.})
assert structure.solver['modal'].solve()
f = structure.results['modal'].frequencies
for mode, coeff in ((0, 3.15623), (2, 16.258)):
    expected = coeff * math.sqrt(EE * I / (mass * L ** 4)) / (2 * math.pi)
    print('mode %d: textbook %.3f Hz, computed %.3f Hz' % (mode + 1, expected, f[mode]))
print('positions_to_eliminate', structure.positions_to_eliminate)
