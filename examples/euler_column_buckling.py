"""Linear buckling of a pinned-pinned column through the drop-in API (the reference's BucklingSolver, Solver.py:90-180,
is unfinished): 30 elements, unit compressive load, first three critical loads against n^2 pi^2 EI / L^2."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beamfe2_b200 import HermitianBeam_2D as HB, Structure, Node      # noqa: E402

E, A, I, L, n_el = 2.1e5, 30.0, 22.5, 900.0, 30
nodes = [Node.Node(ID=k + 1, coords=(L * k / n_el, 0.0)) for k in range(n_el + 1)]
beams = [HB.HermitianBeam2D(ID=k + 1, E=E, I=I, A=A, rho=7.85e-9, i=nodes[k], j=nodes[k + 1]) for k in range(n_el)]
column = Structure.Structure(beams=beams, supports={1: ['ux', 'uy'], n_el + 1: ['uy']})
column.add_nodal_load(nodeID=n_el + 1, dynam={'FX': -1.0}, clear=True)
column.solver['buckling'].nr_modes = 3
assert column.solver['buckling'].solve()
euler = math.pi ** 2 * E * I / L ** 2
for k, pcr in enumerate(column.results['buckling'].criticals, start=1):
    print('mode %d: critical load %.6f  (Euler %.6f, ratio %.9f)' % (k, pcr, k * k * euler, pcr / (k * k * euler)))
