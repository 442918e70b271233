#!/usr/bin/env python
"""A multi-storey frame with more than 116 free DOFs through the drop-in API: the static solve runs on the normal
batched path, the modal solve takes the first-k route (frame_modes.FrameModes, lowest 40 modes)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beamfe2_b200 import HermitianBeam_2D as HB, Node, Structure     # noqa: E402

storeys, bays, H, L = 4, 5, 3000.0, 5000.0
nodes, index = [], {}


def node(x, y):
    key = (round(x, 6), round(y, 6))
    if key not in index:
        index[key] = Node.Node(ID=len(nodes) + 1, coords=(x, y))
        nodes.append(index[key])
    return index[key]


beams = []
for b in range(bays + 1):
    for s in range(storeys):
        for k in range(2):
            beams.append(HB.HermitianBeam2D(ID=len(beams) + 1, E=2.1e5, rho=7.85e-9, A=300.0 * 300.0, I=300.0 ** 4 / 12,
                                            i=node(b * L, s * H + k * H / 2), j=node(b * L, s * H + (k + 1) * H / 2)))
for s in range(1, storeys + 1):
    for b in range(bays):
        for k in range(2):
            beams.append(HB.HermitianBeam2D(ID=len(beams) + 1, E=2.1e5, rho=7.85e-9, A=200.0 * 500.0,
                                            I=200.0 * 500.0 ** 3 / 12,
                                            i=node(b * L + k * L / 2, s * H), j=node(b * L + (k + 1) * L / 2, s * H)))
supports = {n.ID: ['ux', 'uy', 'rotz'] for n in nodes if n.y == 0.0}
st = Structure.Structure(beams=beams, supports=supports)
for n in nodes:
    if n.y > 0 and abs(n.y / H - round(n.y / H)) < 1e-9:
        st.add_nodal_mass(nodeID=n.ID, mass={'mx': 2.0, 'my': 2.0})
top = max(nodes, key=lambda n: (n.y, -n.x))
st.add_nodal_load(nodeID=top.ID, dynam={'FX': 5.0e4}, clear=True)
print('nodes %d, beams %d, free DOFs %d' % (len(nodes), len(beams), 3 * len(nodes) - 3 * len(supports)))
assert st.solver['linear static'].solve()
u = st.results['linear static'].displacement_results[0]
print('roof sway under 50 kN: %.3f mm' % float(np.asarray(u).ravel()[3 * (top.ID - 1)]))
assert st.solver['modal'].solve()
print('first five frequencies [Hz]:', ['%.3f' % f for f in st.results['modal'].frequencies[:5]])
