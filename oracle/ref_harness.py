"""Harness that drives the REAL reference (imported read-only from /root/reference)  --  TEST
INFRASTRUCTURE for the build container only.

The reference cannot travel to the GPU box, so nothing under tests/ -m gpu, smoke() or bench.py
imports this module; it is used by oracle/make_golden.py (fixture generation) and by the CPU-only
pinning tests when /root/reference happens to be present.

A "model" here is a plain dict:
    nodes      : list of (x, y)                      node k has ID k+1
    beams      : list of dict(i, j, E, A, I, rho[, lumped])   i, j are 1-based node IDs
    supports   : {nodeID: [dofnames]}
    masses     : list of (nodeID, {mx, my})
    load_cases : list of dict(nodal=[(nodeID, {FX,FY,MZ})], beam=[(beam_index, kind, kwargs)])
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np

REFERENCE_ROOT = os.environ.get('BEAMFE2_REFERENCE_ROOT', '/root/reference')
_STUB = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'mpl_stub')


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'BeamFE2'))


def import_reference():
    """Import the reference package with the matplotlib stub (drawing/__init__.py:3-12 is broken
    without matplotlib) and without writing bytecode into the read-only mount."""
    if not reference_available():
        raise RuntimeError('reference not present at %s' % REFERENCE_ROOT)
    sys.dont_write_bytecode = True
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        if _STUB not in sys.path:
            sys.path.insert(0, _STUB)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    saved = np.get_printoptions()
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        from BeamFE2 import HermitianBeam_2D, Structure, Node, Solver, Results, Loads  # noqa: F401
    np.set_printoptions(**saved)           # BeamFE2/__init__.py:2 changes global print options
    import BeamFE2
    return BeamFE2


def build_structure(model):
    import_reference()
    from BeamFE2 import HermitianBeam_2D as HB
    from BeamFE2 import Structure as ST
    from BeamFE2 import Node as ND
    nodes = [ND.Node(ID=k + 1, coords=(float(x), float(y))) for k, (x, y) in enumerate(model['nodes'])]
    beams = []
    for k, b in enumerate(model['beams']):
        hb = HB.HermitianBeam2D(ID=k + 1, i=nodes[b['i'] - 1], j=nodes[b['j'] - 1], E=b['E'], I=b['I'],
                                A=b['A'], rho=b['rho'])
        if b.get('lumped'):
            hb.mass_matrix = 'lumped'
        beams.append(hb)
    s = ST.Structure(beams=beams, supports={int(k): list(v) for k, v in model['supports'].items()})
    for node_id, mass in model.get('masses', []):
        s.add_nodal_mass(nodeID=node_id, mass=dict(mass))
    return s


def apply_load_case(structure, lc):
    structure.clear_loads()
    for node_id, dyn in lc.get('nodal', []):
        structure.add_nodal_load(nodeID=node_id, dynam=dict(dyn))
    for bidx, kind, kw in lc.get('beam', []):
        structure.add_internal_loads(beam=structure.beams[bidx], loadtype=kind, **kw)


def run_reference(model, modal=True):
    """Run the reference's own solvers; returns numpy arrays in reference layout/order."""
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        s = build_structure(model)
        out = {
            'sumdof': s.sumdof,
            'positions_to_eliminate': np.array(s.positions_to_eliminate, dtype=np.int64),
            'K': np.asarray(s.K), 'M': np.asarray(s.M),
            'Ke_global': np.stack([np.asarray(b.matrix_in_global(b.Ke)) for b in s.beams]),
            'Me_global': np.stack([np.asarray(b.matrix_in_global(b.Me)) for b in s.beams]),
            'length': np.array([b.l for b in s.beams]),
            'direction': np.array([b.direction for b in s.beams]),
        }
        fs, us, efs, rs = [], [], [], []
        for lc in model.get('load_cases', []):
            apply_load_case(s, lc)
            ok = s.solver['linear static'].solve()
            assert ok
            res = s.results['linear static']
            fs.append(np.asarray(s.load_vector).ravel().copy())
            us.append(np.asarray(res.displacement_results[0]).ravel().copy())
            ef = []
            for b in s.beams:
                d = res.element_displacements(local=True, beam=b, asvector=True)
                ef.append(np.asarray(b.nodal_reactions_asvector(disps=d)).ravel())
            efs.append(np.stack(ef))
            rs.append(np.asarray(res.reaction_forces_asvector).ravel().copy())
        if fs:
            out.update(f=np.stack(fs), u=np.stack(us), endforces=np.stack(efs), reactions=np.stack(rs))
        if modal:
            out['M_with_masses'] = np.asarray(s.M_with_masses)
            ok = s.solver['modal'].solve()
            out['modal_ok'] = bool(ok)
            if ok:
                res = s.results['modal']
                out['circular_freq'] = np.array(res.circular_frequencies)
                out['shapes'] = np.stack([np.asarray(x).ravel() for x in res.displacement_results])
    return out


def pack_model(model):
    """model dict -> the packed arrays of oracle/beamfe2_oracle.py (host logic mirrored by
    beamfe2_b200.plan; kept independent here on purpose so the two can be cross-checked)."""
    from oracle import beamfe2_oracle as orc
    nN = len(model['nodes'])
    nE = len(model['beams'])
    coords = np.array(model['nodes'], dtype=np.float64).reshape(nN, 2)
    conn = np.array([[b['i'] - 1, b['j'] - 1] for b in model['beams']], dtype=np.int32)
    props = np.array([[b['E'], b['A'], b['I'], b['rho']] for b in model['beams']], dtype=np.float64)
    lumped = np.array([bool(b.get('lumped', False)) for b in model['beams']])
    fixed = np.array(orc.positions_to_eliminate(model['supports']), dtype=np.int32)
    mass = np.zeros((nN, 2))
    for node_id, m in model.get('masses', []):
        mass[node_id - 1, 0] += m.get('mx', 0)
        mass[node_id - 1, 1] += m.get('my', 0)
    lcs = model.get('load_cases', [])
    C = len(lcs)
    f_nodal = np.zeros((C, 3 * nN))
    r_local = np.zeros((C, nE, 6))
    L, _ = orc.length_direction(coords, conn)
    for c, lc in enumerate(lcs):
        for node_id, dyn in lc.get('nodal', []):
            for k, name in enumerate(orc.LOADNAMES):
                f_nodal[c, (node_id - 1) * 3 + k] += dyn.get(name, 0)
        for bidx, kind, kw in lc.get('beam', []):
            r_local[c, bidx] += orc.beam_load_reactions_local(kind, kw['value'], L[bidx], kw.get('position', 0.0))
    return dict(conn=conn, fixed=fixed, lumped=lumped, coords=coords, props=props, nodal_mass=mass,
                f_nodal=f_nodal, r_local=r_local)
