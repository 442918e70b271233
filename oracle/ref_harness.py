"""Harness that drives the REAL reference  --  TEST / BASELINE INFRASTRUCTURE, never the product path.

The reference is imported from the read-only checkout in the build container, or -- on the GPU box, where that
checkout does not exist -- from baseline/_ref, the git-ignored `pip install --target` copy that
install_into_baseline() makes (called by __graft_entry__.build()).  Users: oracle/make_golden.py (fixture
generation), the CPU-only pinning tests, bench.py's reference arm (A) (object_path_seconds below) and
tests/test_reference_suite.py (the reference's own Tests/ run through the GPU boundary).

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

REFERENCE_SOURCE = '/root/reference'                       # read-only checkout, build container only
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASELINE_REF = os.path.join(_REPO, 'baseline', '_ref')     # git-ignored install of the reference; travels to the GPU box
_STUB = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'mpl_stub')


def _resolve_root():
    env = os.environ.get('BEAMFE2_REFERENCE_ROOT')
    if env:
        return env
    if os.path.isdir(os.path.join(REFERENCE_SOURCE, 'BeamFE2')):
        return REFERENCE_SOURCE
    return BASELINE_REF


REFERENCE_ROOT = _resolve_root()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'BeamFE2')) and os.path.isdir(os.path.join(REFERENCE_ROOT, 'drawing'))


def install_into_baseline(verbose=False) -> bool:
    """`pip install --target baseline/_ref` of the UNMODIFIED reference (the one offline install the task allows),
    plus its `drawing` package (BeamFE2 imports it, setup.py does not list it), its own `Tests/` and its `Examples/`
    scripts, so that the reference arm of bench.py, the reference's test-suite and its examples can run on the GPU box, where the read-only checkout
    does not exist.  baseline/_ref is git-ignored: nothing of the reference enters the history.
    Returns True when baseline/_ref is complete."""
    import shutil
    import subprocess
    import tempfile
    have = all(os.path.isdir(os.path.join(BASELINE_REF, d)) for d in ('BeamFE2', 'drawing', 'Tests', 'Examples'))
    if have or not os.path.isdir(os.path.join(REFERENCE_SOURCE, 'BeamFE2')):
        return have
    os.makedirs(BASELINE_REF, exist_ok=True)
    if not os.path.isdir(os.path.join(BASELINE_REF, 'BeamFE2')):
        with tempfile.TemporaryDirectory() as tmp:         # setup.py writes egg-info next to itself: install from a copy
            src = os.path.join(tmp, 'ref')
            shutil.copytree(REFERENCE_SOURCE, src)
            r = subprocess.run([sys.executable, '-m', 'pip', 'install', '--no-index', '--no-build-isolation', '--no-deps',
                                '--find-links', '/opt/wheelhouse', '--target', BASELINE_REF, src],
                               capture_output=True, text=True)
            if verbose or r.returncode != 0:
                sys.stderr.write(r.stdout[-2000:] + r.stderr[-2000:])
            if r.returncode != 0:
                return False
    for d in ('drawing', 'Tests', 'Examples'):
        if not os.path.isdir(os.path.join(BASELINE_REF, d)):
            shutil.copytree(os.path.join(REFERENCE_SOURCE, d), os.path.join(BASELINE_REF, d),
                            ignore=shutil.ignore_patterns('__pycache__'))
    return True


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


def object_path_seconds(models):
    """CPU baseline arm (A) of SURVEY.md 8(d): the reference's own object path, timed.  Per model: build
    Node / HermitianBeam2D / Structure objects, every load case = clear + reload + solver['linear static'].solve()
    + reaction_forces_asvector, then solver['modal'].solve() (Solver.py:29-87, Results.py:92-127).
    Returns the seconds spent (imports excluded)."""
    import time
    import_reference()
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        t0 = time.perf_counter()
        for model in models:
            s = build_structure(model)
            for lc in model.get('load_cases', []):
                apply_load_case(s, lc)
                assert s.solver['linear static'].solve()
                s.results['linear static'].reaction_forces_asvector
            assert s.solver['modal'].solve()
        return time.perf_counter() - t0


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
