"""Differential test of the drop-in API against the REAL reference (baseline/_ref, CPU): the same random sequence of
public API calls -- nodal loads through both entry points, every beam-load type, nodal masses, `clear=True`, lumped /
consistent mass per beam -- is applied to a random frame built twice, once from the reference's classes and once from
beamfe2_b200's, and everything a user can read back is compared: load / mass vectors, DOF maps, the dense matrices,
displacements, reactions (vector and dict), element displacements and end forces, internal actions, frequencies and
mode shapes.  (The reference's own tests and examples exercise a handful of call sequences; this walks the rest.)"""
import numpy as np
import pytest

from beamfe2_b200 import HermitianBeam_2D as HB
from beamfe2_b200 import Loads, Node, Structure
from oracle import ref_harness as rh
from tests import util

pytestmark = pytest.mark.gpu
EPS = np.finfo(float).eps


class Side(object):
    """The three modules a model is built from, on one side of the comparison."""
    def __init__(self, node, hb, structure):
        self.Node, self.HB, self.Structure = node, hb, structure


def make_frame(rng, n_nodes):
    xy = np.stack([np.arange(n_nodes) * 900.0, 400.0 * np.sin(1.3 * np.arange(n_nodes))], axis=1)
    xy += rng.uniform(-100, 100, xy.shape)
    conn = [(k, k + 1) for k in range(n_nodes - 1)]
    for _ in range(int(rng.integers(0, 3))):                     # a few extra members: wider bands, several beams per node
        a, b = sorted(int(v) for v in rng.choice(n_nodes, 2, replace=False))
        if (a, b) not in conn:
            conn.append((a, b))
    props = [dict(E=float(rng.uniform(1.9e5, 2.2e5)), A=float(rng.uniform(2e3, 2e4)), I=float(rng.uniform(1e6, 1e8)),
                  rho=float(rng.uniform(7e-9, 8.5e-9))) for _ in conn]
    supports = {1: ['ux', 'uy', 'rotz']}
    if n_nodes > 3 and rng.uniform() < 0.7:
        supports[n_nodes] = [['uy'], ['ux', 'uy'], ['ux', 'uy', 'rotz']][int(rng.integers(0, 3))]
    return xy, conn, props, supports


def make_ops(rng, n_nodes, n_beams):
    ops = []
    kinds = list(Loads.BEAM_LOAD_TYPES)
    for _ in range(int(rng.integers(4, 12))):
        r = rng.uniform()
        node = int(rng.integers(2, n_nodes + 1))
        if r < 0.25:
            names = [n for n in ('FX', 'FY', 'MZ') if rng.uniform() < 0.6] or ['FY']
            ops.append(('nodal_load', node, {n: float(rng.uniform(-1e4, 1e4)) for n in names}, bool(rng.uniform() < 0.15)))
        elif r < 0.4:
            names = [n for n in ('FX', 'FY', 'MZ') if rng.uniform() < 0.5] or ['FX']
            ops.append(('single_dynam', node, {n: float(rng.uniform(-1e4, 1e4)) for n in names}, bool(rng.uniform() < 0.1)))
        elif r < 0.75:
            kind = kinds[int(rng.integers(0, len(kinds)))]
            kw = dict(value=float(rng.uniform(-50, 50)))
            if kind.startswith('concentrated'):
                kw['position'] = float(rng.uniform(0.05, 0.95))
            ops.append(('beam_load', int(rng.integers(0, n_beams)), kind, kw))
        elif r < 0.9:
            names = [n for n in ('mx', 'my') if rng.uniform() < 0.7] or ['mx']
            ops.append(('nodal_mass', node, {n: float(rng.uniform(0.1, 5.0)) for n in names}, bool(rng.uniform() < 0.1)))
        else:
            ids = 'all' if rng.uniform() < 0.5 else [int(v) + 1 for v in rng.choice(n_beams, max(1, n_beams // 2), replace=False)]
            ops.append(('mass_type', ['lumped', 'consistent'][int(rng.integers(0, 2))], ids))
    ops.append(('nodal_load', n_nodes if n_nodes > 2 else 2, {'FX': 1234.5, 'FY': -678.9}, False))   # never an empty load vector
    return ops


def build(side, xy, conn, props, supports, ops):
    nodes = [side.Node.Node(ID=k + 1, coords=(float(x), float(y))) for k, (x, y) in enumerate(xy)]
    beams = [side.HB.HermitianBeam2D(ID=k + 1, i=nodes[a], j=nodes[b], **p) for k, ((a, b), p) in enumerate(zip(conn, props))]
    st = side.Structure.Structure(beams=beams, supports={k: list(v) for k, v in supports.items()})
    for op in ops:
        if op[0] == 'nodal_load':
            st.add_nodal_load(nodeID=op[1], dynam=dict(op[2]), clear=op[3])
        elif op[0] == 'single_dynam':
            st.add_single_dynam_to_node(nodeID=op[1], dynam=dict(op[2]), clear=op[3])
        elif op[0] == 'beam_load':
            st.add_internal_loads(beam=st.beams[op[1]], loadtype=op[2], **op[3])
        elif op[0] == 'nodal_mass':
            st.add_nodal_mass(nodeID=op[1], mass=dict(op[2]), clear=op[3])
        elif op[0] == 'mass_type':
            st.set_mass_matrix_type(matrixtype=op[1], beam_IDs=op[2])
    return st


def arr(m):
    return np.asarray(m, dtype=np.float64)


def rel(a, b):
    a, b = arr(a), arr(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300) if a.size else 0.0


@pytest.mark.parametrize('seed', range(48))
def test_random_api_sequences_agree_with_the_reference(seed):
    if not rh.reference_available():
        pytest.skip('baseline/_ref is not installed')
    ref = rh.import_reference()
    rng = np.random.default_rng(9000 + seed)
    n_nodes = int(rng.integers(3, 11))
    xy, conn, props, supports = make_frame(rng, n_nodes)
    ops = make_ops(rng, n_nodes, len(conn))
    theirs = build(Side(ref.Node, ref.HermitianBeam_2D, ref.Structure), xy, conn, props, supports, ops)
    mine = build(Side(Node, HB, Structure), xy, conn, props, supports, ops)
    # ---- bookkeeping: bit-exact
    assert mine.sumdof == theirs.sumdof and list(mine.positions_to_eliminate) == list(theirs.positions_to_eliminate)
    assert np.array_equal(arr(mine.load_vector), arr(theirs.load_vector)), ops
    assert (mine.mass_vector is None) == (theirs.mass_vector is None)           # no nodal mass was added
    assert mine.mass_vector is None or np.array_equal(arr(mine.mass_vector), arr(theirs.mass_vector)), ops
    assert [b.mass_matrix for b in mine.beams] == [b.mass_matrix for b in theirs.beams]
    assert len(mine.nodal_loads) == len(theirs.nodal_loads) and len(mine.nodal_masses) == len(theirs.nodal_masses)
    for bm, bt in zip(mine.beams, theirs.beams):
        assert len(bm.internal_loads) == len(bt.internal_loads)
        assert rel(bm.reduced_internal_loads_asvector, bt.reduced_internal_loads_asvector) < 1e-14
        assert list(bm.internal_action_points) == list(bt.internal_action_points)
    # ---- matrices (GPU element + assembly kernels vs the reference's numpy)
    for name in ('K', 'M', 'M_with_masses', 'K_with_BC'):
        assert rel(getattr(mine, name), getattr(theirs, name)) < 5e-14, name
    for name in ('node_numbering_is_ok', 'stiffness_matrix_is_symmetric', 'stiffness_matrix_is_nonsingular',
                 'stiffness_matrix_is_ok', 'mass_matrix_is_ok'):
        assert getattr(mine, name) is True and getattr(theirs, name) is True, name
    Kc = arr(theirs.condense(theirs.K))
    tol = max(1e-9, 100 * EPS * np.linalg.cond(Kc))              # the reference inverts the penalty matrix
    # ---- static
    assert theirs.solver['linear static'].solve() is True and mine.solver['linear static'].solve() is True
    rm, rt = mine.results['linear static'], theirs.results['linear static']
    assert rel(rm.displacement_results[0], rt.displacement_results[0]) < tol
    assert rel(rm.reaction_forces_asvector, rt.reaction_forces_asvector) < tol, ops
    dm, dt = rm.reaction_forces, rt.reaction_forces
    assert sorted(dm) == sorted(dt)
    scale = np.abs(arr(rt.reaction_forces_asvector)).max()
    for nid in dt:
        assert sorted(dm[nid]) == sorted(dt[nid])
        for k in dt[nid]:
            assert abs(dm[nid][k] - dt[nid][k]) <= tol * scale + 2e-5      # rounded to 5 decimals on both sides
    for bm, bt in zip(mine.beams, theirs.beams):
        dl_m = rm.element_displacements(local=True, beam=bm, asvector=True)
        dl_t = rt.element_displacements(local=True, beam=bt, asvector=True)
        assert rel(dl_m, dl_t) < tol
        assert rel(rm.element_displacements(local=False, beam=bm, asvector=True),
                   rt.element_displacements(local=False, beam=bt, asvector=True)) < tol
        part_m, part_t = rm.element_displacements(local=True, beam=bm), rt.element_displacements(local=True, beam=bt)
        assert sorted(part_m) == sorted(part_t) and all(rel(part_m[k], part_t[k]) < tol or np.abs(arr(part_t[k])).max() < tol for k in part_t)
        fm, ft = arr(bm.nodal_reactions_asvector(disps=dl_m)), arr(bt.nodal_reactions_asvector(disps=dl_t))
        fscale = max(np.abs(ft).max(), scale)
        assert np.abs(fm - ft).max() <= tol * fscale
        for action in ('axial', 'shear', 'moment'):
            for pos in (0.0, 0.25, 0.5, 0.9, 1.0):
                a_m = bm.internal_action(action=action, disp=dl_m, pos=pos)
                a_t = bt.internal_action(action=action, disp=dl_t, pos=pos)
                assert abs(a_m - a_t) <= tol * max(fscale, fscale * bt.l if action == 'moment' else fscale), (action, pos)
    # ---- modal.  With lumped beams the reference's mass matrix carries m * 1e-10 on the rotations
    #      (HermitianBeam_2D.py:225-231): lam_max / lam_min ~ 1e18 and scipy.linalg.eigh returns low modes with
    #      residuals of order one, sometimes a negative eigenvalue that the frequency filter then drops (SURVEY.md
    #      7.3-3).  So every mode is first judged by its own residual |K v - lam M v| / |K v| on the reference's
    #      matrices; the two sides are compared where the reference resolved the pair, and the GPU side has to resolve
    #      all of them.
    assert theirs.solver['modal'].solve() is True and mine.solver['modal'].solve() is True
    mm, mt = mine.results['modal'], theirs.results['modal']
    fa, fb = np.array(mm.frequencies), np.array(mt.frequencies)
    Kc, Mc = arr(theirs.condense(theirs.K_with_BC)), arr(theirs.condense(theirs.M_with_masses))
    keep = np.setdiff1d(np.arange(theirs.sumdof), theirs.positions_to_eliminate)

    def residuals(res, f):
        out = []
        for k in range(len(f)):
            v, lam = arr(res.modeshape(k)).ravel()[keep], (2 * np.pi * f[k]) ** 2
            out.append(np.abs(Kc @ v - lam * (Mc @ v)).max() / np.abs(Kc @ v).max())
        return np.array(out)
    res_m, res_t = residuals(mm, fa), residuals(mt, fb)
    lumped = any(b.mass_matrix == 'lumped' for b in theirs.beams)
    assert len(fa) == len(keep)
    if lumped:
        # a residual cannot be EVALUATED in double precision on such a pencil (|K v| for a low mode is 1e-18 of
        # |K| |v|): the judge is a 60-digit solution of the same double-precision matrices, for every mode
        ft = np.sqrt(util.truth_eigenvalues(Kc, Mc)) / (2 * np.pi)
        assert np.abs(fa / ft - 1).max() < 1e-9, np.abs(fa / ft - 1)
    else:
        assert (res_m < 1e-8).all(), res_m
    assert np.allclose(mm.periods, 1.0 / fa) and len(mm.displacement_results) == len(keep)
    if not lumped:
        assert len(fb) == len(fa) and (res_t < 1e-6).all()
    if len(fb) == len(fa):
        ftol = np.maximum(1e-8, 25 * EPS * (fb.max() / fb) ** 2)     # LAPACK's accuracy is relative to the largest eigenvalue
        lam = (2 * np.pi * fb) ** 2
        gap = np.minimum(np.r_[np.inf, np.diff(lam) / lam[1:]], np.r_[np.diff(lam) / lam[1:], np.inf])
        for k in np.where(res_t < 1e-9)[0]:
            assert abs(fa[k] / fb[k] - 1) < ftol[k], (k, fa[k], fb[k])
            if gap[k] > 1e-3:                                      # isolated: the vector is defined up to its sign
                assert util.mode_err(arr(mm.modeshape(k)).ravel(), arr(mt.modeshape(k)).ravel()) < 1e-6 / min(gap[k], 1.0), k
    Mfull = arr(theirs.M_with_masses)
    for k in range(len(fa)):
        v = arr(mm.modeshape(k)).ravel()
        assert abs(v @ Mfull @ v - 1) < 1e-8


def test_patched_reference_agrees_with_itself():
    """The same random models through the unmodified reference and through the reference with only its two solve()
    bodies replaced by the ctypes stub (INTEGRATION.md option B) -- in a child process, because the patch is global."""
    import os
    import subprocess
    import sys
    if not rh.reference_available():
        pytest.skip('baseline/_ref is not installed')
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE='1', PYTHONPATH=util.ROOT)
    r = subprocess.run([sys.executable, '-m', 'tests.differential_patch_child', '24'], cwd=util.ROOT, env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and 'OK 24 models' in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
