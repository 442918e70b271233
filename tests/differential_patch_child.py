"""Child process of tests/test_gpu_differential.py::test_patched_reference_agrees_with_itself: the UNMODIFIED reference
solves a set of random models on the CPU, then its two solve() bodies are replaced by the ctypes stub of
INTEGRATION.md option B (integration/solver_patch.py) and the same models, rebuilt by the same API calls, are solved
again -- everything else (Structure, Results, Loads) stays the reference's own code."""
import sys

import numpy as np

from oracle import ref_harness as rh
from tests import test_gpu_differential as T

EPS = np.finfo(float).eps
ref = rh.import_reference()
side = T.Side(ref.Node, ref.HermitianBeam_2D, ref.Structure)


def models(seeds):
    for seed in seeds:
        rng = np.random.default_rng(7000 + seed)
        n_nodes = int(rng.integers(3, 11))
        xy, conn, props, supports = T.make_frame(rng, n_nodes)
        ops = T.make_ops(rng, n_nodes, len(conn))
        yield seed, T.build(side, xy, conn, props, supports, ops)


def snapshot(st):
    assert st.solver['linear static'].solve() is True and st.solver['modal'].solve() is True
    ls, mo = st.results['linear static'], st.results['modal']
    keep = np.setdiff1d(np.arange(st.sumdof), st.positions_to_eliminate)
    Kc, Mc = T.arr(st.condense(st.K_with_BC)), T.arr(st.condense(st.M_with_masses))
    f = np.array(mo.frequencies)
    res = []
    for k in range(len(f)):
        v, lam = T.arr(mo.modeshape(k)).ravel()[keep], (2 * np.pi * f[k]) ** 2
        res.append(np.abs(Kc @ v - lam * (Mc @ v)).max() / np.abs(Kc @ v).max())
    beams = []
    for b in st.beams:
        d = ls.element_displacements(local=True, beam=b, asvector=True)
        beams.append((T.arr(d).ravel(), T.arr(b.nodal_reactions_asvector(disps=d)).ravel(),
                      [b.internal_action(action=a, disp=d, pos=p) for a in ('axial', 'shear', 'moment') for p in (0.0, 0.4, 1.0)]))
    lumped = any(b.mass_matrix == 'lumped' for b in st.beams)
    truth = np.sqrt(T.util.truth_eigenvalues(Kc, Mc)) / (2 * np.pi) if lumped else None
    return dict(truth=truth, u=T.arr(ls.displacement_results[0]).ravel(), react=T.arr(ls.reaction_forces_asvector).ravel(),
                rdict=ls.reaction_forces, f=f, res=np.array(res), beams=beams, n_free=len(keep),
                cond=np.linalg.cond(T.arr(st.condense(st.K))), lumped=lumped)


seeds = range(int(sys.argv[1]) if len(sys.argv) > 1 else 24)
before = {seed: snapshot(st) for seed, st in models(seeds)}
from integration import solver_patch                                     # noqa: E402
solver_patch.apply(ref.Solver, ref.Results)
for seed, st in models(seeds):
    a, b = snapshot(st), before[seed]
    tol = max(1e-9, 100 * EPS * b['cond'])
    assert T.rel(a['u'], b['u']) < tol and T.rel(a['react'], b['react']) < tol, seed
    scale = np.abs(b['react']).max()
    for nid in b['rdict']:
        for k in b['rdict'][nid]:
            assert abs(a['rdict'][nid][k] - b['rdict'][nid][k]) <= tol * scale + 2e-5, seed
    for (da, fa, ia), (db, fb, ib) in zip(a['beams'], b['beams']):
        fs = max(np.abs(fb).max(), scale)
        assert T.rel(da, db) < tol and np.abs(fa - fb).max() <= tol * fs, seed
        assert np.abs(np.array(ia) - np.array(ib)).max() <= tol * fs * 1e4, seed      # moments carry a length
    assert len(a['f']) == a['n_free'], seed
    if a['lumped']:                       # judged by a 60-digit solution of the same matrices (see test_gpu_differential)
        assert np.abs(a['f'] / a['truth'] - 1).max() < 1e-9, (seed, np.abs(a['f'] / a['truth'] - 1))
    else:
        assert (a['res'] < 1e-8).all(), (seed, a['res'])
    if len(a['f']) == len(b['f']):
        ftol = np.maximum(1e-8, 25 * EPS * (b['f'].max() / b['f']) ** 2)
        ok = b['res'] < 1e-9
        assert (np.abs(a['f'] / b['f'] - 1)[ok] < ftol[ok]).all(), seed
print('OK %d models' % len(before))
