"""Post-processing kernels: internal-action fields / maxima and modal participation
factors, against the REAL reference's HermitianBeam2D.internal_action and the Examples/modal_masses.py
arithmetic (goldens in tests/golden/ref_post.npz, made by oracle/make_golden.py) and the oracle."""
import os

import numpy as np
import pytest
import torch

from beamfe2_b200 import batch, sweep
from beamfe2_b200.plan import Plan
from oracle import beamfe2_oracle as orc
from tests import util

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
POST = dict(np.load(os.path.join(util.GOLD, 'ref_post.npz')))


def T(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=DEV)


def uniform_loads(model, nE):
    C = len(model['load_cases'])
    qa, qp = np.zeros((C, nE)), np.zeros((C, nE))
    for c, lc in enumerate(model['load_cases']):
        for bi, kind, kw in lc.get('beam', []):
            if kind == 'uniform axial force':
                qa[c, bi] += kw['value']
            elif kind == 'uniform perpendicular force':
                qp[c, bi] += kw['value']
            else:
                raise AssertionError('case has non-uniform beam loads')
    return qa, qp


@pytest.mark.parametrize('case', ['p20_v0', 'p20_v1', 'allresults_1', 'allresults_2', 'chopra_105'])
def test_internal_actions_and_participation_vs_reference(case):
    model = util.golden_models()[case]
    pm = util.pack(model)
    nN, nE = len(model['nodes']), len(model['beams'])
    plan = Plan.from_supports(nN, pm['conn'], model['supports'], pm['lumped'])
    out = batch.solve_fused(plan, T(pm['coords'][None]), T(pm['props'][None]), T(pm['nodal_mass'][None]),
                            T(pm['f_nodal'][None]), T(pm['r_local'][None]), n_modes=min(6, plan.n_free))
    qa, qp = uniform_loads(model, nE)
    act, mx = batch.internal_actions(plan, T(pm['coords'][None]), out['endforces'], T(qa[None]), T(qp[None]), npts=5)
    ref = POST[case + '/internal_actions']                       # [C, nE, 3, 5] from beam.internal_action
    got = act[0].cpu().numpy()
    scale = np.abs(ref).max(axis=(1, 3), keepdims=True)
    assert (np.abs(got - ref) / scale).max() < 1e-9
    assert np.allclose(mx[0].cpu().numpy(), np.abs(got).max(axis=-1), rtol=0, atol=0)
    # field-free call only returns the maxima
    none, mx2 = batch.internal_actions(plan, T(pm['coords'][None]), out['endforces'], T(qa[None]), T(qp[None]), npts=5,
                                       want_field=False)
    assert none is None and torch.equal(mx, mx2)
    # participation factors
    full = Plan(nN, pm['conn'], [], pm['lumped'])               # all DOFs: the reference uses the uncondensed M
    Kb, Mb = batch.assemble_banded(full, T(pm['coords'][None]), T(pm['props'][None]), T(pm['nodal_mass'][None]))
    gam = batch.modal_participation(full, Mb, out['phi'])[0].cpu().numpy()
    gref = POST[case + '/gamma'][: gam.shape[0]]
    phi = out['phi'][0].cpu().numpy()
    sref = np.sign(np.sum(util.golden(case)['shapes'][: gam.shape[0]] * phi, axis=1))[:, None]   # sign convention
    assert np.abs(gam * sref - gref).max() / np.abs(gref).max() < 1e-8


def test_post_kernels_batch_vs_oracle():
    B = 2000
    P = sweep.p20_draw_params(B)
    pk = sweep.p20_pack(torch.from_numpy(P).to(DEV))
    conn, sup = sweep.p20_topology()
    plan = Plan.from_supports(21, conn, sup)
    out = batch.solve_fused(plan, pk['coords'], pk['props'], pk['nodal_mass'], pk['f_nodal'], pk['r_local'], n_modes=10)
    s = torch.from_numpy(P[:, 9]).to(DEV)
    qp = torch.zeros((B, 3, 20), dtype=torch.float64, device=DEV)
    qp[:, 1, 7:13] = (-s * 10.0)[:, None]
    qp[:, 2, 7:13] = (-s * 10.0)[:, None]
    act, mx = batch.internal_actions(plan, pk['coords'], out['endforces'], None, qp, npts=11)
    h = {k: v.cpu().numpy() for k, v in pk.items()}
    L, _ = orc.length_direction(h['coords'], conn)
    ref = orc.internal_action_field(out['endforces'].cpu().numpy(), L[:, None, :], None, qp.cpu().numpy(), npts=11)
    assert util.relmax(act.cpu().numpy(), ref) < 1e-12
    assert np.array_equal(mx.cpu().numpy(), np.abs(act.cpu().numpy()).max(axis=-1))
    # moments are continuous across the girder nodes and vanish nowhere special: equilibrium check at a joint
    a = act.cpu().numpy()
    assert np.abs(a[:, :, 8, 2, 0] - a[:, :, 7, 2, -1]).max() / np.abs(a[:, :, :, 2]).max() < 1e-9
    full = Plan(21, conn, [])
    Kb, Mb = batch.assemble_banded(full, pk['coords'], pk['props'], pk['nodal_mass'])
    gam = batch.modal_participation(full, Mb, out['phi']).cpu().numpy()
    o = orc.solve_batch(conn, orc.positions_to_eliminate(sup), h['coords'][:64], h['props'][:64], h['nodal_mass'][:64])
    gref = orc.participation_factors(out['phi'].cpu().numpy()[:64], o['M'])
    assert np.abs(gam[:64] - gref).max() / np.abs(gref).max() < 1e-10
    # sum of effective modal masses over ALL modes equals the total translational mass seen by the free DOFs
    # completeness: with ALL modes, sum_k Gamma_k^2 = (M r)_f^T M_ff^-1 (M r)_f  (f = free DOFs)
    allm = batch.solve_fused(plan, pk['coords'][:32], pk['props'][:32], pk['nodal_mass'][:32], None, None, n_modes=57)
    g_all = batch.modal_participation(full, Mb[:32], allm['phi']).cpu().numpy()
    keep = plan.pos_of_free
    for k in (0, 1):
        r = np.zeros(63)
        r[k::3] = 1.0
        mr = np.einsum('bij,j->bi', o['M'][:32], r)[:, keep]
        Mff = o['M'][:32][:, keep[:, None], keep[None, :]]
        total = np.einsum('bi,bi->b', mr, np.linalg.solve(Mff, mr[..., None])[..., 0])
        assert np.abs((g_all[:, :, k] ** 2).sum(axis=1) / total - 1).max() < 1e-9
