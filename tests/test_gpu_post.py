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


def beam_loads(model, nE, n_conc=2):
    """uniform q_axial / q_perp [C, nE] and concentrated-load descriptors [C, nE, n_conc, 3] of a golden model"""
    C = len(model['load_cases'])
    qa, qp, conc = np.zeros((C, nE)), np.zeros((C, nE)), np.zeros((C, nE, n_conc, 3))
    used = np.zeros((C, nE), dtype=int)
    for c, lc in enumerate(model['load_cases']):
        for bi, kind, kw in lc.get('beam', []):
            if kind == 'uniform axial force':
                qa[c, bi] += kw['value']
            elif kind == 'uniform perpendicular force':
                qp[c, bi] += kw['value']
            else:
                conc[c, bi, used[c, bi]] = (batch.LOAD_KINDS[kind], kw['value'], kw.get('position', 0.0))
                used[c, bi] += 1
    return qa, qp, conc


@pytest.mark.parametrize('case', ['clamped_clamped_loads', 'vertical_beam', 'cantilever_example'])
def test_concentrated_load_curves_vs_reference(case):
    """N / V / M fields of beams carrying the three CONCENTRATED load types (Loads.py:347-365, 421-432,
    489-502), against HermitianBeam2D.internal_action of the real reference at 11 positions (xi = 0.3, the load
    position of two of the cases, is sampled exactly: the reference's comparisons at xi == position are pinned), and
    the per-beam maxima, which must see the one-sided limits at the load positions."""
    model = util.golden_models()[case]
    pm = util.pack(model)
    nN, nE = len(model['nodes']), len(model['beams'])
    plan = Plan.from_supports(nN, pm['conn'], model['supports'], pm['lumped'])
    out = batch.solve_fused(plan, T(pm['coords'][None]), T(pm['props'][None]), None, T(pm['f_nodal'][None]),
                            T(pm['r_local'][None]), modal=False)
    assert int(out['info'][0]) == 0
    qa, qp, conc = beam_loads(model, nE)
    act, mx = batch.internal_actions(plan, T(pm['coords'][None]), out['endforces'], T(qa[None]), T(qp[None]), npts=11,
                                     conc=T(conc[None]))
    ref = POST[case + '/internal_actions']                       # [C, nE, 3, 11]
    got = act[0].cpu().numpy()
    # relative to the largest force / moment of the load case: a vertical member has N = 6e-17 in the reference
    # (cos(arctan2(dy, 0)) is not exactly 0, helpers.py:47-48) and exactly 0 here
    scale = np.abs(ref).max(axis=(1, 2, 3), keepdims=True)
    assert (np.abs(got - ref) / scale).max() < 1e-9
    # maxima: at least the sampled ones; where a load sits between / on samples, the one-sided limits as well
    mxn = mx[0].cpu().numpy()
    assert (mxn >= np.abs(got).max(axis=-1) - 1e-12 * scale[..., 0]).all()
    if case == 'vertical_beam':                              # M = 7 at 0.25: -4.375 | +2.625 around the load
        assert abs(mxn[0, 0, 2] - 4.375) < 1e-9
    if case == 'clamped_clamped_loads':
        # LC 2: P = -30 at 0.3 of a clamped-clamped beam: |V| jumps from 23.52 to 6.48 at the load; the moment kink
        # under the load, M(0.3) = 441 - 23.52 * 30 = -264.6, is a sampled point; LC 3: M = -200 at 0.3: the moment
        # line climbs 14 ... 89.6 and drops by 200 AT the load: the reference returns the left value there, the largest
        # |M| = 110.4 is the right-hand limit, which only the one-sided evaluation sees
        assert abs(mxn[1, 0, 1] - 23.52) < 1e-9 and abs(mxn[1, 0, 2] - 441.0) < 1e-9
        left = ref[2, 0, 2, 3]                                    # reference value at xi = 0.3 (its `xi > pos` is False)
        assert abs(mxn[2, 0, 2] - max(np.abs(ref[2, 0, 2]).max(), abs(left), abs(left + (-200)))) < 1e-9


def test_concentrated_load_maxima_between_samples():
    """A load position that no sample hits: the extreme shear / moment only shows in the one-sided limits."""
    model = util.golden_models()['clamped_clamped_loads']
    pm = util.pack(model)
    plan = Plan.from_supports(2, pm['conn'], model['supports'], pm['lumped'])
    # simply supported reading of the same beam: end forces of P at a = 0.37 on a pinned-pinned span are
    # (V1, V2) = (b P, a P), M1 = M2 = 0; feed them directly
    P, a, L = 10.0, 0.37, 100.0
    ef = np.zeros((1, 1, 1, 6))
    ef[0, 0, 0] = [0.0, (1 - a) * P, 0.0, 0.0, a * P, 0.0]
    conc = np.zeros((1, 1, 1, 1, 3))
    conc[0, 0, 0, 0] = (batch.LOAD_KINDS['concentrated perpendicular force'], -P, a)
    act, mx = batch.internal_actions(plan, T(pm['coords'][None]), T(ef), None, None, npts=5, conc=T(conc))
    got, mxn = act[0, 0, 0].cpu().numpy(), mx[0, 0, 0].cpu().numpy()
    # the largest moment sits under the load: M = a b P L = 233.1 -- none of the 5 samples sees it
    assert abs(mxn[2] - a * (1 - a) * P * L) < 1e-9 and np.abs(got[2]).max() < mxn[2] - 1.0
    assert abs(mxn[1] - (1 - a) * P) < 1e-12


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
