"""GPU parity tests (-m gpu): every CUDA kernel, through the C ABI, against the CPU oracle, the
golden vectors of the real reference and the 40-digit truth fixture.

Tolerances (north_star): 1e-10 relative for displacements, eigenvalues and M-normalised modes (modes
up to sign); integer maps bit-exact (tests/test_plan_abi.py).  Where scipy's own FP64 error exceeds
1e-10 (small eigenvalues of a pencil with lam_max/lam_min ~ 1e7) the comparison with scipy uses
LAPACK's error bound and the 1e-10 claim is checked against the 40-digit truth instead."""
import math

import numpy as np
import pytest
import torch

from beamfe2_b200 import batch, sweep
from beamfe2_b200.plan import Plan
from oracle import beamfe2_oracle as orc
from tests import util

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
EPS = np.finfo(float).eps


def T(a, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device=DEV)


def gpu_model(model, modal=True, n_modes=None, fused=True):
    """One model dict through the GPU path as a batch of 1."""
    pm = util.pack(model)
    nN = len(model['nodes'])
    plan = Plan.from_supports(nN, pm['conn'], model['supports'], pm['lumped'])
    C = pm['f_nodal'].shape[0]
    kw = dict(coords=T(pm['coords'][None]), props=T(pm['props'][None]), nodal_mass=T(pm['nodal_mass'][None]))
    if C:
        kw.update(f_nodal=T(pm['f_nodal'][None]), r_local=T(pm['r_local'][None]))
    out = batch.solve_fused(plan, modal=modal and plan.n_free > 0, n_modes=n_modes, **kw)
    torch.cuda.synchronize()
    return plan, pm, {k: (v[0].cpu().numpy() if v is not None else None) for k, v in out.items()}


def p20_batch(B, shard=None):
    P = sweep.p20_draw_params(B, shard)
    pk = sweep.p20_pack(torch.from_numpy(P).to(DEV))
    conn, sup = sweep.p20_topology()
    plan = Plan.from_supports(21, conn, sup)
    return P, pk, plan


def oracle_p20(pk, n_modes=None):
    conn, sup = sweep.p20_topology()
    h = {k: v.cpu().numpy() for k, v in pk.items()}
    return orc.solve_batch(conn, orc.positions_to_eliminate(sup), h['coords'], h['props'], h['nodal_mass'],
                           h['f_nodal'], h['r_local'], n_modes=n_modes)


# ------------------------------------------------------------------------------------------- K1
def test_k1_element_known_answers():
    """Tests/test_beam_linear_elastic.py:37-46 (unit beam integer matrix) and :140-151 (rotated)."""
    one = T([1.0])
    Kg, Mg = batch.element_km(T([0.]), T([0.]), one, T([0.]), one, one, one, one)
    exp = np.array([[1, 0, 0, -1, 0, 0], [0, 12, 6, 0, -12, 6], [0, 6, 4, 0, -6, 2],
                    [-1, 0, 0, 1, 0, 0], [0, -12, -6, 0, 12, -6], [0, 6, 2, 0, -6, 4]], dtype=float)
    assert np.array_equal(Kg[0].cpu().numpy(), exp)
    for case in ('rot_beam', 'vertical_beam', 'unit_beam'):
        m = util.golden_models()[case]
        g = util.golden(case)
        (xi, yi), (xj, yj) = m['nodes']
        b = m['beams'][0]
        Kg, Mg = batch.element_km(T([xi]), T([yi]), T([xj]), T([yj]), T([b['E']]), T([b['A']]), T([b['I']]),
                                  T([b['rho']]))
        assert util.relmax(Kg.cpu().numpy(), g['Ke_global']) < 1e-14
        assert util.relmax(Mg.cpu().numpy(), g['Me_global']) < 1e-14


@pytest.mark.parametrize('nE', [1, 31, 32, 33, 10007, 300000])
def test_k1_element_random(nE):
    rng = np.random.default_rng(nE)
    xi, yi = rng.uniform(-1e3, 1e3, nE), rng.uniform(-1e3, 1e3, nE)
    ang = rng.uniform(-math.pi, math.pi, nE)
    ang[: min(nE, 8)] = [0, math.pi / 2, -math.pi / 2, math.pi, math.atan(0.5), 1e-9, 3.0, -3.0][: min(nE, 8)]
    L = rng.uniform(0.5, 5e3, nE)
    xj, yj = xi + L * np.cos(ang), yi + L * np.sin(ang)
    E, A = rng.uniform(1e4, 3e5, nE), rng.uniform(1, 1e5, nE)
    I, rho = rng.uniform(1, 1e9, nE), rng.uniform(1e-9, 1e-8, nE)
    lumped = (rng.uniform(size=nE) < 0.25)
    Kg, Mg = batch.element_km(*(T(a) for a in (xi, yi, xj, yj, E, A, I, rho)), lumped=T(lumped, torch.uint8))
    coords = np.stack([np.stack([xi, yi], -1), np.stack([xj, yj], -1)], axis=1)      # [nE, 2, 2]
    props = np.stack([E, A, I, rho], -1)[:, None, :]
    Ko, Mo = orc.element_km_global(coords, np.array([[0, 1]]), props, lumped[:, None])
    Kg, Mg = Kg.cpu().numpy(), Mg.cpu().numpy()
    ek = np.abs(Kg - Ko[:, 0]).max(axis=(1, 2)) / np.abs(Ko[:, 0]).max(axis=(1, 2))
    em = np.abs(Mg - Mo[:, 0]).max(axis=(1, 2)) / np.abs(Mo[:, 0]).max(axis=(1, 2))
    assert ek.max() < 1e-14 and em.max() < 1e-14       # relative to each matrix's largest entry
    assert np.array_equal(Kg, np.swapaxes(Kg, 1, 2)) or util.relmax(Kg, np.swapaxes(Kg, 1, 2)) < 1e-15


# ------------------------------------------------------------------------------------------- K2
def test_k2_assembly_p20_vs_oracle():
    P, pk, plan = p20_batch(512)
    Kb, Mb = batch.assemble_banded(plan, pk['coords'], pk['props'], pk['nodal_mass'])
    o = oracle_p20(pk)
    Kd = plan.band_to_dense(Kb.cpu().numpy())
    Md = plan.band_to_dense(Mb.cpu().numpy())
    for b in range(512):
        assert util.relmax(np.tril(Kd[b]), np.tril(o['Kc'][b])) < 1e-14
        assert util.relmax(np.tril(Md[b]), np.tril(o['Mc'][b])) < 1e-14
    Kb2, none = batch.assemble_banded(plan, pk['coords'], pk['props'], None, want_mass=False)
    assert none is None and torch.equal(Kb, Kb2)


@pytest.mark.parametrize('case', ['chopra_102', 'chopra_103', 'chopra_105', 'allresults_1_mass', 'allresults_2_mass',
                                  'rod30_lumped', 'cantilever_example'])
def test_k2_assembly_golden(case):
    model = util.golden_models()[case]
    g = util.golden(case)
    pm = util.pack(model)
    plan = Plan.from_supports(len(model['nodes']), pm['conn'], model['supports'], pm['lumped'])
    Kb, Mb = batch.assemble_banded(plan, T(pm['coords'][None]), T(pm['props'][None]), T(pm['nodal_mass'][None]))
    Kc = orc.condense(g['K'], pm['fixed'])
    Mc = orc.condense(g['M_with_masses'], pm['fixed'])
    assert util.relmax(np.tril(plan.band_to_dense(Kb[0].cpu().numpy())), np.tril(Kc)) < 1e-14
    assert util.relmax(np.tril(plan.band_to_dense(Mb[0].cpu().numpy())), np.tril(Mc)) < 1e-14


# ------------------------------------------------------------------------------------------- K3
STATIC_CASES = ['unit_beam', 'rot_beam', 'vertical_beam', 'cantilever_example', 'chopra_105', 'allresults_1',
                'allresults_2', 'allresults_4', 'three_elements', 'clamped_clamped_loads',
                'p20_v0', 'p20_v1', 'p20_v2', 'p20_v3']


@pytest.mark.parametrize('case', STATIC_CASES)
def test_static_golden(case):
    """Displacements, local end forces and reactions against what the REAL reference computed."""
    model = util.golden_models()[case]
    g = util.golden(case)
    plan, pm, out = gpu_model(model, modal=False)
    assert out['info'] == 0
    keep = plan.pos_of_free
    fixed = plan.fixed
    assert (out['u'][:, fixed] == 0).all()             # exact zeros at supports (reference: -R/1e41)
    if len(keep):
        o = orc.solve_batch(pm['conn'], pm['fixed'], pm['coords'], pm['props'], f_nodal=pm['f_nodal'],
                            r_local=pm['r_local'])
        assert util.relmax(out['u'], o['u']) < 1e-10                      # vs numpy.linalg.solve
        assert util.relmax(out['u'][:, keep], g['u'][:, keep]) < 1e-9     # vs the reference's inv(K+1e41)
    assert util.relmax(out['endforces'], g['endforces']) < 1e-9
    assert util.relmax(out['reactions'], g['reactions']) < 1e-9


def test_static_unfused_equals_fused_and_oracle():
    P, pk, plan = p20_batch(300)
    Kb, _ = batch.assemble_banded(plan, pk['coords'], pk['props'], None, want_mass=False)
    s = batch.static_solve(plan, Kb, pk['coords'], pk['props'], pk['f_nodal'], pk['r_local'])
    f = batch.solve_fused(plan, pk['coords'], pk['props'], None, pk['f_nodal'], pk['r_local'], modal=False)
    # K3 is the warp-per-structure throughput kernel (warp-cooperative Cholesky), the fused kernel factorises in one
    # thread with rsqrt seeds: same algorithm, different rounding in the last bits
    assert torch.equal(s['info'], f['info']) and (s['info'] == 0).all()
    for k in ('u', 'endforces', 'reactions'):
        a, b = s[k].cpu().numpy(), f[k].cpu().numpy()
        assert (np.abs(a - b).max(axis=tuple(range(1, a.ndim))) / np.abs(b).max(axis=tuple(range(1, a.ndim)))).max() < 1e-10, k
    o = oracle_p20(pk)
    u = s['u'].cpu().numpy()
    for b in range(300):
        assert util.relmax(u[b], o['u'][b]) < 1e-10
        assert util.relmax(s['endforces'][b].cpu().numpy(), o['endforces'][b]) < 1e-10
        assert util.relmax(s['reactions'][b].cpu().numpy(), o['reactions'][b]) < 1e-10
    t, pk = util.p20_truth_batch(DEV)                  # 40-digit solve of the same K, f: 64 variants picked by condition
    s = batch.solve_fused(plan, pk['coords'], pk['props'], None, pk['f_nodal'], pk['r_local'], modal=False)
    assert util.relmax(s['u'].cpu().numpy(), t['u']) < 1e-10


# ------------------------------------------------------------------------------------------- K4
MODAL_CASES = ['unit_beam', 'cantilever_example', 'chopra_102', 'chopra_103', 'chopra_105', 'allresults_1',
               'allresults_1_mass', 'allresults_2', 'allresults_2_mass', 'three_elements', 'rod30_consistent',
               'p20_v0', 'p20_v1', 'p20_v2', 'p20_v3']


@pytest.mark.parametrize('case', MODAL_CASES)
def test_modal_golden(case):
    """Circular frequencies and mode shapes against the REAL reference (scipy.linalg.eigh inside)."""
    model = util.golden_models()[case]
    g = util.golden(case)
    plan, pm, out = gpu_model(model)
    assert out['info'] == 0 and 0 <= out['sweeps'] < 30          # 0: the two-ended engine (n_free <= 16 here asks <= 16 shapes)
    lam = out['lam']
    assert (np.diff(lam) >= 0).all()
    w = np.sqrt(lam)
    wr = g['circular_freq']
    bound = 1e-10 + 64 * EPS * (wr[-1] / wr) ** 2
    assert (np.abs(w / wr - 1) < bound).all()
    # shapes: zeros at supports, M-orthonormal, eigen-residual; per-mode match where the gap allows
    phi = out['phi']
    assert (phi[:, plan.fixed] == 0).all()
    Mc = orc.condense(g['M_with_masses'], pm['fixed'])
    Kc = orc.condense(g['K'], pm['fixed'])
    Pf = phi[:, plan.pos_of_free]
    assert np.abs(Pf @ Mc @ Pf.T - np.eye(len(lam))).max() < 1e-9
    res = Kc @ Pf.T - (Mc @ Pf.T) * lam[None, :]
    assert (np.abs(res).max(axis=0) / (np.abs(Kc).max() * np.abs(Pf).max(axis=1))).max() < 1e-11
    gap = np.minimum(np.diff(lam, prepend=-np.inf), np.diff(lam, append=np.inf))
    tol = 1e-10 + 256 * EPS * lam[-1] / gap
    assert (util.mode_err(phi, g['shapes']) < tol).all()


@pytest.fixture(params=['jacobi', 'two_ended'])
def engine(request):
    """both modal engines of the fused path (include/bfe2.h: BFE2_ENGINE_*) must meet the same bar"""
    prev = batch.set_modal_engine(request.param)
    yield request.param
    batch.set_modal_engine(prev)


def test_modal_p20_truth_1e10(engine):
    """The north_star bar, against a 40-digit solution of the same pencils: ALL 57 eigenvalues and the first
    10 M-normalised modes to 1e-10 relative, on 64 variants picked to span the sweep (lam_max / lam_min from 4.7e5 to
    6.4e7, near-degenerate pairs down to a relative gap of 4e-5; oracle/make_truth.py).  Modes of a pair closer than
    1e-3 are compared as a subspace."""
    t, pk = util.p20_truth_batch(DEV)
    conn, sup = sweep.p20_topology()
    plan = Plan.from_supports(21, conn, sup)
    out = batch.solve_fused(plan, pk['coords'], pk['props'], pk['nodal_mass'], pk['f_nodal'], pk['r_local'],
                            n_modes=10)
    assert (out['info'] == 0).all()
    lam = out['lam'].cpu().numpy()
    phi = out['phi'].cpu().numpy()
    assert lam.shape[0] >= 64
    assert np.abs(lam / t['lam'] - 1).max() < 1e-10
    n_pairs = 0
    for b in range(lam.shape[0]):
        e = util.mode_err_clustered(phi[b], t['phi'][b], t['lam'][b])
        n_pairs += int(np.sum(np.diff(t['lam'][b][:10]) / t['lam'][b][1:10] < 1e-3))
        assert np.nanmax(e) < 1e-10, (b, str(t['kind'][b]), e)
    assert n_pairs >= 16                                 # the close pairs really are in the fixture
    # sign convention: the largest-magnitude entry of every mode is positive
    big = np.take_along_axis(phi, np.abs(phi).argmax(axis=-1)[..., None], axis=-1)
    assert (big > 0).all()


@pytest.mark.parametrize('shard', [3, 5])
def test_modal_p20_vs_scipy_batch(shard, engine):
    """400 variants per shard: every one of the 57 eigenvalues within 1e-10 of scipy's eigenpairs refined by one
    extended-precision Rayleigh-quotient step (tests/util.py::refined_eigenvalues agrees with the 40-digit truth to
    1e-13, tests/test_oracle_pinning.py).  The DIRECT comparison with scipy's own eigenvalues is kept beside it with
    LAPACK's bound, because dsygvd itself is only absolutely accurate (up to 5e-10 off on the smallest ones)."""
    B = 400
    P, pk, plan = p20_batch(B, shard=shard)
    out = batch.solve_fused(plan, pk['coords'], pk['props'], pk['nodal_mass'], None, None, n_modes=10)
    assert (out['info'] == 0).all()
    sw = out['sweeps'].cpu().numpy()
    assert (sw.min() >= 5 and sw.max() <= 16) if engine == 'jacobi' else (sw == 0).all()   # no sweeps in the two-ended engine
    o = oracle_p20(pk, n_modes=57)
    lam = out['lam'].cpu().numpy()
    keep = plan.pos_of_free
    rq = util.refined_eigenvalues(o['Kc'], o['Mc'], o['phi'][:, :, keep])
    assert np.abs(lam / rq - 1).max() < 1e-10                                  # all 57 eigenvalues, no slack
    bound = 1e-10 + 64 * EPS * o['lam'][:, -1:] / o['lam']            # documented: scipy's own (LAPACK) error bound
    assert (np.abs(lam / o['lam'] - 1) < bound).all()
    phi = out['phi'].cpu().numpy()
    gap = np.minimum(np.diff(o['lam'], axis=1, prepend=-np.inf), np.diff(o['lam'], axis=1, append=np.inf))[:, :10]
    tol = 1e-10 + 256 * EPS * o['lam'][:, -1:] / gap
    assert (util.mode_err(phi, o['phi'][:, :10]) < tol).all()
    # eigenvectors without scipy in the loop: residual and M-orthonormality of the first 10 modes in extended precision
    K, M = o['Kc'].astype(np.longdouble), o['Mc'].astype(np.longdouble)
    V = phi[:, :, keep].astype(np.longdouble)
    R = np.einsum('bij,bmj->bmi', K, V) - lam[:, :10, None] * np.einsum('bij,bmj->bmi', M, V)
    scale = np.abs(o['Kc']).max(axis=(1, 2))[:, None] * np.abs(phi).max(axis=2)
    assert (np.abs(R).max(axis=2).astype(np.float64) / scale).max() < 1e-13
    G = np.einsum('bmi,bij,bnj->bmn', V, M, V).astype(np.float64)
    assert np.abs(G - np.eye(10)).max() < 1e-10


def test_modal_unfused_equals_fused():
    P, pk, plan = p20_batch(200)
    Kb, Mb = batch.assemble_banded(plan, pk['coords'], pk['props'], pk['nodal_mass'])
    m = batch.modal_solve(plan, Kb, Mb, n_modes=10)
    f = batch.solve_fused(plan, pk['coords'], pk['props'], pk['nodal_mass'], pk['f_nodal'], pk['r_local'], n_modes=10)
    # K2 forms the beam matrices in closed form, the fused kernel by the reference's two 6 x 6 products: the bands
    # differ in the last bit, the results within the rounding amplification of the eigenproblem
    assert torch.equal(m['info'], f['info'])
    lam_m, lam_f = m['lam'].cpu().numpy(), f['lam'].cpu().numpy()
    assert (np.abs(lam_m / lam_f - 1) < 1e-11 + 64 * EPS * lam_f[:, -1:] / lam_f).all()
    keep = plan.pos_of_free
    for b in range(0, 200, 40):                                         # mode shapes: sign-free, gap-scaled
        pm, pf = m['phi'][b].cpu().numpy()[:, keep], f['phi'][b].cpu().numpy()[:, keep]
        gap = np.minimum(np.diff(lam_f[b][:11]), np.r_[np.inf, np.diff(lam_f[b][:10])])
        tol = 1e-10 + 256 * EPS * lam_f[b, -1] / gap
        pm = pm * np.sign(np.sum(pm * pf, axis=1))[:, None]            # largest-entry sign rule can tie
        err = np.abs(pm - pf).max(axis=1) / np.abs(pf).max(axis=1)
        assert (err < tol).all(), (b, err, tol)


def test_rod30_analytic_and_lumped():
    """Tests/test_beam_modal.py:61-84 -- first five omegas within 1 % of [3.52, 22.03, 61.7, 121, 200] *
    sqrt(EI / rho A L^4), consistent and lumped mass (n_free = 90: the 4-lanes-per-pair Jacobi path)."""
    E, rho, L = 2.1e11, 7850, 0.45
    A, I = 0.003 * 0.02, 0.02 * 0.003 ** 3 / 12
    ana = np.array([3.52, 22.03, 61.7, 121, 200]) * math.sqrt(E * I / (rho * A * L ** 4))
    for case in ('rod30_consistent', 'rod30_lumped'):
        plan, pm, out = gpu_model(util.golden_models()[case], n_modes=12)
        assert plan.n_free == 90 and out['info'] == 0
        w = np.sqrt(out['lam'])
        bending = [x for x in w if x < 0.9 * math.sqrt(E / rho) * math.pi / (2 * L)][:5]   # below 1st axial
        assert np.abs(np.array(bending) / ana - 1).max() < 0.01, case


# ------------------------------------------------------------------------------------------- edges
def test_failure_reporting_and_edge_cases():
    conn, sup = sweep.p20_topology()
    P, pk, plan = p20_batch(8)
    # (a) K not positive definite -> info = 1-based column of the bad pivot, NaN outputs.  A free-free
    #     structure is singular only up to rounding, so the robust trigger is a negative modulus;
    #     for the free-free case flagged variants must be NaN and unflagged ones finite.
    neg = pk['props'].clone()
    neg[..., 0] = -neg[..., 0]
    out = batch.solve_fused(plan, pk['coords'], neg, pk['nodal_mass'], pk['f_nodal'], pk['r_local'], modal=False)
    assert (out['info'] == 1).all() and torch.isnan(out['u']).all() and torch.isnan(out['reactions']).all()
    out = batch.solve_fused(plan, pk['coords'], neg, pk['nodal_mass'], pk['f_nodal'], pk['r_local'], n_modes=3)
    assert (out['info'] == 1).all() and torch.isnan(out['u']).all() and torch.isnan(out['lam']).all()
    out = batch.solve_fused(plan, pk['coords'], neg, pk['nodal_mass'], None, None, n_modes=3)      # modal only
    assert (out['info'] == 1).all() and torch.isnan(out['lam']).all() and torch.isnan(out['phi']).all()
    free_plan = Plan(21, conn, [])
    out = batch.solve_fused(free_plan, pk['coords'], pk['props'], pk['nodal_mass'], pk['f_nodal'], pk['r_local'],
                            modal=False)
    bad = out['info'] > 0
    assert bad.sum() >= 6                        # the relative pivot test catches (almost) all of them
    assert torch.isnan(out['u'][bad]).all() and torch.isfinite(out['u'][~bad]).all()
    # (b) massless structure: M not positive definite -> info = -1 (reference returns False, Solver.py:34-35)
    props0 = pk['props'].clone()
    props0[..., 3] = 0.0
    out = batch.solve_fused(plan, pk['coords'], props0, None, None, None, n_modes=3)
    assert (out['info'] == -1).all() and torch.isnan(out['lam']).all()
    # (c) one bad variant does not disturb its neighbours
    props1 = pk['props'].clone()
    props1[3, :, 3] = 0.0
    out = batch.solve_fused(plan, pk['coords'], props1, None, pk['f_nodal'], pk['r_local'], n_modes=2)
    good = batch.solve_fused(plan, pk['coords'], pk['props'], None, pk['f_nodal'], pk['r_local'], n_modes=2)
    info = out['info'].cpu().numpy()
    assert info[3] == -1 and (np.delete(info, 3) == 0).all()
    keepers = [b for b in range(8) if b != 3]
    assert torch.equal(out['lam'][keepers], good['lam'][keepers])
    assert torch.equal(out['u'][keepers], good['u'][keepers])
    # the massless variant still has a valid STATIC result (K factored fine): only lam / phi are NaN
    assert torch.isnan(out['lam'][3]).all() and torch.isnan(out['phi'][3]).all()
    assert torch.equal(out['u'][3], good['u'][3]) and torch.equal(out['endforces'][3], good['endforces'][3])
    assert torch.equal(out['reactions'][3], good['reactions'][3])
    # (d) empty batch and wrong shapes
    e = batch.solve_fused(plan, pk['coords'][:0], pk['props'][:0], None, None, None, n_modes=0)
    assert e['lam'].shape == (0, 57)
    with pytest.raises(Exception):
        batch.solve_fused(plan, pk['coords'][:, :5], pk['props'], None, None, None)
    with pytest.raises(Exception):
        batch.solve_fused(plan, pk['coords'].cpu(), pk['props'].cpu(), None, None, None)


def test_large_batch_properties():
    """Config-3 size (1e5 variants): properties that need no CPU solve -- info, ordering, K u = f via the
    reactions at free DOFs (sum of end forces minus loads must vanish), M-orthonormality, eigen-residual."""
    B = 100000
    P, pk, plan = p20_batch(B)
    out = batch.solve_fused(plan, pk['coords'], pk['props'], pk['nodal_mass'], pk['f_nodal'], pk['r_local'], n_modes=10)
    torch.cuda.synchronize()
    assert (out['info'] == 0).all()
    lam = out['lam']
    assert (lam[:, 1:] >= lam[:, :-1]).all() and (lam > 0).all()
    free = torch.as_tensor(plan.pos_of_free.astype(np.int64), device=DEV)
    react = out['reactions']
    scale = react.abs().amax(dim=(1, 2), keepdim=True)
    assert (react[:, :, free].abs() / scale).max() < 1e-9            # equilibrium at every free DOF
    # global equilibrium: support reactions balance the applied loads (FX, FY per load case)
    conn, sup = sweep.p20_topology()
    Kb, Mb = batch.assemble_banded(plan, pk['coords'], pk['props'], pk['nodal_mass'])
    n, hb = plan.n_free, plan.hbw
    sub = slice(0, 4096)
    Md = torch.zeros((4096, n, n), dtype=torch.float64, device=DEV)
    Kd = torch.zeros_like(Md)
    for d in range(hb + 1):
        idx = torch.arange(n - d, device=DEV)
        Md[:, idx + d, idx] = Mb[sub, d, : n - d]
        Md[:, idx, idx + d] = Mb[sub, d, : n - d]
        Kd[:, idx + d, idx] = Kb[sub, d, : n - d]
        Kd[:, idx, idx + d] = Kb[sub, d, : n - d]
    Pf = out['phi'][sub][:, :, free]
    G = Pf @ Md @ Pf.transpose(1, 2)
    assert (G - torch.eye(10, dtype=torch.float64, device=DEV)).abs().max() < 1e-10
    R = Kd @ Pf.transpose(1, 2) - (Md @ Pf.transpose(1, 2)) * lam[sub, None, :10]
    rel = R.abs().amax(dim=1) / (Kd.abs().amax(dim=(1, 2))[:, None] * Pf.abs().amax(dim=2))
    assert rel.max() < 1e-12
    # trace identity: sum of all eigenvalues == trace(M^-1 K)
    tr = torch.linalg.solve(Md, Kd).diagonal(dim1=1, dim2=2).sum(dim=1)
    assert ((lam[sub].sum(dim=1) / tr) - 1).abs().max() < 1e-10


def test_degenerate_eigenvalues_subspace():
    """Two identical, disconnected cantilevers in one structure: every eigenvalue is exactly double.  Mode shapes
    of a cluster are only defined up to a rotation inside the eigenspace, so they are compared by projector."""
    nE = 6
    xy = np.array([(k * 100.0, 0.0) for k in range(nE + 1)] + [(k * 100.0, 500.0) for k in range(nE + 1)])
    conn = np.array([[k, k + 1] for k in range(nE)] + [[nE + 1 + k, nE + 2 + k] for k in range(nE)], dtype=np.int32)
    fixed = [0, 1, 2, 3 * (nE + 1), 3 * (nE + 1) + 1, 3 * (nE + 1) + 2]
    props = np.tile(np.array([2.1e5, 30.0, 22.5, 7.85e-9]), (2 * nE, 1))
    plan = Plan(2 * (nE + 1), conn, fixed)
    n = plan.n_free
    out = batch.solve_fused(plan, T(xy[None]), T(props[None]), None, None, None, n_modes=n)
    assert int(out['info'][0]) == 0
    lam = out['lam'][0].cpu().numpy()
    phi = out['phi'][0].cpu().numpy()
    o = orc.solve_batch(conn, fixed, xy, props, nodal_mass=np.zeros((2 * (nE + 1), 2)))
    assert np.abs(lam / o['lam'] - 1).max() < 1e-10 + 64 * EPS * o['lam'][-1] / o['lam'][0]
    assert np.abs(lam[0::2] / lam[1::2] - 1).max() < 1e-11          # exact pairs
    keep = plan.pos_of_free
    Pg, Ps, M = phi[:, keep], o['phi'][:, keep], o['Mc']
    assert np.abs(Pg @ M @ Pg.T - np.eye(n)).max() < 1e-9
    for k in range(0, 8, 2):                                           # projector onto each 2-dimensional eigenspace
        Pr_g = Pg[k:k + 2].T @ Pg[k:k + 2] @ M
        Pr_s = Ps[k:k + 2].T @ Ps[k:k + 2] @ M
        assert np.abs(Pr_g - Pr_s).max() < 1e-8


def test_maximum_size_of_the_jacobi_path():
    """n_free = 116 is the largest generalized eigenproblem the shared-memory Jacobi accepts; 117 is refused loudly
    (the static path has no such limit)."""
    nE = 39
    xy = np.stack([np.arange(nE + 1) * 50.0, np.zeros(nE + 1)], axis=1)
    conn = np.stack([np.arange(nE), np.arange(1, nE + 1)], axis=1).astype(np.int32)
    props = np.tile(np.array([2.1e5, 30.0, 22.5, 7.85e-9]), (nE, 1))
    f = np.zeros((1, 1, 3 * (nE + 1)))
    f[0, 0, -2] = -10.0
    for fixed, ok in (([0, 1, 2, 3 * nE + 1], True), ([0, 1, 2], False)):
        plan = Plan(nE + 1, conn, fixed)
        assert plan.n_free == (116 if ok else 117)
        if ok:
            out = batch.solve_fused(plan, T(xy[None]), T(props[None]), None, T(f), None, n_modes=5)
            assert int(out['info'][0]) == 0
            o = orc.solve_batch(conn, fixed, xy, props, np.zeros((nE + 1, 2)), f[0], None)
            lam = out['lam'][0].cpu().numpy()
            assert (np.abs(lam / o['lam'] - 1) < 1e-10 + 64 * EPS * o['lam'][-1] / o['lam']).all()
            assert util.relmax(out['u'][0].cpu().numpy(), o['u']) < 1e-9
        else:
            with pytest.raises(Exception, match='Jacobi'):
                batch.solve_fused(plan, T(xy[None]), T(props[None]), None, T(f), None, n_modes=5)
            st = batch.solve_fused(plan, T(xy[None]), T(props[None]), None, T(f), None, modal=False)
            assert int(st['info'][0]) == 0


@pytest.mark.parametrize('n_nodes,fixed', [(20, [0, 1, 2, 58]), (21, [0, 1, 2]), (22, [0, 1, 2]), (23, [0, 1, 2, 66, 67]),
                                            (23, [0, 1, 2])])
def test_register_path_sizes_59_to_64_and_first_smem_size(n_nodes, fixed):
    """n_free = 56, 60, 63, 64 run the register-resident Jacobi with 32 rows per lane (64 = no dummy column, all 32
    lanes active); 66 is the first size on the shared-memory variant.  Same checks on both sides of the boundary."""
    rng = np.random.default_rng(n_nodes)
    B = 6
    xy = np.stack([np.arange(n_nodes) * 400.0, 150.0 * np.cos(np.arange(n_nodes))], axis=1)
    conn = np.stack([np.arange(n_nodes - 1), np.arange(1, n_nodes)], axis=1).astype(np.int32)
    coords = xy[None] + rng.uniform(-20, 20, (B, n_nodes, 2))
    props = np.tile(np.array([2.1e5, 30.0, 22.5, 7.85e-9]), (B, n_nodes - 1, 1)) * rng.uniform(0.8, 1.2, (B, n_nodes - 1, 1))
    mass = rng.uniform(0, 1e-4, (B, n_nodes, 2))
    f = np.zeros((B, 2, 3 * n_nodes))
    f[:, 0, -2] = -10.0
    f[:, 1, 3 * (n_nodes // 2)] = 5.0
    plan = Plan(n_nodes, conn, fixed)
    out = batch.solve_fused(plan, T(coords), T(props), T(mass), T(f), None, n_modes=plan.n_free)
    assert (out['info'] == 0).all()
    o = orc.solve_batch(conn, fixed, coords, props, mass, f, None)
    lam = out['lam'].cpu().numpy()
    assert (np.abs(lam / o['lam'] - 1) < 1e-10 + 64 * EPS * o['lam'][:, -1:] / o['lam']).all()
    condK = np.linalg.cond(o['Kc']).max()
    assert util.relmax(out['u'].cpu().numpy(), o['u']) < max(1e-10, 50 * EPS * condK)
    keep = plan.pos_of_free
    phi = out['phi'].cpu().numpy()
    for b in range(B):
        Pf = phi[b][:, keep]
        assert np.abs(Pf @ o['Mc'][b] @ Pf.T - np.eye(plan.n_free)).max() < 1e-9
        res = o['Kc'][b] @ Pf.T - (o['Mc'][b] @ Pf.T) * lam[b][None, :]
        assert (np.abs(res).max(axis=0) / (np.abs(o['Kc'][b]).max() * np.abs(Pf).max(axis=1))).max() < 1e-11


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_one_plan_on_two_devices():
    """A plan keeps one device copy per GPU: alternating devices must neither re-upload nor free a copy in use."""
    conn, sup = sweep.p20_topology()
    plan = Plan.from_supports(21, conn, sup)
    P = sweep.p20_draw_params(64)
    outs = []
    for rep in range(3):
        for d in (0, 1):
            pk = sweep.p20_pack(torch.from_numpy(P).to('cuda:%d' % d))
            outs.append(batch.solve_fused(plan, pk['coords'], pk['props'], pk['nodal_mass'], pk['f_nodal'],
                                          pk['r_local'], n_modes=4))
    for d in (0, 1):
        torch.cuda.synchronize(d)
    ref = outs[0]
    for o in outs[1:]:
        assert torch.equal(o['lam'].cpu(), ref['lam'].cpu()) and torch.equal(o['u'].cpu(), ref['u'].cpu())


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_one_plan_from_two_host_threads():
    """One plan, one host thread per GPU, concurrent launches (ctypes drops the GIL): every call must get the plan
    copy of ITS device -- the plan keeps no 'current device' state."""
    import threading
    conn, sup = sweep.p20_topology()
    plan = Plan.from_supports(21, conn, sup)                    # first upload happens inside the threads
    P = sweep.p20_draw_params(256)
    results, errors = {}, []

    def work(d):
        try:
            with torch.cuda.device(d):
                pk = sweep.p20_pack(torch.from_numpy(P).to('cuda:%d' % d))
                outs = [batch.solve_fused(plan, pk['coords'], pk['props'], pk['nodal_mass'], pk['f_nodal'],
                                          pk['r_local'], n_modes=4) for _ in range(20)]
                torch.cuda.synchronize(d)
                results[d] = [(o['lam'].cpu(), o['u'].cpu(), o['info'].cpu()) for o in outs]
        except Exception as e:                                      # noqa: BLE001
            errors.append((d, repr(e)))

    threads = [threading.Thread(target=work, args=(d,)) for d in (0, 1)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    ref = results[0][0]
    for d in (0, 1):
        for lam, u, info in results[d]:
            assert int(info.abs().sum()) == 0
            assert torch.equal(lam, ref[0]) and torch.equal(u, ref[1])


def _compare_packed(io, views, ref, plan, modal, exact):
    """compact results of the packed path against the dense tensors of solve_fused"""
    pos = torch.as_tensor(plan.pos_of_free.astype('int64'), device=ref['info'].device)
    fixed = torch.as_tensor(plan.fixed.astype('int64'), device=ref['info'].device)

    def same(a, b, name):
        if exact:
            assert torch.equal(a, b), name
        else:
            assert ((a - b).abs().max() / b.abs().max()).item() < 1e-12, name
    same(views['u'], ref['u'][:, :, pos], 'u')
    same(views['endforces'], ref['endforces'], 'endforces')
    same(views['reactions'], ref['reactions'][:, :, fixed], 'reactions')
    assert torch.equal(views['info'], ref['info'])
    if modal:
        same(views['lam'], ref['lam'], 'lam')
        same(views['phi'], ref['phi'][:, :, pos], 'phi')
        assert torch.equal(views['sweeps'], ref['sweeps'])
    dense = io.expand(views)
    same(dense['u'], ref['u'], 'u dense')
    assert (dense['u'][:, :, fixed] == 0).all()
    if modal:
        same(dense['phi'], ref['phi'], 'phi dense')


@pytest.mark.parametrize('modal,compact_loads', [(True, True), (True, False), (False, True)])
def test_packed_records_equal_the_dense_call(modal, compact_loads):
    """bfe2_solve_fused_packed (one record per structure, non-zero load entries only, u / phi on the free DOFs,
    reactions at the supports) returns bit for bit what bfe2_solve_fused returns on the dense tensors.  Static-only
    runs go through the block-per-structure kernel instead of the 16-lane group kernel: same arithmetic up to the
    order of the substitutions, compared to 1e-12."""
    P, pk, plan = p20_batch(777)
    if compact_loads:
        io = batch.PackedIO.from_dense(plan, pk['f_nodal'], pk['r_local'], n_modes=6, modal=modal)
        assert io.f_index.tolist() == [21, 2 * 63 + 21, 2 * 63 + 41]            # FX node 8 (LC1, LC3), MZ node 14 (LC3)
        assert io.r_index.size == 2 * 6 * 4                                     # 6 girder beams x (V, M) x 2 ends, LC2 + LC3
        assert io.w_in == 42 + 80 + 42 + 3 + 48
        assert io.w_out == 3 * 57 + 360 + 18 + (57 + 6 * 57 if modal else 0) + 1
    else:
        io = batch.PackedIO(plan, 3, 6, modal)
        assert io.w_in == 42 + 80 + 42 + 189 + 360
    rec = torch.zeros((777, io.w_in), dtype=torch.float64, device=DEV)
    io.pack_inputs(rec, pk['coords'], pk['props'], pk['nodal_mass'], pk['f_nodal'], pk['r_local'])
    out = io.solve(rec)
    ref = batch.solve_fused(plan, pk['coords'], pk['props'], pk['nodal_mass'], pk['f_nodal'], pk['r_local'],
                            modal=modal, n_modes=6)
    assert (ref['info'] == 0).all()
    _compare_packed(io, io.out_views(out), ref, plan, modal, exact=modal)


def test_packed_records_failure_codes_and_nan_fill():
    P, pk, plan = p20_batch(8)
    io = batch.PackedIO.from_dense(plan, pk['f_nodal'], pk['r_local'], n_modes=4)
    props = pk['props'].clone()
    props[2, :, 0] = -props[2, :, 0]               # K not positive definite: everything NaN
    props[5, :, 3] = 0.0                           # massless, no nodal mass below: modal NaN, static valid
    rec = torch.zeros((8, io.w_in), dtype=torch.float64, device=DEV)
    io.pack_inputs(rec, pk['coords'], props, None, pk['f_nodal'], pk['r_local'])
    v = io.out_views(io.solve(rec))
    info = v['info'].cpu().numpy()
    assert info[2] == 1 and info[5] == -1 and (np.delete(info, [2, 5]) == 0).all()
    assert torch.isnan(v['u'][2]).all() and torch.isnan(v['reactions'][2]).all() and torch.isnan(v['lam'][2]).all()
    assert torch.isnan(v['lam'][5]).all() and torch.isnan(v['phi'][5]).all() and torch.isfinite(v['u'][5]).all()
    ok = [b for b in range(8) if b not in (2, 5)]
    assert torch.isfinite(v['u'][ok]).all() and torch.isfinite(v['phi'][ok]).all()


@pytest.mark.parametrize('modal', [True, False])
def test_host_pipeline_equals_the_device_call(modal):
    """The end-to-end path (pinned host records, chunks over three streams, ragged last chunk, staging buffers
    reused, ONE copy per direction per chunk) returns exactly what one device-resident call returns."""
    P, pk, plan = p20_batch(1000)
    io = batch.PackedIO.from_dense(plan, pk['f_nodal'], pk['r_local'], n_modes=6, modal=modal)
    pipe = batch.HostPipeline(plan, n_loadcases=3, n_modes=6, chunk=333, modal=modal, io=io)
    h_in, h_out = pipe.alloc_host(1000)
    io.pack_inputs(h_in, pk['coords'].cpu(), pk['props'].cpu(), pk['nodal_mass'].cpu(), pk['f_nodal'].cpu(),
                   pk['r_local'].cpu())
    pipe.run(h_in, h_out)
    assert pipe.launches == 4 and pipe.copies == 8
    ref = batch.solve_fused(plan, pk['coords'], pk['props'], pk['nodal_mass'], pk['f_nodal'], pk['r_local'],
                            modal=modal, n_modes=6)
    ref = {k: (v.cpu() if v is not None else None) for k, v in ref.items()}
    _compare_packed(io, io.out_views(h_out), ref, plan, modal, exact=modal)
    bi, bo = pipe.bytes_per_structure()
    assert bi == 8 * (42 + 80 + 42 + 3 + 48)
    assert bo == 8 * (171 + 360 + 18 + (57 + 342 if modal else 0) + 1)


def test_two_ended_engine_vs_jacobi_and_repair_path():
    """The two engines of the fused path agree to the accuracy both have against the truth (a few 1e-11), on the dense
    and on the packed entry; a structure the two-ended engine declines (it cannot: n_modes > 16 is routed to Jacobi
    up front) and the static results are engine-independent bit for bit."""
    P, pk, plan = p20_batch(2000, shard=9)
    res = {}
    for eng in ('two_ended', 'jacobi'):
        prev = batch.set_modal_engine(eng)
        try:
            res[eng] = batch.solve_fused(plan, pk['coords'], pk['props'], pk['nodal_mass'], pk['f_nodal'], pk['r_local'],
                                         n_modes=10)
            io = batch.PackedIO.from_dense(plan, pk['f_nodal'], pk['r_local'], n_modes=10)
            rec = torch.zeros((2000, io.w_in), dtype=torch.float64, device=DEV)
            io.pack_inputs(rec, pk['coords'], pk['props'], pk['nodal_mass'], pk['f_nodal'], pk['r_local'])
            res[eng + '_packed'] = io.out_views(io.solve(rec))
        finally:
            batch.set_modal_engine(prev)
    a, j = res['two_ended'], res['jacobi']
    assert (a['info'] == 0).all() and (j['info'] == 0).all()
    assert (a['sweeps'] == 0).all() and (j['sweeps'] > 0).all()
    la, lj = a['lam'].cpu().numpy(), j['lam'].cpu().numpy()
    assert np.abs(la / lj - 1).max() < 2e-10                    # each engine is within 1e-10 of the truth
    pa, pj = a['phi'].cpu().numpy(), j['phi'].cpu().numpy()
    for b in range(0, 2000, 7):
        e = util.mode_err_clustered(pa[b], pj[b], lj[b])
        assert np.nanmax(e) < 1e-10, (b, e)
    # static part: same factor of K up to the order of the elimination: 1e-12
    assert ((a['u'] - j['u']).abs().max() / j['u'].abs().max()).item() < 1e-12
    assert ((a['endforces'] - j['endforces']).abs().max() / j['endforces'].abs().max()).item() < 1e-12
    # packed entry = dense entry, bit for bit, for either engine
    pos = torch.as_tensor(plan.pos_of_free.astype('int64'), device=DEV)
    for eng in ('two_ended', 'jacobi'):
        assert torch.equal(res[eng + '_packed']['lam'], res[eng]['lam'])
        assert torch.equal(res[eng + '_packed']['phi'], res[eng]['phi'][:, :, pos])
        assert torch.equal(res[eng + '_packed']['u'], res[eng]['u'][:, :, pos])
    # more shapes than the two-ended engine returns: routed to Jacobi whatever the setting
    many = batch.solve_fused(plan, pk['coords'][:16], pk['props'][:16], pk['nodal_mass'][:16], None, None, n_modes=30)
    assert (many['sweeps'] > 0).all() and (many['info'] == 0).all()


@pytest.mark.parametrize('packed', [False, True])
def test_change_of_units_is_only_a_change_of_units(packed):
    """Dimensional consistency (no hidden absolute threshold anywhere on the path): the P20 batch in N-mm-t units and the
    same structures in SI (m, Pa, kg/m^3, N) must give the same frequencies, and displacements / end forces / reactions
    that differ by the unit factors only -- to rounding, although every intermediate number differs by 3 to 12 decades."""
    _, pk, plan = p20_batch(256)
    si = dict(coords=pk['coords'] * 1e-3,
              props=pk['props'] * torch.tensor([1e6, 1e-6, 1e-12, 1e12], dtype=torch.float64, device=DEV),
              nodal_mass=pk['nodal_mass'] * 1e3,                       # t -> kg
              f_nodal=pk['f_nodal'].clone(), r_local=pk['r_local'].clone())
    si['f_nodal'][..., 2::3] *= 1e-3                                   # moments N mm -> N m
    si['r_local'][..., 2::3] *= 1e-3
    if packed:
        def solve(d):
            io = batch.PackedIO.from_dense(plan, d['f_nodal'], d['r_local'], n_modes=10)
            rec = torch.zeros((d['coords'].shape[0], io.w_in), dtype=torch.float64, device=DEV)
            io.pack_inputs(rec, d['coords'], d['props'], d['nodal_mass'], d['f_nodal'], d['r_local'])
            return io.expand(io.out_views(io.solve(rec)))
    else:
        def solve(d):
            return batch.solve_fused(plan, d['coords'], d['props'], d['nodal_mass'], d['f_nodal'], d['r_local'], n_modes=10)
    a, b = solve(pk), solve(si)
    assert int(a['info'].abs().max()) == 0 and int(b['info'].abs().max()) == 0
    assert torch.equal(a['sweeps'], b['sweeps']) or (a['sweeps'] - b['sweeps']).abs().max() <= 1
    assert float((b['lam'] / a['lam'] - 1).abs().max()) < 1e-10
    uf = torch.tensor([1e-3, 1e-3, 1.0], dtype=torch.float64, device=DEV).repeat(21)          # ux, uy [m], rotz [rad]
    ff = torch.tensor([1.0, 1.0, 1e-3], dtype=torch.float64, device=DEV)
    for key, fac in (('u', uf), ('reactions', ff.repeat(21)), ('endforces', ff.repeat(2))):
        x, y = a[key] * fac, b[key]
        scale = (x.abs() / fac).amax(dim=-1, keepdim=True) * fac       # per structure and load case, per unit class
        assert float(((x - y).abs() / scale.clamp_min(1e-300)).max()) < 1e-9, key
    # M-normalised shapes: phi = u / q with q in sqrt(mass) x length -> translations x sqrt(t / kg), rotations x 1e3 more
    pf = torch.tensor([1.0, 1.0, 1e3], dtype=torch.float64, device=DEV).repeat(21) * np.sqrt(1e-3)
    phi_a, phi_b = a['phi'] * pf, b['phi']
    sgn = torch.sign((phi_a * phi_b).sum(dim=-1, keepdim=True))
    gap_ok = ((a['lam'][:, 1:11] - a['lam'][:, :10]) / a['lam'][:, 1:11] > 1e-3) & \
        torch.cat([torch.ones_like(a['lam'][:, :1], dtype=torch.bool), (a['lam'][:, 1:10] - a['lam'][:, :9]) / a['lam'][:, 1:10] > 1e-3], 1)
    err = (phi_a * sgn - phi_b).abs().amax(dim=-1) / phi_b.abs().amax(dim=-1)
    assert float(err[gap_ok].max()) < 1e-7


@pytest.mark.parametrize('modal', [False, True])
def test_superposition_of_load_cases_at_full_size(modal):
    """Linearity at config-3 size (1e5 variants, no CPU solve): with load case 3 := load case 1 + load case 2 (nodal
    loads and beam loads), displacements, end forces and reactions of case 3 are the sums -- on the static-only
    kernel and on the fused static + modal kernel."""
    B = 100000
    _, pk, plan = p20_batch(B)
    f, r = pk['f_nodal'].clone(), pk['r_local'].clone()
    f[:, 2] = f[:, 0] + f[:, 1]
    r[:, 2] = r[:, 0] + r[:, 1]
    out = batch.solve_fused(plan, pk['coords'], pk['props'], pk['nodal_mass'] if modal else None, f, r, modal=modal,
                            n_modes=3 if modal else None)
    assert (out['info'] == 0).all()
    for key in ('u', 'endforces', 'reactions'):
        x = out[key].reshape(B, 3, -1)
        scale = x.abs().amax(dim=(1, 2))
        assert float(((x[:, 0] + x[:, 1] - x[:, 2]).abs().amax(dim=1) / scale).max()) < 1e-11, key


def test_every_entry_point_accepts_an_empty_batch():
    """B = 0 (an empty shard of a ragged split, tests/test_dist_gloo.py) is legal everywhere: outputs of the right
    shape, no launch, no error."""
    _, pk, plan = p20_batch(4)
    z = {k: v[:0] for k, v in pk.items()}
    Kb, Mb = batch.assemble_banded(plan, z['coords'], z['props'], z['nodal_mass'])
    assert Kb.shape == (0, plan.hbw + 1, plan.n_free) and Mb.shape == Kb.shape
    st = batch.static_solve(plan, Kb, z['coords'], z['props'], z['f_nodal'], z['r_local'])
    assert st['u'].shape == (0, 3, 63) and st['endforces'].shape == (0, 3, 20, 6) and st['info'].shape == (0,)
    mo = batch.modal_solve(plan, Kb, Mb, n_modes=10)
    assert mo['lam'].shape == (0, 57) and mo['phi'].shape == (0, 10, 63)
    for modal in (False, True):
        out = batch.solve_fused(plan, z['coords'], z['props'], z['nodal_mass'], z['f_nodal'], z['r_local'], modal=modal,
                                n_modes=10 if modal else None)
        assert out['u'].shape == (0, 3, 63) and out['reactions'].shape == (0, 3, 63)
    io = batch.PackedIO.from_dense(plan, pk['f_nodal'], pk['r_local'], n_modes=10)
    rec = torch.zeros((0, io.w_in), dtype=torch.float64, device=DEV)
    v = io.out_views(io.solve(rec))
    assert v['lam'].shape == (0, 57) and v['info'].shape == (0,)
    e = torch.zeros(0, dtype=torch.float64, device=DEV)
    Kg, Mg = batch.element_km(e, e, e, e, e, e, e, e)
    assert Kg.shape == (0, 6, 6) and Mg.shape == (0, 6, 6)
    act, mx = batch.internal_actions(plan, z['coords'], st['endforces'], None, None, npts=5)
    assert act.shape[0] == 0 and mx.shape[0] == 0
    torch.cuda.synchronize()
