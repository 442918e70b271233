"""Config 4 (one large chain structure): partitioned block-tridiagonal static solve, DMMA block products,
dense Rayleigh-Ritz eigen-solve and shift-invert subspace iteration, against scipy's banded / sparse
solvers and the Euler-Bernoulli closed forms.  Parity criterion per SURVEY.md 7.3-2: residuals and
closed-form error, because cond(K) ~ n^4 makes digit-for-digit agreement between two correct FP64
solvers impossible at large n (the tolerances below scale with the measured condition)."""
import math

import numpy as np
import pytest
import scipy.linalg as sla
import torch

from beamfe2_b200 import sweep
from beamfe2_b200.beam import LargeBeam, Q
from oracle import beam_oracle as bo
from tests import util

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def make(n_elem, length=1000.0, n_partitions=0, precision='dd'):
    pk = sweep.beam_pack(n_elem, length=length, device=DEV)
    beam = LargeBeam(pk['coords'][0], pk['props'][0], [0, 1, 2], n_partitions=n_partitions, precision=precision)
    return pk, beam


@pytest.mark.parametrize('n_elem,parts', [(3, 0), (10, 1), (10, 3), (257, 0), (257, 5), (1000, 0), (1000, 17), (4096, 64),
                                          (30, -8), (257, -40), (1000, -200), (4096, 0), (4500, -1400)])   # < 0, and 0 from 4096 nodes on: two levels
def test_static_fp64_kernels_vs_banded_cholesky(n_elem, parts):
    """The plain-FP64 instantiation of the chain kernels: as accurate as scipy's banded Cholesky, i.e. within the
    conditioning of the FP64-assembled matrix (kept for comparison; the default is double-double, below)."""
    length = 1000.0
    pk, beam = make(n_elem, length, parts, precision='fp64')
    f = pk['f_nodal'][0, 0].clone()
    un = beam.static(f).cpu().numpy()
    assert (un[:3] == 0).all()
    coords, props = pk['coords'][0].cpu().numpy(), pk['props'][0].cpu().numpy()
    K, M, keep = bo.chain_banded(coords, props, [0, 1, 2])
    r = K @ un[keep] - f.cpu().numpy()[keep]
    assert np.abs(r).max() / (abs(K).max() * np.abs(un).max()) < 1e-14
    us = bo.static(coords, props, [0, 1, 2], f.cpu().numpy())
    cond_est = 16 * n_elem ** 4                      # ~ cond(K) of the cantilever
    assert util.relmax(un, us) < max(1e-11, 1e-15 * cond_est)


@pytest.mark.parametrize('n_elem,parts', [(3, 0), (10, 3), (257, 5), (1000, 17), (4096, 64), (10000, 0), (100000, 0),
                                          (100000, 500), (257, -40), (10000, 100), (100000, -4000)])
def test_static_double_double_vs_closed_form_and_scipy(n_elem, parts):
    """SURVEY 7.3-2 at the sizes BASELINE.json names (config 4: n = 1e5): (i) scaled residual no worse than
    scipy.linalg.solveh_banded's, (ii) error against the closed form PL^3/3EI, PL^2/2EI (nodal values of Hermitian
    elements are exact) no worse than scipy's -- in fact <= 1e-10 at every size, while scipy loses eps * cond."""
    length = 1000.0
    pk, beam = make(n_elem, length, parts)
    f = pk['f_nodal'][0, 0].clone()
    u, ulo = beam.static(f, with_lo=True)
    un = u.cpu().numpy()
    assert (un[:3] == 0).all()
    coords, props = pk['coords'][0].cpu().numpy(), pk['props'][0].cpu().numpy()
    fn = f.cpu().numpy()
    K, M, keep = bo.chain_banded(coords, props, [0, 1, 2])
    us = bo.static(coords, props, [0, 1, 2], fn)
    knorm = abs(K).max()
    # the stored double-double K, rounded, is the FP64 K of the oracle to the last bits
    blk = beam.blocks().cpu().numpy()
    for k in range(1, min(3, n_elem)):                  # free numbering starts at node 1 (node 0 is clamped)
        ref = K[3 * (k - 1):3 * k, 3 * (k - 1):3 * k].toarray()
        assert np.abs(blk[k, 0] - ref).max() <= 4e-16 * knorm
    # (i) scaled residuals, both evaluated the same way in FP64 with the oracle's FP64 K
    res_gpu = np.abs(K @ un[keep] - fn[keep]).max() / (knorm * np.abs(un).max())
    res_cpu = np.abs(K @ us[keep] - fn[keep]).max() / (knorm * np.abs(us).max())
    assert res_gpu < 1e-15
    # ... and the residual of the full double-double solution evaluated in double-double on the device
    r_dd = beam.residual(f, u, ulo)
    res_dd = float(r_dd.abs().max().item()) / (knorm * np.abs(un).max())
    assert res_dd < 1e-28 and res_dd < res_cpu
    # (ii) closed form
    E, A, I = props[0, 0], props[0, 1], props[0, 2]
    ltot = float(coords[-1, 0] - coords[0, 0])
    d, rot, _ = bo.cantilever_closed_form(100.0, ltot, E, I, props[0, 3], A, 1)
    err_gpu, err_cpu = abs(un[-2] / d - 1), abs(us[-2] / d - 1)
    rot_gpu, rot_cpu = abs(un[-1] / rot - 1), abs(us[-1] / rot - 1)
    assert err_gpu < 1e-10 and rot_gpu < 1e-10
    assert err_gpu <= max(err_cpu, 1e-14) and rot_gpu <= max(rot_cpu, 1e-14)
    # the whole deflection line against the closed form u(x) = P x^2 (3L - x) / 6EI
    x = coords[:, 0] - coords[0, 0]
    line = 100.0 * x ** 2 * (3 * ltot - x) / (6 * E * I)
    assert np.abs(un[1::3] - line).max() / line.max() < 1e-10
    # (iii) where FP64 can still represent the problem the two routes agree within scipy's conditioning
    if n_elem <= 4096:
        assert util.relmax(un, us) < max(1e-11, 1e-15 * 16 * n_elem ** 4)


def test_matvec_and_multi_rhs_solve():
    pk, beam = make(777, 500.0, 9)
    coords, props = pk['coords'][0].cpu().numpy(), pk['props'][0].cpu().numpy()
    K, M, keep = bo.chain_banded(coords, props, [0, 1, 2])
    g = torch.Generator().manual_seed(1)
    X = torch.randn((beam.n, 5), generator=g, dtype=torch.float64).to(DEV)
    X[:3] = 0
    Xn = X.cpu().numpy()
    KX = beam.matvec(X).cpu().numpy()
    MX = beam.matvec(X, mass=True).cpu().numpy()
    assert util.relmax(KX[keep], K @ Xn[keep]) < 1e-14
    assert util.relmax(MX[keep], M @ Xn[keep]) < 1e-14
    assert np.array_equal(KX[:3], Xn[:3]) and (MX[:3] == 0).all()      # supports: K_ii = 1, M_ii = 0
    F = torch.as_tensor(KX, device=DEV)
    U = beam.solve(F).cpu().numpy()
    assert util.relmax(U, Xn) < 1e-15 * 16 * 777 ** 4                   # K^-1 (K X) = X up to cond * eps
    R = beam.matvec(torch.as_tensor(U, device=DEV)).cpu().numpy() - KX
    assert np.abs(R).max() / (abs(K).max() * np.abs(U).max()) < 1e-14


def test_dmma_gram_and_update():
    pk, beam = make(1500)
    g = torch.Generator().manual_seed(2)
    X = torch.randn((beam.n, Q), generator=g, dtype=torch.float64).to(DEV)
    Y = torch.randn((beam.n, Q), generator=g, dtype=torch.float64).to(DEV)
    Qm = torch.randn((Q, Q), generator=g, dtype=torch.float64).to(DEV)
    C = beam.gram(X, Y)
    ref = X.t() @ Y
    assert ((C - ref).abs().max() / ref.abs().max()).item() < 1e-13
    Z = beam.update(X, Qm)
    ref = X @ Qm
    assert ((Z - ref).abs().max() / ref.abs().max()).item() < 1e-13
    assert torch.equal(beam.gram(X, Y), C)                              # deterministic reduction order


def test_small_dense_eig_vs_scipy():
    pk, beam = make(64)
    rng = np.random.default_rng(5)
    R1, R2 = rng.standard_normal((Q, Q)), rng.standard_normal((Q, Q))
    A = R1 @ R1.T + Q * np.diag(np.logspace(0, 6, Q))
    B = R2 @ R2.T + Q * np.eye(Q)
    lam, V, info = beam.small_eig(torch.as_tensor(A, device=DEV), torch.as_tensor(B, device=DEV))
    assert int(info.item()) == 0
    w, v = sla.eigh(A, B)
    lam, V = lam.cpu().numpy(), V.cpu().numpy()
    assert (np.abs(lam / w - 1) < 1e-11 + 64 * np.finfo(float).eps * w[-1] / w).all()      # scipy's own bound
    assert np.abs(V @ B @ V.T - np.eye(Q)).max() < 1e-11
    assert np.abs(A @ V.T - B @ V.T * lam[None, :]).max() / np.abs(A).max() < 1e-11


@pytest.mark.parametrize('n', [7, 24, 40, 58])
def test_dense_eig_register_path(n):
    """bfe2_sym_eig_dense for n <= 58: the register-resident Jacobi path with the pivoted-Cholesky preconditioner
    on a full (not banded) pencil; a batch of graded matrices, one of them with an indefinite A."""
    import ctypes as C
    from beamfe2_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(n)
    nb = 4
    A = np.empty((nb, n, n))
    Bm = np.empty((nb, n, n))
    for b in range(nb):
        R1, R2 = rng.standard_normal((n, n)), rng.standard_normal((n, n))
        D = np.diag(np.logspace(0, 2 * b, n))                           # graded: cond up to 1e6 * cond(R1 R1^T)
        A[b] = D @ (R1 @ R1.T + n * np.eye(n)) @ D
        Bm[b] = R2 @ R2.T + n * np.eye(n)
    A[3] = -A[3]                                                        # not positive definite
    h = C.c_void_p()
    _lib.check(L.bfe2_plan_create_dense(C.byref(h), n))
    tA, tB = torch.as_tensor(A, device=DEV), torch.as_tensor(Bm, device=DEV)
    lam = torch.empty((nb, n), dtype=torch.float64, device=DEV)
    V = torch.empty((nb, n, n), dtype=torch.float64, device=DEV)
    info = torch.zeros(nb, dtype=torch.int32, device=DEV)
    sw = torch.zeros(nb, dtype=torch.int32, device=DEV)
    ws = torch.empty(nb * 2 * n * n, dtype=torch.float64, device=DEV)
    _lib.check(L.bfe2_sym_eig_dense(h, nb, _lib.ptr(tA), _lib.ptr(tB), _lib.ptr(lam), _lib.ptr(V), _lib.ptr(info),
                                    _lib.ptr(sw), _lib.ptr(ws), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    L.bfe2_plan_destroy(h)
    info = info.cpu().numpy()
    assert (info[:3] == 0).all() and info[3] > 0 and torch.isnan(lam[3]).all()
    assert (sw[:3].cpu().numpy() <= 12).all()
    for b in range(3):
        w = sla.eigh(A[b], Bm[b], eigvals_only=True)
        l, v = lam[b].cpu().numpy(), V[b].cpu().numpy()
        assert (np.abs(l / w - 1) < 1e-11 + 64 * np.finfo(float).eps * w[-1] / w).all()
        assert np.abs(v @ Bm[b] @ v.T - np.eye(n)).max() < 1e-10
        assert np.abs(A[b] @ v.T - Bm[b] @ v.T * l[None, :]).max() / np.abs(A[b]).max() < 1e-11


@pytest.mark.parametrize('n_elem,n_modes', [(10000, 20), (100000, 50)])
def test_modal_config4_named_size(n_elem, n_modes):
    """Config 4 at its named size: 50 lowest modes of the 1e5-element cantilever.  The continuous member's
    frequencies (bending (beta L)^2 sqrt(EI / rho A L^4) and axial (2k-1) pi/2 sqrt(E/rho)/L families) are the
    comparator; the discretisation error of cubic Hermite elements is O(h^4) ~ 1e-16 here.  scipy's eigsh on the
    FP64-assembled matrices is 58x off on omega_1 at this size (profiles/r1_config4_beam_1e5.json)."""
    pk, beam = make(n_elem)
    out = beam.modal(n_modes=n_modes, tol=1e-10, maxit=60)
    assert out['converged'] and not out['failed']
    props = pk['props'][0].cpu().numpy()
    coords = pk['coords'][0].cpu().numpy()
    E, A, I, rho = props[0]
    exact = bo.cantilever_spectrum(float(coords[-1, 0] - coords[0, 0]), E, I, rho, A, n_modes)
    om = np.sqrt(out['lam'].cpu().numpy())
    assert np.abs(om[:5] / exact[:5] - 1).max() < 1e-9           # VERDICT bar: 1e-6
    assert np.abs(om / exact - 1).max() < 1e-7
    assert float(out['residual'].max().item()) < 1e-14
    Phi = out['phi']
    G = (Phi @ beam.matvec(Phi.t().contiguous(), mass=True)).cpu().numpy()
    assert np.abs(G - np.eye(n_modes)).max() < 1e-9


@pytest.mark.parametrize('n_elem', [300, 2000])
def test_modal_subspace_iteration(n_elem):
    length = 1000.0
    pk, beam = make(n_elem, length)
    out = beam.modal(n_modes=12, tol=1e-11, maxit=60)
    cond_est = 16 * n_elem ** 4
    assert out['converged'] or out['noise_floor'] < 1e-15 * cond_est        # stops at the eps*cond floor
    lam = out['lam'].cpu().numpy()
    coords, props = pk['coords'][0].cpu().numpy(), pk['props'][0].cpu().numpy()
    ls, ps = bo.modal(coords, props, [0, 1, 2], 12)
    tol = max(1e-9, 1e-15 * cond_est)
    assert np.abs(lam / ls - 1).max() < tol
    # closed form: bending omegas (the discretisation error of cubic Hermite elements is O(h^4))
    E, A, I, rho = props[0]
    _, _, w = bo.cantilever_closed_form(1.0, length, E, I, rho, A, 6)
    om = np.sqrt(lam)
    axial1 = math.pi / 2 * math.sqrt(E / rho) / length
    bending = om[om < 0.99 * axial1][:6]
    assert np.abs(bending / w[:len(bending)] - 1).max() < max(1e-4, 1e-16 * cond_est)
    # residuals and M-orthonormality
    assert float(out['residual'].max().item()) < 1e-10
    Phi = out['phi']
    MPhi = beam.matvec(Phi.t().contiguous(), mass=True)
    G = (Phi @ MPhi).cpu().numpy()
    assert np.abs(G - np.eye(12)).max() < 1e-9
    assert util.mode_err(Phi.cpu().numpy()[:3], ps[:3]).max() < 1e-5


@pytest.mark.parametrize('parts', [0, 40, -120, -900])
def test_general_chain_with_interior_supports(parts):
    """Not a cantilever: a bent chain of 3000 beams with varying sections, rollers / pins / a clamp at interior nodes
    (some of them on separator nodes of the partitioning) and four load vectors -- a well-conditioned problem (many
    supports): the double-double path agrees with scipy's sparse LU to 1e-10 (the FP64 comparator's own error; every
    partitioning lands on the same 1.1e-11) and leaves a scaled residual below 1e-14, on one and on two levels."""
    import scipy.sparse.linalg as spl
    rng = np.random.default_rng(31)
    nE = 3000
    ang = np.cumsum(rng.uniform(-0.02, 0.02, nE))
    step = rng.uniform(80.0, 120.0, nE)
    xy = np.zeros((nE + 1, 2))
    xy[1:, 0], xy[1:, 1] = np.cumsum(step * np.cos(ang)), np.cumsum(step * np.sin(ang))
    props = np.stack([rng.uniform(1.9e5, 2.2e5, nE), rng.uniform(2e3, 2e4, nE), rng.uniform(1e6, 1e8, nE),
                      rng.uniform(7e-9, 8.5e-9, nE)], axis=1)
    fixed = [0, 1, 2]
    for node in range(25, nE, 25):                                   # a support every 25 nodes
        kind = node // 25 % 3
        fixed += [3 * node + 1] if kind == 0 else ([3 * node, 3 * node + 1] if kind == 1 else [3 * node, 3 * node + 1, 3 * node + 2])
    fixed += [3 * nE + 1]
    T = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=DEV)     # noqa: E731
    beam = LargeBeam(T(xy), T(props), fixed, n_partitions=parts)
    K, M, keep = bo.chain_banded(xy, props, fixed)
    F = np.zeros((3 * (nE + 1), 4))
    F[keep] = rng.uniform(-1e3, 1e3, (len(keep), 4))
    U = beam.solve(T(F)).cpu().numpy()
    assert (U[np.asarray(fixed)] == 0).all()
    ref = spl.splu(K.tocsc()).solve(F[keep])
    assert util.relmax(U[keep], ref) < 1e-10
    KU = beam.matvec(T(U)).cpu().numpy()
    assert np.abs(KU[keep] - F[keep]).max() / (abs(K).max() * np.abs(U).max()) < 1e-14      # residual of the FP64-rounded U
