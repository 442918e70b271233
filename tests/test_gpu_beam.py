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


def make(n_elem, length=1000.0, n_partitions=0):
    pk = sweep.beam_pack(n_elem, length=length, device=DEV)
    beam = LargeBeam(pk['coords'][0], pk['props'][0], [0, 1, 2], n_partitions=n_partitions)
    return pk, beam


@pytest.mark.parametrize('n_elem,parts', [(3, 0), (10, 1), (10, 3), (257, 0), (257, 5), (1000, 0), (1000, 17), (4096, 64)])
def test_static_vs_banded_cholesky_and_closed_form(n_elem, parts):
    length = 1000.0
    pk, beam = make(n_elem, length, parts)
    f = pk['f_nodal'][0, 0].clone()
    u = beam.static(f)
    un = u.cpu().numpy()
    assert (un[:3] == 0).all()
    coords, props = pk['coords'][0].cpu().numpy(), pk['props'][0].cpu().numpy()
    K, M, keep = bo.chain_banded(coords, props, [0, 1, 2])
    # (i) scaled residual ||K u - f|| / (||K|| ||u||)
    r = K @ un[keep] - f.cpu().numpy()[keep]
    res = np.abs(r).max() / (abs(K).max() * np.abs(un).max())
    assert res < 1e-14
    # (ii) closed form (nodal values of Hermitian elements are exact): tip deflection PL^3/3EI, rotation PL^2/2EI
    E, A, I = props[0, 0], props[0, 1], props[0, 2]
    d, rot, _ = bo.cantilever_closed_form(100.0, length, E, I, props[0, 3], A, 1)
    us = bo.static(coords, props, [0, 1, 2], f.cpu().numpy())
    err_gpu = abs(un[-2] / d - 1)
    err_cpu = abs(us[-2] / d - 1)
    cond_est = 16 * n_elem ** 4                      # ~ cond(K) of the cantilever
    assert err_gpu < max(1e-11, 20 * err_cpu, 1e-16 * cond_est)
    assert abs(un[-1] / rot - 1) < max(1e-11, 20 * abs(us[-1] / rot - 1), 1e-16 * cond_est)
    # (iii) against scipy.linalg.solveh_banded within the conditioning
    assert util.relmax(un, us) < max(1e-11, 1e-15 * cond_est)


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
