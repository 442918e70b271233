"""One LARGE chain structure (a single beam discretised in up to ~1e6 elements).

Static: partitioned block-tridiagonal Cholesky on the GPU (libbfe2 `bfe2_beam_*`).
Modal : shift-invert (sigma = 0) subspace iteration with a block of 64 vectors; the tall-skinny block
        products run on the FP64 tensor cores (DMMA), the 64 x 64 Rayleigh-Ritz problem on the same
        Cholesky-reduction + Jacobi kernel as the batched path.  The iteration loop below is host-side
        orchestration only; every flop is in a libbfe2 kernel (torch is used for allocation and a few
        element-wise copies / symmetrisation of 64 x 64 matrices).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

Q = 64          # block size of the subspace (DMMA kernels are specialised for it)


def _sp():
    return torch.cuda.current_stream().cuda_stream


class SubspaceIteration:
    """Shift-invert (sigma = 0) subspace iteration with a block of 64 vectors.  A subclass provides the operator:
    `n`, `device`, `fixed_mask` (uint8 [n], 1 = DOF held at zero), `solve(F, out=None)` = K^-1 F and
    `matvec(X, mass, out=None)` = K X or M X for F, X [n, 64]."""

    _dense_plan = None
    _gram_ws = None
    _eig_ws = None

    def _free_small(self):
        try:
            if getattr(self, '_dense_plan', None):
                _lib.lib().bfe2_plan_destroy(self._dense_plan)
                self._dense_plan = None
        except Exception:
            pass

    def gram(self, X, Y):
        """X^T Y on the FP64 tensor cores, X, Y [n, 64]."""
        L = _lib.lib()
        assert X.shape == (self.n, Q) and Y.shape == (self.n, Q) and X.is_contiguous() and Y.is_contiguous()
        if self._gram_ws is None:
            self._gram_ws = torch.empty(L.bfe2_gram_workspace_bytes(self.n) // 8, dtype=torch.float64, device=self.device)
        Cm = torch.empty((Q, Q), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(L.bfe2_gram64(self.n, _lib.ptr(X), _lib.ptr(Y), _lib.ptr(Cm), _lib.ptr(self._gram_ws), _sp()))
        return Cm

    def update(self, Z, Qm, out=None):
        """Z @ Qm on the FP64 tensor cores, Z [n, 64], Qm [64, 64]."""
        X = torch.empty_like(Z) if out is None else out
        Qm = Qm.contiguous()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().bfe2_update64(self.n, _lib.ptr(Z), _lib.ptr(Qm), _lib.ptr(X), _sp()))
        return X

    def small_eig(self, A, B):
        """Generalized symmetric eigenproblem of the 64 x 64 Rayleigh-Ritz matrices (Jacobi kernel).
        Returns theta [64] ascending and V [64, 64] with ROWS = eigenvectors, V B V^T = I."""
        L = _lib.lib()
        if self._dense_plan is None:
            h = C.c_void_p()
            _lib.check(L.bfe2_plan_create_dense(C.byref(h), Q))
            self._dense_plan = h
            self._eig_ws = torch.empty(2 * Q * Q, dtype=torch.float64, device=self.device)
        lam = torch.empty(Q, dtype=torch.float64, device=self.device)
        V = torch.empty((Q, Q), dtype=torch.float64, device=self.device)
        info = torch.zeros(1, dtype=torch.int32, device=self.device)
        sw = torch.zeros(1, dtype=torch.int32, device=self.device)
        A, B = A.contiguous(), B.contiguous()
        with torch.cuda.device(self.device):
            _lib.check(L.bfe2_sym_eig_dense(self._dense_plan, 1, _lib.ptr(A), _lib.ptr(B), _lib.ptr(lam), _lib.ptr(V),
                                            _lib.ptr(info), _lib.ptr(sw), _lib.ptr(self._eig_ws), _sp()))
        return lam, V, info


    def _rayleigh_ritz(self, A, B):
        """Ritz values theta (ascending) and coefficients Qm (columns, M-normalised: Q^T B Q = I) of the pencil
        A q = theta B q with A = Z^T K Z, B = Z^T M Z.

        Z = K^-1 M X has columns scaled like 1/lambda_j, so cond(B) ~ (lambda_q / lambda_1)^2 reaches 1e16 for a
        64-vector block on a beam spectrum: B has no usable Cholesky factor.  The pencil is therefore solved the
        other way round, B p = mu A p (mu = 1/theta, cond(A) is only the square root of cond(B)), after
        equilibrating with D = diag(A)^-1/2 and with an exactly-known shift delta that makes B + delta A
        safely positive definite:  (B + delta A) p = (mu + delta) A p."""
        A = 0.5 * (A + A.t())
        B = 0.5 * (B + B.t())
        d = torch.rsqrt(A.diagonal().clamp_min(1e-300))
        As = d[:, None] * A * d[None, :]
        Bs = d[:, None] * B * d[None, :]
        delta = 1e-9 * float(Bs.diagonal().max().item())
        info = None
        for _ in range(6):
            mus, V, info = self.small_eig(Bs + delta * As, As)          # rows p: p As p^T = 1
            if int(info.item()) == 0:
                break
            delta *= 100.0
        mu = (mus - delta).flip(0)                                      # descending mu = ascending theta
        P = V.flip(0)
        mu_c = mu.clamp_min(1e-18 * float(mu.max().item()))
        theta = 1.0 / mu_c
        Qm = (d[:, None] * P.t()) * torch.rsqrt(mu_c)[None, :]          # q_j = D p_j / sqrt(mu_j)
        return theta, Qm.contiguous(), info

    def modal(self, n_modes=50, tol=1e-10, maxit=80, seed=0, verbose=False):
        """Lowest n_modes (<= 56) eigenpairs by subspace iteration.  Returns dict(lam [n_modes],
        phi [n_modes, n] M-normalised, iterations, residual [n_modes] = ||K phi - lam M phi||_inf /
        (||K||_max ||phi||_inf), converged)."""
        assert 1 <= n_modes <= Q - 8, 'keep at least 8 guard vectors'
        g = torch.Generator(device=self.device)                 # start block drawn on the device: a host draw + copy of
        g.manual_seed(seed)                                     # n x 64 doubles cost more than the whole iteration at n = 3e5
        X = torch.randn((self.n, Q), generator=g, dtype=torch.float64, device=self.device)
        fm = torch.as_tensor(self.fixed_mask.astype(bool), device=self.device)
        X[fm] = 0.0
        Y = self.matvec(X, mass=True)
        Z = torch.empty_like(X)
        Y2 = torch.empty_like(X)
        theta_old = None
        it = 0
        conv = False
        best, stall = float('inf'), 0
        failed = False
        for it in range(1, maxit + 1):
            self.solve(Y, out=Z)                                   # Z = K^-1 M X
            self.matvec(Z, mass=True, out=Y2)                      # Y2 = M Z
            A = self.gram(Z, Y)                                    # Z^T K Z  (K Z = Y; no cancellation this way)
            B = self.gram(Z, Y2)                                   # Z^T M Z
            theta, Qm, info = self._rayleigh_ritz(A, B)
            if int(info.item()) != 0 or not bool(torch.isfinite(theta[:n_modes]).all()):
                # K^-1 is no longer a positive operator in FP64 (cond(K) ~ 16 n^4 > 1/eps): report, do not iterate
                failed = True
                break
            self.update(Z, Qm, out=X)                              # X = Z Q : X^T M X = I
            self.update(Y2, Qm, out=Y)                             # Y = M X
            th = theta[:n_modes]
            if theta_old is not None:
                change = float(((th - theta_old).abs() / th.abs()).max().item())
                if verbose:
                    print('it %d  max rel change %.3e  info %d' % (it, change, int(info.item())))
                theta_old = th.clone()
                if change < tol:
                    conv = True
                    break
                # cond(K) ~ n^4 puts a floor of ~eps*cond under the attainable change: stop when the change
                # has not improved for 4 iterations (the floor is reported, not hidden)
                if change < 0.7 * best:
                    best, stall = change, 0
                else:
                    stall += 1
                    if stall >= 4 and it >= 8:
                        break
            else:
                theta_old = th.clone()
        if theta_old is None:
            nan = torch.full((n_modes,), float('nan'), dtype=torch.float64, device=self.device)
            return dict(lam=nan, phi=None, iterations=it, residual=nan, converged=False, noise_floor=None, failed=True)
        lam = theta_old
        Phi = X[:, :n_modes]
        KX = self.matvec(X, mass=False)
        R = KX[:, :n_modes] - Y[:, :n_modes] * lam[None, :]
        kmax = float(self.matvec(torch.ones(self.n, dtype=torch.float64, device=self.device)).abs().max().item())
        res = R.abs().amax(dim=0) / (max(kmax, 1e-300) * Phi.abs().amax(dim=0))
        return dict(lam=lam, phi=Phi.t().contiguous(), iterations=it, residual=res, converged=conv,
                    noise_floor=None if conv else best, failed=failed)


class LargeBeam(SubspaceIteration):
    """coords [nN, 2], props [nE, 4] (E, A, I, rho): CUDA float64; nodes numbered along the chain (beam k joins
    nodes k and k+1); fixed_positions = Structure.positions_to_eliminate (global DOF positions).

    precision: 'dd' (default) evaluates element matrices, factorisation and substitutions in double-double --
    cond(K) ~ 16 n^4 makes the FP64-assembled problem meaningless at n >~ 1e4; 'fp64' runs the same kernels in
    plain FP64 (comparison only)."""

    def __init__(self, coords, props, fixed_positions, n_partitions: int = 0, precision: str = 'dd'):
        L = _lib.lib()
        assert coords.is_cuda and props.is_cuda and coords.dtype == torch.float64 and props.dtype == torch.float64
        assert precision in ('dd', 'fp64')
        self.precision = precision
        self.coords, self.props = coords.contiguous(), props.contiguous()
        nE = props.shape[0]
        assert coords.shape == (nE + 1, 2) and props.shape == (nE, 4)
        self.n_elem, self.n_nodes, self.n = nE, nE + 1, 3 * (nE + 1)
        mask = np.zeros(self.n, dtype=np.uint8)
        mask[np.asarray(list(fixed_positions), dtype=np.int64)] = 1
        self.fixed_mask = mask
        self.device = coords.device
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(L.bfe2_beam_create_ex(C.byref(h), nE, _lib.ptr(self.coords), _lib.ptr(self.props),
                                             _lib.ptr(mask), int(n_partitions), 1 if precision == 'dd' else 0, _sp()))
        self._h = h
        nn, nd, npart = C.c_int64(), C.c_int64(), C.c_int32()
        _lib.check(L.bfe2_beam_dims(h, C.byref(nn), C.byref(nd), C.byref(npart)))
        self.n_partitions = npart.value

    def __del__(self):
        try:
            if getattr(self, '_h', None):
                _lib.lib().bfe2_beam_destroy(self._h)
                self._h = None
            if getattr(self, '_dense_plan', None):
                _lib.lib().bfe2_plan_destroy(self._dense_plan)
                self._dense_plan = None
        except Exception:
            pass

    # ---- building blocks ---------------------------------------------------------------------
    def solve(self, F, out=None, with_lo=False):
        """U = K^-1 F for F [n, nrhs] (or [n]).  with_lo: also return the low words of the double-double
        solution (U + U_lo carries ~32 digits; zeros for an fp64 handle)."""
        vec = F.dim() == 1
        F2 = F.reshape(self.n, -1).contiguous()
        U = torch.empty_like(F2) if out is None else out
        assert U.is_contiguous() and U.shape == F2.shape
        Ulo = torch.empty_like(F2) if with_lo else None
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().bfe2_beam_solve_dd(self._h, F2.shape[1], _lib.ptr(F2), None, _lib.ptr(U),
                                                     _lib.ptr(Ulo), _sp()))
        if with_lo:
            return (U.reshape(self.n), Ulo.reshape(self.n)) if vec else (U, Ulo)
        return U.reshape(self.n) if vec else U

    def matvec(self, X, mass=False, out=None):
        vec = X.dim() == 1
        X2 = X.reshape(self.n, -1).contiguous()
        Y = torch.empty_like(X2) if out is None else out
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().bfe2_beam_matvec(self._h, 1 if mass else 0, X2.shape[1], _lib.ptr(X2), _lib.ptr(Y), _sp()))
        return Y.reshape(self.n) if vec else Y

    def residual(self, F, U, U_lo=None):
        """F - K (U + U_lo), evaluated in the handle's precision and rounded once."""
        F2 = F.reshape(self.n, -1).contiguous()
        U2 = U.reshape(self.n, -1).contiguous()
        L2 = None if U_lo is None else U_lo.reshape(self.n, -1).contiguous()
        R = torch.empty_like(F2)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().bfe2_beam_residual(self._h, F2.shape[1], _lib.ptr(F2), _lib.ptr(U2), _lib.ptr(L2),
                                                     _lib.ptr(R), _sp()))
        return R.reshape(F.shape)

    def blocks(self, mass=False):
        """The stored K (or M) rounded to FP64: [n_nodes, 2, 3, 3] = blocks (k, k) and (k+1, k)."""
        out = torch.empty((self.n_nodes, 2, 3, 3), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().bfe2_beam_export_blocks(self._h, 1 if mass else 0, _lib.ptr(out), _sp()))
        return out

    # ---- analyses ----------------------------------------------------------------------------
    def static(self, f, with_lo=False):
        """Displacements for the load vector(s) f [n] or [n, C]; supported DOFs are forced to zero load."""
        f = f.clone()
        fm = torch.as_tensor(self.fixed_mask.astype(bool), device=self.device)
        f[fm] = 0.0
        return self.solve(f, with_lo=with_lo)
