"""First-k modes of ONE frame that is too large for the batched Jacobi path (n_free > 116).

Same shift-invert subspace iteration as the large-beam path (`beam.SubspaceIteration`: block of 64 vectors, Gram
products on the FP64 tensor cores, 64 x 64 Rayleigh-Ritz on the Jacobi kernel), with the operators of a general
plan: K and M banded from `bfe2_assemble_banded`; K^-1 applied to the 64 columns by ONE `bfe2_static_solve`
launch (each column is a "structure" of the batch that shares the band of K); K X and M X as banded products.
The reference returns ALL eigenpairs of such a structure through a dense LAPACK call (Solver.py:29-69); this
path returns the lowest <= 56, which is what a modal analysis of a large frame uses.
"""
from __future__ import annotations

import numpy as np
import torch

from . import batch
from .beam import Q, SubspaceIteration
from .plan import Plan


class FrameModes(SubspaceIteration):
    """coords [nN, 2], props [nE, 4] (E, A, I, rho), nodal_mass [nN, 2] or None: CUDA float64 of ONE structure."""

    def __init__(self, plan: Plan, coords, props, nodal_mass=None):
        assert coords.is_cuda and coords.dtype == torch.float64 and props.dtype == torch.float64
        self.plan = plan
        self.device = coords.device
        self.n = plan.n_free
        self.fixed_mask = np.zeros(self.n, dtype=np.uint8)            # the operators live on the free DOFs
        c1, p1 = coords.reshape(1, plan.n_nodes, 2).contiguous(), props.reshape(1, plan.n_elem, 4).contiguous()
        m1 = None if nodal_mass is None else nodal_mass.reshape(1, plan.n_nodes, 2).contiguous()
        Kb, Mb = batch.assemble_banded(plan, c1, p1, m1)
        self.Kb, self.Mb = Kb[0].contiguous(), Mb[0].contiguous()     # [hbw + 1, n]
        # the Q right-hand sides of one K^-1 application are Q "structures" with the same band and geometry
        self._Kb_rep = Kb.expand(Q, -1, -1).contiguous()
        self._c_rep = c1.expand(Q, -1, -1).contiguous()
        self._p_rep = p1.expand(Q, -1, -1).contiguous()
        self._pos = torch.as_tensor(plan.pos_of_free.astype(np.int64), device=self.device)
        self._f = torch.zeros((Q, 1, plan.sumdof), dtype=torch.float64, device=self.device)

    def __del__(self):
        self._free_small()

    def solve(self, F, out=None):
        """U = K^-1 F on the free DOFs, F [n, 64]."""
        assert F.shape == (self.n, Q)
        self._f[:, 0, self._pos] = F.t()
        r = batch.static_solve(self.plan, self._Kb_rep, self._c_rep, self._p_rep, self._f, None,
                               want_endforces=False, want_reactions=False)
        if int(r['info'].abs().sum().item()) != 0:
            raise np.linalg.LinAlgError('stiffness matrix is singular (pivot %d)' % int(r['info'].max().item()))
        U = r['u'][:, 0, self._pos].t()
        if out is None:
            return U.contiguous()
        out.copy_(U)
        return out

    def matvec(self, X, mass=False, out=None):
        """K X or M X from the band (libbfe2 `bfe2_band_matvec`: thread per entry of the result)."""
        from . import _lib
        band = self.Mb if mass else self.Kb
        vec = X.dim() == 1
        X2 = X.reshape(self.n, -1).contiguous()
        Y = torch.empty_like(X2) if (out is None or vec or not out.is_contiguous()) else out
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().bfe2_band_matvec(self.n, self.plan.hbw, X2.shape[1], _lib.ptr(band), _lib.ptr(X2),
                                                   _lib.ptr(Y), torch.cuda.current_stream().cuda_stream))
        if vec:
            Y = Y.reshape(self.n)
        if out is not None and Y is not out:
            out.copy_(Y)
            return out
        return Y

    def modes(self, n_modes=20, tol=1e-11, maxit=80):
        """dict(lam [k] ascending, phi [k, sumdof] M-normalised with zeros at supports, iterations, residual, converged)."""
        n_modes = int(min(n_modes, Q - 8, self.n))
        out = self.modal(n_modes=n_modes, tol=tol, maxit=maxit)
        if out['phi'] is not None:
            full = torch.zeros((n_modes, self.plan.sumdof), dtype=torch.float64, device=self.device)
            full[:, self._pos] = out['phi']
            out['phi'] = full
        return out


class BatchedFrameModes:
    """First-k modes of MANY frames that share one topology plan and are too large for the batched all-modes kernel
    (n_free > 116) -- the same shift-invert subspace iteration as `FrameModes`, with the batch as
    an extra axis of every operator, so that B structures cost one sequence of launches instead of B:

      K^-1 on the 64 vectors  `bfe2_static_solve` with the vectors as load cases, 8 at a time (one band factorisation
                              per launch and structure)
      K X, M X                `bfe2_band_matvec_batched`
      X^T Y, Z Q              `bfe2_gram64_batched` / `bfe2_update64_batched` (FP64 tensor cores, block per structure)
      64 x 64 Rayleigh-Ritz   `bfe2_sym_eig_dense` on the whole batch (Cholesky reduction + Jacobi kernel)

    coords [B, nN, 2], props [B, nE, 4], nodal_mass [B, nN, 2] or None: CUDA float64."""

    RHS_PER_LAUNCH = 8

    def __init__(self, plan: Plan, coords, props, nodal_mass=None):
        assert coords.is_cuda and coords.dtype == torch.float64 and props.dtype == torch.float64 and coords.dim() == 3
        self.plan, self.device = plan, coords.device
        self.B, self.n = coords.shape[0], plan.n_free
        self.coords, self.props = coords.contiguous(), props.contiguous()
        self.Kb, self.Mb = batch.assemble_banded(plan, self.coords, self.props, nodal_mass)
        self._pos = torch.as_tensor(plan.pos_of_free.astype(np.int64), device=self.device)
        self._f = torch.zeros((self.B, self.RHS_PER_LAUNCH, plan.sumdof), dtype=torch.float64, device=self.device)
        self._dense_plan = None
        self._eig_ws = None

    def __del__(self):
        try:
            if getattr(self, '_dense_plan', None):
                from . import _lib
                _lib.lib().bfe2_plan_destroy(self._dense_plan)
                self._dense_plan = None
        except Exception:
            pass

    # ---- operators on [B, n, 64] multi-vectors ------------------------------------------------
    def solve(self, F):
        U = torch.empty_like(F)
        C = self.RHS_PER_LAUNCH
        for c0 in range(0, Q, C):
            self._f[:, :, self._pos] = F[:, :, c0:c0 + C].transpose(1, 2)
            r = batch.static_solve(self.plan, self.Kb, self.coords, self.props, self._f, None, want_endforces=False,
                                   want_reactions=False)
            if c0 == 0 and int(r['info'].abs().sum().item()) != 0:
                raise np.linalg.LinAlgError('stiffness matrix is singular (pivot %d)' % int(r['info'].max().item()))
            U[:, :, c0:c0 + C] = r['u'][:, :, self._pos].transpose(1, 2)
        return U

    def matvec(self, X, mass=False):
        from . import _lib
        Y = torch.empty_like(X)
        band = self.Mb if mass else self.Kb
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().bfe2_band_matvec_batched(self.n, self.plan.hbw, Q, self.B, _lib.ptr(band), _lib.ptr(X),
                                                           _lib.ptr(Y), torch.cuda.current_stream().cuda_stream))
        return Y

    def gram(self, X, Y):
        from . import _lib
        Cm = torch.empty((self.B, Q, Q), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().bfe2_gram64_batched(self.n, self.B, _lib.ptr(X), _lib.ptr(Y), _lib.ptr(Cm),
                                                      torch.cuda.current_stream().cuda_stream))
        return Cm

    def update(self, Z, Qm):
        from . import _lib
        X = torch.empty_like(Z)
        Qm = Qm.contiguous()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().bfe2_update64_batched(self.n, self.B, _lib.ptr(Z), _lib.ptr(Qm), _lib.ptr(X),
                                                        torch.cuda.current_stream().cuda_stream))
        return X

    def small_eig(self, A, Bm):
        """batched 64 x 64 generalized symmetric eigenproblem: theta [B, 64] ascending, V [B, 64, 64] (rows = vectors)"""
        import ctypes as C
        from . import _lib
        L = _lib.lib()
        if self._dense_plan is None:
            h = C.c_void_p()
            _lib.check(L.bfe2_plan_create_dense(C.byref(h), Q))
            self._dense_plan = h
            self._eig_ws = torch.empty(2 * self.B * Q * Q, dtype=torch.float64, device=self.device)
        lam = torch.empty((self.B, Q), dtype=torch.float64, device=self.device)
        V = torch.empty((self.B, Q, Q), dtype=torch.float64, device=self.device)
        info = torch.zeros(self.B, dtype=torch.int32, device=self.device)
        sw = torch.zeros(self.B, dtype=torch.int32, device=self.device)
        A, Bm = A.contiguous(), Bm.contiguous()
        with torch.cuda.device(self.device):
            _lib.check(L.bfe2_sym_eig_dense(self._dense_plan, self.B, _lib.ptr(A), _lib.ptr(Bm), _lib.ptr(lam), _lib.ptr(V),
                                            _lib.ptr(info), _lib.ptr(sw), _lib.ptr(self._eig_ws),
                                            torch.cuda.current_stream().cuda_stream))
        return lam, V, info

    def _rayleigh_ritz(self, A, Bm):
        """batched form of SubspaceIteration._rayleigh_ritz: the pencil is solved as B p = mu A p (mu = 1 / theta) after
        diagonal equilibration, with an exactly known shift that makes the left matrix safely definite"""
        A = 0.5 * (A + A.transpose(1, 2))
        Bm = 0.5 * (Bm + Bm.transpose(1, 2))
        d = torch.rsqrt(A.diagonal(dim1=1, dim2=2).clamp_min(1e-300))
        As = d[:, :, None] * A * d[:, None, :]
        Bs = d[:, :, None] * Bm * d[:, None, :]
        delta = 1e-9 * Bs.diagonal(dim1=1, dim2=2).amax(dim=1)
        for _ in range(6):
            mus, V, info = self.small_eig(Bs + delta[:, None, None] * As, As)
            bad = info != 0
            if not bool(bad.any()):
                break
            delta = torch.where(bad, delta * 100.0, delta)
        mu = (mus - delta[:, None]).flip(1)
        P = V.flip(1)
        mu_c = torch.maximum(mu, 1e-18 * mu.amax(dim=1, keepdim=True))
        theta = 1.0 / mu_c
        Qm = (d[:, :, None] * P.transpose(1, 2)) * torch.rsqrt(mu_c)[:, None, :]
        return theta, Qm.contiguous(), info

    def modes(self, n_modes=20, tol=1e-11, maxit=80, seed=0):
        """dict(lam [B, k] ascending, phi [B, k, sumdof] M-normalised with zeros at supports, iterations,
        converged [B], residual [B, k] = |K phi - lam M phi|_inf / (|K|_max |phi|_inf))."""
        n_modes = int(min(n_modes, Q - 8, self.n))
        g = torch.Generator(device='cpu')
        g.manual_seed(seed)
        X = torch.randn((self.n, Q), generator=g, dtype=torch.float64).to(self.device)
        X = X[None].repeat(self.B, 1, 1).contiguous()
        Y = self.matvec(X, mass=True)
        theta_old = None
        conv = torch.zeros(self.B, dtype=torch.bool, device=self.device)
        it = 0
        for it in range(1, maxit + 1):
            Z = self.solve(Y)                                   # Z = K^-1 M X
            Y2 = self.matvec(Z, mass=True)
            A = self.gram(Z, Y)                                 # Z^T K Z
            Bm = self.gram(Z, Y2)                               # Z^T M Z
            theta, Qm, info = self._rayleigh_ritz(A, Bm)
            X = self.update(Z, Qm)
            Y = self.update(Y2, Qm)
            th = theta[:, :n_modes]
            if theta_old is not None:
                change = ((th - theta_old).abs() / th.abs()).amax(dim=1)
                conv = change < tol
                if bool(conv.all()):
                    theta_old = th.clone()
                    break
            theta_old = th.clone()
        lam = theta_old
        KX = self.matvec(X, mass=False)
        R = KX[:, :, :n_modes] - Y[:, :, :n_modes] * lam[:, None, :]
        kmax = self.Kb.abs().amax(dim=(1, 2))
        Phi = X[:, :, :n_modes]
        res = R.abs().amax(dim=1) / (kmax[:, None] * Phi.abs().amax(dim=1))
        full = torch.zeros((self.B, n_modes, self.plan.sumdof), dtype=torch.float64, device=self.device)
        full[:, :, self._pos] = Phi.transpose(1, 2)
        return dict(lam=lam, phi=full, iterations=it, converged=conv, residual=res)
