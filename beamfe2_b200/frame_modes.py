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
