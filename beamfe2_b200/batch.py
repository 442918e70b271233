"""Batched front end: many structure variants sharing one topology plan -> libbfe2 kernels.

This is the data-parallel axis the reference does not have (it solves one Structure at a time in
Python, BeamFE2/Solver.py:29-87).  Tensors are torch CUDA float64; nothing here computes on the CPU.
"""
from __future__ import annotations

import torch

from . import _lib
from .plan import Plan


def _stream_ptr():
    return torch.cuda.current_stream().cuda_stream


MODAL_ENGINES = {'jacobi': 0, 'two_ended': 1}


def set_modal_engine(name: str):
    """Process-wide choice of the modal engine of the fused path (include/bfe2.h): 'jacobi' (default) = one-sided Jacobi
    on the pivoted-Cholesky factor; 'two_ended' (n_free <= 64, n_modes <= 16) = tridiagonalisation of C and of C^-1 +
    Sturm bisection + inverse iteration, with the Jacobi engine repairing what it declines.  Returns the previous name."""
    L = _lib.lib()
    prev = {v: k for k, v in MODAL_ENGINES.items()}[L.bfe2_get_modal_engine()]
    _lib.check(L.bfe2_set_modal_engine(MODAL_ENGINES[name]))
    return prev


def _chk(t, shape, name, dtype=torch.float64):
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.Bfe2Error('%s must be a CUDA tensor (libbfe2 has no CPU fallback)' % name)
    if t.dtype != dtype:
        raise _lib.Bfe2Error('%s must be %s' % (name, dtype))
    if tuple(t.shape) != tuple(shape):
        raise _lib.Bfe2Error('%s has shape %s, expected %s' % (name, tuple(t.shape), tuple(shape)))
    return t.contiguous()


def element_km(xi, yi, xj, yj, E, A, I, rho, lumped=None):
    """K1: per-element global 6x6 K and M.  All inputs [nE] CUDA float64 (lumped: uint8 or None)."""
    nE = xi.numel()
    args = [_chk(t, (nE,), 'elem input') for t in (xi, yi, xj, yj, E, A, I, rho)]
    if lumped is not None:
        lumped = _chk(lumped, (nE,), 'lumped', torch.uint8)
    Kg = torch.empty((nE, 6, 6), dtype=torch.float64, device=xi.device)
    Mg = torch.empty((nE, 6, 6), dtype=torch.float64, device=xi.device)
    with torch.cuda.device(xi.device):
        _lib.check(_lib.lib().bfe2_element_km(nE, *[_lib.ptr(a) for a in args], _lib.ptr(lumped),
                                             _lib.ptr(Kg), _lib.ptr(Mg), _stream_ptr()))
    return Kg, Mg


def assemble_banded(plan: Plan, coords, props, nodal_mass=None, want_mass=True):
    """K2: condensed band matrices Kb, Mb [B, hbw+1, n]."""
    B = coords.shape[0]
    coords = _chk(coords, (B, plan.n_nodes, 2), 'coords')
    props = _chk(props, (B, plan.n_elem, 4), 'props')
    nodal_mass = _chk(nodal_mass, (B, plan.n_nodes, 2), 'nodal_mass')
    Kb = torch.empty((B, plan.hbw + 1, plan.n_free), dtype=torch.float64, device=coords.device)
    Mb = torch.empty_like(Kb) if want_mass else None
    with torch.cuda.device(coords.device):
        _lib.check(_lib.lib().bfe2_assemble_banded(plan.handle, B, _lib.ptr(coords), _lib.ptr(props),
                                                  _lib.ptr(nodal_mass), _lib.ptr(Kb), _lib.ptr(Mb),
                                                  _stream_ptr()))
    return Kb, Mb


def static_solve(plan: Plan, Kb, coords, props, f_nodal, r_local=None, want_endforces=True,
                 want_reactions=True):
    """K3: u [B,C,sumdof], endforces [B,C,nE,6], reactions [B,C,sumdof], info [B]."""
    B, C = f_nodal.shape[0], f_nodal.shape[1]
    dev = f_nodal.device
    Kb = _chk(Kb, (B, plan.hbw + 1, plan.n_free), 'Kb')
    coords = _chk(coords, (B, plan.n_nodes, 2), 'coords')
    props = _chk(props, (B, plan.n_elem, 4), 'props')
    f_nodal = _chk(f_nodal, (B, C, plan.sumdof), 'f_nodal')
    r_local = _chk(r_local, (B, C, plan.n_elem, 6), 'r_local')
    u = torch.empty((B, C, plan.sumdof), dtype=torch.float64, device=dev)
    ef = torch.empty((B, C, plan.n_elem, 6), dtype=torch.float64, device=dev) if want_endforces else None
    re = torch.empty((B, C, plan.sumdof), dtype=torch.float64, device=dev) if want_reactions else None
    info = torch.empty((B,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().bfe2_static_solve(plan.handle, B, C, _lib.ptr(Kb), _lib.ptr(coords),
                                               _lib.ptr(props), _lib.ptr(f_nodal), _lib.ptr(r_local),
                                               _lib.ptr(u), _lib.ptr(ef), _lib.ptr(re), _lib.ptr(info),
                                               _stream_ptr()))
    return dict(u=u, endforces=ef, reactions=re, info=info)


def modal_solve(plan: Plan, Kb, Mb, n_modes=None):
    """K4: lam [B,n] ascending, phi [B,n_modes,sumdof] M-normalised, info [B], sweeps [B]."""
    B = Kb.shape[0]
    dev = Kb.device
    n_modes = plan.n_free if n_modes is None else int(n_modes)
    Kb = _chk(Kb, (B, plan.hbw + 1, plan.n_free), 'Kb')
    Mb = _chk(Mb, (B, plan.hbw + 1, plan.n_free), 'Mb')
    lam = torch.empty((B, plan.n_free), dtype=torch.float64, device=dev)
    phi = torch.empty((B, n_modes, plan.sumdof), dtype=torch.float64, device=dev) if n_modes > 0 else None
    info = torch.empty((B,), dtype=torch.int32, device=dev)
    sweeps = torch.empty((B,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().bfe2_modal_solve(plan.handle, B, n_modes, _lib.ptr(Kb), _lib.ptr(Mb),
                                              _lib.ptr(lam), _lib.ptr(phi), _lib.ptr(info),
                                              _lib.ptr(sweeps), _stream_ptr()))
    return dict(lam=lam, phi=phi, info=info, sweeps=sweeps)


def assemble_geometric(plan: Plan, coords, axial):
    """Condensed band Kgb [B, hbw+1, n] of the geometric stiffness K_g(N) for axial forces axial [B, nE] (tension
    positive; HermitianBeam_2D.py:261-282, Structure.py:211-214)."""
    B = coords.shape[0]
    coords = _chk(coords, (B, plan.n_nodes, 2), 'coords')
    axial = _chk(axial, (B, plan.n_elem), 'axial')
    Kgb = torch.empty((B, plan.hbw + 1, plan.n_free), dtype=torch.float64, device=coords.device)
    with torch.cuda.device(coords.device):
        _lib.check(_lib.lib().bfe2_assemble_geometric(plan.handle, B, _lib.ptr(coords), _lib.ptr(axial), _lib.ptr(Kgb),
                                                      _stream_ptr()))
    return Kgb


def buckling(plan: Plan, coords, props, f_nodal, r_local=None, n_modes=3, load_case=0):
    """Linear buckling of a batch (the reference's BucklingSolver, Solver.py:90-180, is unfinished).

    1. static solve for the reference loads -> axial forces N_e = endforces[.., e, 3] (tension positive);
    2. K_g(N) assembled on the device (bfe2_assemble_geometric);
    3. (K + lambda K_g) phi = 0  <=>  (-K_g) phi = mu K phi, mu = 1 / lambda.  Both passes run on the dense
       Cholesky-reduction + Jacobi kernel (bfe2_modal_solve) with K -- positive definite -- in M's role:
         pass 1 (eigenvalues only): (-K_g + s K, K), a small shift s making the left matrix definite -> mu_max, mu_min
         pass 2: (K_g + sigma K, K) with sigma = 2 mu_max: its LOWEST eigenvalues theta = sigma - mu belong to the
                 largest mu, i.e. to the smallest positive critical factors, and come with their shapes.
    Returns dict(load_factor [B, n_modes] ascending (inf where fewer positive factors exist), phi [B, n_modes, sumdof]
    (K-normalised: phi^T K phi = 1), axial [B, nE], info [B])."""
    B = coords.shape[0]
    dev = coords.device
    n = plan.n_free
    st = solve_fused(plan, coords, props, None, f_nodal, r_local, modal=False, want_reactions=False)
    axial = st['endforces'][:, load_case, :, 3].contiguous()
    Kb, _ = assemble_banded(plan, coords, props, None, want_mass=False)
    Kg = assemble_geometric(plan, coords, axial)
    info = st['info'].clone()
    # scale of mu: |K_g| / |K| on the diagonals
    ratio = (Kg[:, 0].abs().amax(dim=1) / Kb[:, 0].abs().amax(dim=1)).clamp_min(1e-300)
    mu_max = torch.zeros(B, dtype=torch.float64, device=dev)
    todo = info == 0
    shift = 1e-9 * ratio
    for _ in range(8):                                   # members in tension make -K_g indefinite: raise the shift
        if not bool(todo.any()):
            break
        idx = torch.nonzero(todo).reshape(-1)
        A = -Kg[idx] + shift[idx, None, None] * Kb[idx]
        r = modal_solve(plan, A, Kb[idx], n_modes=0)
        ok = r['info'] == 0
        mu_max[idx[ok]] = r['lam'][ok, -1] - shift[idx[ok]]
        todo[idx[ok]] = False
        shift[idx[~ok]] *= 100.0
    info[todo] = -2
    good = (info == 0) & (mu_max > 0)                    # mu_max <= 0: nothing buckles under this load
    lf = torch.full((B, n_modes), float('inf'), dtype=torch.float64, device=dev)
    phi = torch.zeros((B, n_modes, plan.sumdof), dtype=torch.float64, device=dev)
    if bool(good.any()):
        idx = torch.nonzero(good).reshape(-1)
        sigma = 2.0 * mu_max[idx]
        A = Kg[idx] + sigma[:, None, None] * Kb[idx]
        r = modal_solve(plan, A, Kb[idx], n_modes=n_modes)
        mu = sigma[:, None] - r['lam'][:, :n_modes]
        lf[idx] = torch.where(mu > 0, 1.0 / mu.clamp_min(1e-300), torch.full_like(mu, float('inf')))
        phi[idx] = r['phi']
        bad = r['info'] != 0
        info[idx[bad]] = r['info'][bad]
    return dict(load_factor=lf, phi=phi, axial=axial, info=info)


def modal_lumped_condensed(plan: Plan, coords, props, nodal_mass=None, n_modes=None):
    """Lumped-mass modal analysis WITHOUT the reference's m * 1e-10 rotary terms (HermitianBeam_2D.py:225-231, which make
    cond(M) ~ 1e10 and produce negative eigenvalues): the rotations are massless, M is positive
    SEMI-definite, and the pencil is solved the other way round, M phi = mu K phi (mu = 1 / lam, K positive definite in
    M's role) -- mathematically the static (Guyan) condensation of the massless rotations, which is exact here.  The
    n_t = number of free translations finite modes come out with mu > 0; the massless directions have mu = 0 exactly
    (lam = inf) and are dropped.  `plan` must have been created with every beam flagged lumped.
    Returns dict(lam [B, n_t] ascending, phi [B, n_modes, sumdof] M-normalised, n_t, info)."""
    B = coords.shape[0]
    n = plan.n_free
    assert plan.lumped is not None and bool(plan.lumped.all()), 'every beam must use the lumped mass matrix'
    Kb, Mb = assemble_banded(plan, coords, props, nodal_mass)
    pos = torch.as_tensor(plan.pos_of_free.astype('int64'), device=coords.device)
    rot = (pos % 3) == 2
    Mb = Mb.clone()
    Mb[:, 0, rot] = 0.0                                  # exactly massless rotations
    n_t = int((~rot).sum().item())
    n_modes = n_t if n_modes is None else min(int(n_modes), n_t)
    # shift s: M + s K is positive definite and (M + s K) phi = (mu + s) K phi EXACTLY, so s costs no accuracy; it only
    # has to keep the 30-odd massless directions (mu = 0) above the pivot test of the kernel, i.e. s ~ 1e-9 mu_max.
    # mu_max is estimated by one step of the power iteration x = K^-1 M 1 (one launch of the static kernel):
    # mu_est = x^T M x / x^T K x <= mu_max, and close to it for this smooth start vector.
    mdiag = torch.zeros((B, 1, plan.sumdof), dtype=torch.float64, device=coords.device)
    mdiag[:, 0, pos] = Mb[:, 0]
    x = static_solve(plan, Kb, coords, props, mdiag, None, want_endforces=False, want_reactions=False)['u'][:, 0]
    mu_est = (mdiag[:, 0] * x * x).sum(dim=1) / (mdiag[:, 0] * x).sum(dim=1)
    shift = 1e-9 * mu_est
    r = modal_solve(plan, Mb + shift[:, None, None] * Kb, Kb, n_modes=n)     # ascending theta = mu + s, all shapes
    mu = (r['lam'] - shift[:, None]).flip(1)[:, :n_t]    # descending mu = ascending lam
    lam = 1.0 / mu
    phi = r['phi'].flip(1)[:, :n_modes] / torch.sqrt(mu[:, :n_modes])[:, :, None]   # phi^T K phi = 1 -> phi^T M phi = 1
    return dict(lam=lam, phi=phi.contiguous(), n_t=n_t, info=r['info'])


def solve_fused(plan: Plan, coords, props, nodal_mass=None, f_nodal=None, r_local=None, modal=True,
                n_modes=None, want_endforces=True, want_reactions=True, out=None):
    """Fused K1->K4 for a batch.  Returns dict(u, endforces, reactions, lam, phi, info, sweeps);
    the static entries are None when f_nodal is None, the modal ones when modal is False.
    `out` may carry preallocated output tensors (same keys) to avoid allocations in timed loops."""
    B = coords.shape[0]
    dev = coords.device
    coords = _chk(coords, (B, plan.n_nodes, 2), 'coords')
    props = _chk(props, (B, plan.n_elem, 4), 'props')
    nodal_mass = _chk(nodal_mass, (B, plan.n_nodes, 2), 'nodal_mass')
    C = 0
    if f_nodal is not None:
        C = f_nodal.shape[1]
        f_nodal = _chk(f_nodal, (B, C, plan.sumdof), 'f_nodal')
        r_local = _chk(r_local, (B, C, plan.n_elem, 6), 'r_local')
    n_modes = (plan.n_free if n_modes is None else int(n_modes)) if modal else 0
    out = {} if out is None else out

    def buf(key, shape, dtype=torch.float64, need=True):
        if not need:
            return None
        t = out.get(key)
        if t is None:
            t = torch.empty(shape, dtype=dtype, device=dev)
        elif not t.is_contiguous():
            # _chk() would make a contiguous COPY: the kernel would fill the copy and the caller's buffer never
            raise ValueError('output buffer %r must be contiguous' % key)
        return _chk(t, shape, key, dtype)

    u = buf('u', (B, C, plan.sumdof), need=C > 0)
    ef = buf('endforces', (B, C, plan.n_elem, 6), need=C > 0 and want_endforces)
    re = buf('reactions', (B, C, plan.sumdof), need=C > 0 and want_reactions)
    lam = buf('lam', (B, plan.n_free), need=modal)
    phi = buf('phi', (B, n_modes, plan.sumdof), need=modal and n_modes > 0)
    info = buf('info', (B,), torch.int32)
    sweeps = buf('sweeps', (B,), torch.int32)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().bfe2_solve_fused(
            plan.handle, B, C, 1 if modal else 0, n_modes, _lib.ptr(coords), _lib.ptr(props),
            _lib.ptr(nodal_mass), _lib.ptr(f_nodal), _lib.ptr(r_local), _lib.ptr(u), _lib.ptr(ef),
            _lib.ptr(re), _lib.ptr(lam), _lib.ptr(phi), _lib.ptr(info), _lib.ptr(sweeps), _stream_ptr()))
    return dict(u=u, endforces=ef, reactions=re, lam=lam, phi=phi, info=info, sweeps=sweeps)


LOAD_KINDS = {'concentrated perpendicular force': 1, 'concentrated axial force': 2, 'concentrated moment': 3}


def internal_actions(plan: Plan, coords, endforces, q_axial=None, q_perp=None, npts=11, want_field=True, conc=None):
    """N / V / M at npts equidistant positions of every beam (HermitianBeam_2D.py:398-431) and their per-beam maxima.
    q_axial / q_perp [B, C, nE]: uniform beam loads; conc [B, C, nE, n_conc, 3]: up to n_conc concentrated loads per
    beam as (kind, value, normalised position), kind from LOAD_KINDS (0 = empty slot).
    Returns (actions [B, C, nE, 3, npts] or None, maxabs [B, C, nE, 3])."""
    B, C = endforces.shape[0], endforces.shape[1]
    dev = endforces.device
    coords = _chk(coords, (B, plan.n_nodes, 2), 'coords')
    endforces = _chk(endforces, (B, C, plan.n_elem, 6), 'endforces')
    q_axial = _chk(q_axial, (B, C, plan.n_elem), 'q_axial')
    q_perp = _chk(q_perp, (B, C, plan.n_elem), 'q_perp')
    n_conc = 0
    if conc is not None:
        n_conc = int(conc.shape[3])
        conc = _chk(conc, (B, C, plan.n_elem, n_conc, 3), 'conc')
    act = torch.empty((B, C, plan.n_elem, 3, npts), dtype=torch.float64, device=dev) if want_field else None
    mx = torch.empty((B, C, plan.n_elem, 3), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().bfe2_internal_actions_ex(plan.handle, B, C, int(npts), _lib.ptr(coords), _lib.ptr(endforces),
                                                       _lib.ptr(q_axial), _lib.ptr(q_perp), n_conc, _lib.ptr(conc),
                                                       _lib.ptr(act), _lib.ptr(mx), _stream_ptr()))
    return act, mx


def modal_participation(plan: Plan, Mb, phi):
    """gamma [B, n_modes, 2] = phi^T M r_x, phi^T M r_y; effective modal masses are gamma ** 2.
    `plan` / `Mb`: the UNSUPPORTED plan (Plan(n_nodes, conn, [])) and its assembled band mass matrix."""
    B, n_modes = phi.shape[0], phi.shape[1]
    Mb = _chk(Mb, (B, plan.hbw + 1, plan.n_free), 'Mb')
    phi = _chk(phi, (B, n_modes, plan.sumdof), 'phi')
    gamma = torch.empty((B, n_modes, 2), dtype=torch.float64, device=phi.device)
    with torch.cuda.device(phi.device):
        _lib.check(_lib.lib().bfe2_modal_participation(plan.handle, B, n_modes, _lib.ptr(Mb), _lib.ptr(phi),
                                                       _lib.ptr(gamma), _stream_ptr()))
    return gamma


class PackedIO:
    """Packed per-structure records with compact load / result layouts (include/bfe2.h: bfe2_pattern_*).

    f_index / r_index: the dense indices (into [C, sumdof] and [C, nE, 6]) of the load entries that may be non-zero
    anywhere in the batch -- the reference's add_nodal_load / add_internal_loads touch a handful of positions
    (Structure.py:244-287).  `PackedIO.from_dense` derives them from example tensors.  None = every entry."""

    IN_FIELDS = ('coords', 'props', 'nodal_mass', 'f_values', 'r_values')
    OUT_FIELDS = ('u', 'endforces', 'reactions', 'lam', 'phi', 'status')

    def __init__(self, plan: Plan, n_loadcases: int, n_modes: int, modal: bool = True, f_index=None, r_index=None):
        import ctypes as C
        import numpy as np
        self.plan, self.C, self.modal = plan, int(n_loadcases), bool(modal)
        self.n_modes = int(n_modes) if modal else 0
        nf, nr = self.C * plan.sumdof, self.C * plan.n_elem * 6
        self.f_index = np.arange(nf, dtype=np.int32) if f_index is None else np.ascontiguousarray(f_index, dtype=np.int32)
        self.r_index = np.arange(nr, dtype=np.int32) if r_index is None else np.ascontiguousarray(r_index, dtype=np.int32)
        h = C.c_void_p()
        _lib.check(_lib.lib().bfe2_pattern_create(C.byref(h), plan.handle, self.C, 1 if modal else 0, self.n_modes,
                                                  int(self.f_index.size), _lib.ptr(self.f_index) if self.f_index.size else None,
                                                  int(self.r_index.size), _lib.ptr(self.r_index) if self.r_index.size else None))
        self._h = h
        wi, wo = C.c_int32(), C.c_int32()
        oi, oo = (C.c_int32 * 5)(), (C.c_int32 * 6)()
        _lib.check(_lib.lib().bfe2_pattern_layout(h, C.byref(wi), C.byref(wo), oi, oo))
        self.w_in, self.w_out = wi.value, wo.value
        self.in_off, self.out_off = list(oi), list(oo)
        p, Cn, n = plan, self.C, plan.n_free
        self.n_fixed = int(plan.fixed.size)
        self.in_shapes = dict(coords=(p.n_nodes, 2), props=(p.n_elem, 4), nodal_mass=(p.n_nodes, 2),
                              f_values=(int(self.f_index.size),), r_values=(int(self.r_index.size),))
        self.out_shapes = dict(u=(Cn, n), endforces=(Cn, p.n_elem, 6), reactions=(Cn, self.n_fixed),
                               lam=(n if modal else 0,), phi=(self.n_modes, n), status=(1,))

    @classmethod
    def from_dense(cls, plan, f_nodal, r_local, n_modes, modal=True):
        """Load pattern = union over the batch of the non-zero entries of example dense tensors
        f_nodal [B, C, sumdof], r_local [B, C, nE, 6] (either may be None)."""
        import numpy as np
        C = 0 if f_nodal is None else int(f_nodal.shape[1])
        fi = np.zeros(0, dtype=np.int32)
        ri = np.zeros(0, dtype=np.int32)
        if f_nodal is not None:
            fi = torch.nonzero((f_nodal != 0).any(dim=0).reshape(-1)).reshape(-1).cpu().numpy().astype(np.int32)
        if r_local is not None:
            ri = torch.nonzero((r_local != 0).any(dim=0).reshape(-1)).reshape(-1).cpu().numpy().astype(np.int32)
        return cls(plan, C, n_modes, modal, fi, ri)

    def __del__(self):
        try:
            if getattr(self, '_h', None):
                _lib.lib().bfe2_pattern_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ---- views into records ([B, w] tensors on any device) -------------------------------------
    def _views(self, rec, offs, shapes, names):
        import math
        out = {}
        for name, o in zip(names, offs):
            shp = shapes[name]
            k = math.prod(shp)
            out[name] = rec[:, o:o + k].unflatten(1, shp) if k else rec[:, o:o].reshape((rec.shape[0],) + tuple(shp))
        return out

    def in_views(self, rec):
        return self._views(rec, self.in_off, self.in_shapes, self.IN_FIELDS)

    def out_views(self, rec):
        v = self._views(rec, self.out_off, self.out_shapes, self.OUT_FIELDS)
        v.pop('status')                                    # one float64 slot holding two int32: info, sweeps
        ints = rec.view(torch.int32)                       # [B, 2 w_out]
        v['info'], v['sweeps'] = ints[:, 2 * self.out_off[5]], ints[:, 2 * self.out_off[5] + 1]
        return v

    def pack_inputs(self, rec, coords, props, nodal_mass=None, f_nodal=None, r_local=None):
        """Fill the input records from dense tensors (gathers the load entries of the pattern)."""
        v = self.in_views(rec)
        v['coords'].copy_(coords)
        v['props'].copy_(props)
        if nodal_mass is None:
            v['nodal_mass'].zero_()
        else:
            v['nodal_mass'].copy_(nodal_mass)
        B = rec.shape[0]
        if self.f_index.size:
            idx = torch.as_tensor(self.f_index, dtype=torch.long, device=f_nodal.device)
            v['f_values'].copy_(f_nodal.reshape(B, -1)[:, idx])
        if self.r_index.size:
            idx = torch.as_tensor(self.r_index, dtype=torch.long, device=r_local.device)
            v['r_values'].copy_(r_local.reshape(B, -1)[:, idx])
        return rec

    def expand(self, views):
        """Compact results -> the dense layouts of solve_fused: u / phi on all DOFs (zeros at supports), reactions
        on all DOFs (NaN-free zeros away from the supports).  Works on host or device tensors."""
        pos = torch.as_tensor(self.plan.pos_of_free.astype('int64'), device=views['info'].device)
        fixed = torch.as_tensor(self.plan.fixed.astype('int64'), device=views['info'].device)
        B, sd = views['info'].shape[0], self.plan.sumdof
        out = dict(endforces=views['endforces'], lam=views['lam'], info=views['info'], sweeps=views['sweeps'])
        if self.C > 0:
            u = torch.zeros((B, self.C, sd), dtype=torch.float64, device=pos.device)
            u[:, :, pos] = views['u']
            re = torch.zeros((B, self.C, sd), dtype=torch.float64, device=pos.device)
            re[:, :, fixed] = views['reactions']
            out.update(u=u, reactions=re)
        if self.n_modes > 0:
            phi = torch.zeros((B, self.n_modes, sd), dtype=torch.float64, device=pos.device)
            phi[:, :, pos] = views['phi']
            out['phi'] = phi
        return out

    def solve(self, in_rec, out_rec=None):
        """Run the fused kernel on device records in_rec [B, w_in] -> out_rec [B, w_out]."""
        B = in_rec.shape[0]
        in_rec = _chk(in_rec, (B, self.w_in), 'in_rec')
        if out_rec is None:
            out_rec = torch.empty((B, self.w_out), dtype=torch.float64, device=in_rec.device)
        elif not out_rec.is_contiguous():
            raise ValueError('out_rec must be contiguous')
        out_rec = _chk(out_rec, (B, self.w_out), 'out_rec')
        with torch.cuda.device(in_rec.device):
            _lib.check(_lib.lib().bfe2_solve_fused_packed(self.plan.handle, self._h, B, _lib.ptr(in_rec),
                                                          _lib.ptr(out_rec), _stream_ptr()))
        return out_rec


class HostPipeline:
    """End-to-end path with HOST buffers: pinned host records -> H2D -> fused kernel -> D2H, chunked over a few
    CUDA streams so that copies overlap the kernels.  This is the call a user of the reference-facing API makes
    when the model data lives on the host (the reference keeps everything in host numpy matrices,
    BeamFE2/Structure.py).  One packed record per structure and direction (PackedIO): ONE cudaMemcpyAsync each way
    per chunk, and only the non-zero load entries, the free-DOF displacements / shapes and the support reactions
    cross the bus (P20: 1.7 KB in + 9.4 KB out per structure instead of 5.7 + 11.4 KB dense)."""

    def __init__(self, plan: Plan, n_loadcases: int, n_modes: int, chunk: int = 4096, n_streams: int = 3,
                 device=None, modal: bool = True, f_index=None, r_index=None, io: PackedIO = None):
        self.io = io if io is not None else PackedIO(plan, n_loadcases, n_modes, modal, f_index, r_index)
        self.plan, self.C, self.modal, self.n_modes = plan, self.io.C, self.io.modal, self.io.n_modes
        self.chunk = int(chunk)
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(n_streams)]
        self.stage = [(torch.empty((self.chunk, self.io.w_in), dtype=torch.float64, device=self.device),
                       torch.empty((self.chunk, self.io.w_out), dtype=torch.float64, device=self.device))
                      for _ in range(n_streams)]
        self.launches = 0
        self.copies = 0

    def alloc_host(self, B: int):
        """Pinned host records for a batch of B: (in_rec [B, w_in], out_rec [B, w_out]).  Field views:
        self.io.in_views(in_rec) / self.io.out_views(out_rec); dense layouts: self.io.expand(views)."""
        return (torch.zeros((B, self.io.w_in), dtype=torch.float64).pin_memory(),
                torch.empty((B, self.io.w_out), dtype=torch.float64).pin_memory())

    def bytes_per_structure(self):
        return 8 * self.io.w_in, 8 * self.io.w_out

    def run(self, h_in, h_out):
        """Enqueue the whole batch; returns after everything has landed in h_out."""
        B = h_in.shape[0]
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            s.wait_stream(cur)
        for ci, lo in enumerate(range(0, B, self.chunk)):
            hi = min(B, lo + self.chunk)
            m = hi - lo
            st = self.streams[ci % len(self.streams)]
            d_in, d_out = self.stage[ci % len(self.streams)]
            with torch.cuda.stream(st):
                d_in[:m].copy_(h_in[lo:hi], non_blocking=True)
                self.io.solve(d_in[:m], d_out[:m])
                h_out[lo:hi].copy_(d_out[:m], non_blocking=True)
                self.launches += 1
                self.copies += 2
        for s in self.streams:
            cur.wait_stream(s)
        cur.synchronize()
        return h_out
