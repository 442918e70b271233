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
