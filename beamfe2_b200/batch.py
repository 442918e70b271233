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


def internal_actions(plan: Plan, coords, endforces, q_axial=None, q_perp=None, npts=11, want_field=True):
    """N, V, M along every beam: actions [B,C,nE,3,npts] (or None) and maxabs [B,C,nE,3]."""
    B, C = endforces.shape[0], endforces.shape[1]
    dev = endforces.device
    coords = _chk(coords, (B, plan.n_nodes, 2), 'coords')
    endforces = _chk(endforces, (B, C, plan.n_elem, 6), 'endforces')
    q_axial = _chk(q_axial, (B, C, plan.n_elem), 'q_axial')
    q_perp = _chk(q_perp, (B, C, plan.n_elem), 'q_perp')
    act = torch.empty((B, C, plan.n_elem, 3, npts), dtype=torch.float64, device=dev) if want_field else None
    mx = torch.empty((B, C, plan.n_elem, 3), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().bfe2_internal_actions(plan.handle, B, C, int(npts), _lib.ptr(coords), _lib.ptr(endforces),
                                                    _lib.ptr(q_axial), _lib.ptr(q_perp), _lib.ptr(act), _lib.ptr(mx),
                                                    _stream_ptr()))
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


class HostPipeline:
    """End-to-end path with HOST buffers: pinned host tensors -> H2D -> fused kernel -> D2H, chunked
    over a few CUDA streams so that copies overlap the kernels.  This is the call a user of the
    reference-facing API makes when the model data lives on the host (the reference keeps everything
    in host numpy matrices, BeamFE2/Structure.py)."""

    IN_KEYS = ('coords', 'props', 'nodal_mass', 'f_nodal', 'r_local')
    OUT_KEYS = ('u', 'endforces', 'reactions', 'lam', 'phi', 'info', 'sweeps')

    def __init__(self, plan: Plan, n_loadcases: int, n_modes: int, chunk: int = 4096, n_streams: int = 3,
                 device=None, modal: bool = True):
        self.plan, self.C, self.modal = plan, int(n_loadcases), bool(modal)
        self.n_modes = int(n_modes) if modal else 0
        self.chunk = int(chunk)
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        p, C = plan, self.C
        self.in_shapes = dict(coords=(p.n_nodes, 2), props=(p.n_elem, 4), nodal_mass=(p.n_nodes, 2),
                              f_nodal=(C, p.sumdof), r_local=(C, p.n_elem, 6))
        self.out_shapes = dict(u=(C, p.sumdof), endforces=(C, p.n_elem, 6), reactions=(C, p.sumdof),
                               lam=(p.n_free,), phi=(self.n_modes, p.sumdof), info=(), sweeps=())
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(n_streams)]
        self.stage = []
        for _ in range(n_streams):
            d_in = {k: torch.empty((self.chunk,) + s, dtype=torch.float64, device=self.device)
                    for k, s in self.in_shapes.items()}
            d_out = {k: torch.empty((self.chunk,) + s, device=self.device,
                                    dtype=torch.int32 if k in ('info', 'sweeps') else torch.float64)
                     for k, s in self.out_shapes.items()}
            self.stage.append((d_in, d_out))
        self.launches = 0

    def alloc_host(self, B: int):
        """Pinned host input and output buffers for a batch of B."""
        h_in = {k: torch.empty((B,) + s, dtype=torch.float64).pin_memory() for k, s in self.in_shapes.items()}
        h_out = {k: torch.empty((B,) + s, dtype=torch.int32 if k in ('info', 'sweeps') else torch.float64).pin_memory()
                 for k, s in self.out_shapes.items()}
        return h_in, h_out

    def bytes_per_structure(self):
        import math
        bi = sum(8 * math.prod(s) for s in self.in_shapes.values())
        bo = sum((4 if k in ('info', 'sweeps') else 8) * math.prod(s) for k, s in self.out_shapes.items())
        return bi, bo

    def run(self, h_in: dict, h_out: dict):
        """Enqueue the whole batch; returns after everything has landed in h_out."""
        B = h_in['coords'].shape[0]
        use_static = self.C > 0
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            s.wait_stream(cur)
        for ci, lo in enumerate(range(0, B, self.chunk)):
            hi = min(B, lo + self.chunk)
            m = hi - lo
            st = self.streams[ci % len(self.streams)]
            d_in, d_out = self.stage[ci % len(self.streams)]
            with torch.cuda.stream(st):
                for k in self.IN_KEYS:
                    if not use_static and k in ('f_nodal', 'r_local'):
                        continue
                    d_in[k][:m].copy_(h_in[k][lo:hi], non_blocking=True)
                solve_fused(self.plan, d_in['coords'][:m], d_in['props'][:m], d_in['nodal_mass'][:m],
                            d_in['f_nodal'][:m] if use_static else None,
                            d_in['r_local'][:m] if use_static else None,
                            modal=self.modal, n_modes=self.n_modes,
                            out={k: v[:m] for k, v in d_out.items()})
                self.launches += 1
                for k in self.OUT_KEYS:
                    if d_out[k].numel() == 0:
                        continue
                    if not use_static and k in ('u', 'endforces', 'reactions'):
                        continue
                    if not self.modal and k in ('lam', 'phi'):
                        continue
                    h_out[k][lo:hi].copy_(d_out[k][:m], non_blocking=True)
        for s in self.streams:
            cur.wait_stream(s)
        cur.synchronize()
        return h_out
