#!/usr/bin/env python
"""DEVELOPMENT TOOL (no GPU needed): runs the config-4 arithmetic (bfe2_beam_core.cuh, the code the CUDA kernels
instantiate) in host loops through tools/beam_host_check.cu and drives the SAME subspace-iteration logic as
beamfe2_b200/beam.py with torch CPU tensors (Gram / update / Rayleigh-Ritz products by torch / scipy instead of the
DMMA and Jacobi kernels).  Purpose: check the double-double formulation at n = 1e5 before spending GPU time.
Nothing in the product imports this."""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import time

import numpy as np
import scipy.linalg as sla
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from beamfe2_b200 import sweep                       # noqa: E402
from beamfe2_b200.beam import SubspaceIteration      # noqa: E402
from oracle import beam_oracle as bo                 # noqa: E402

SO = '/tmp/libbeam_host.so'


def build():
    src = os.path.join(ROOT, 'tools', 'beam_host_check.cu')
    if os.path.exists(SO) and os.path.getmtime(SO) > max(os.path.getmtime(src), os.path.getmtime(
            os.path.join(ROOT, 'beamfe2_b200', 'csrc', 'bfe2_beam_core.cuh')), os.path.getmtime(
            os.path.join(ROOT, 'beamfe2_b200', 'csrc', 'bfe2_dd.cuh'))):
        return
    subprocess.check_call(['nvcc', '-O2', '-std=c++17', '-Xcompiler', '-fPIC,-fopenmp,-ffp-contract=off,-Wno-unknown-pragmas',
                           '-shared', '-o', SO, src])


class HostBeam(SubspaceIteration):
    def __init__(self, coords, props, fixed_positions, prec=1, n_partitions=0):
        build()
        self.L = C.CDLL(SO)
        self.L.hb_create.restype = C.c_void_p
        self.prec = prec
        self.coords = np.ascontiguousarray(coords)
        self.props = np.ascontiguousarray(props)
        nE = props.shape[0]
        self.n = 3 * (nE + 1)
        mask = np.zeros(self.n, dtype=np.uint8)
        mask[list(fixed_positions)] = 1
        self.fixed_mask = mask
        self.device = torch.device('cpu')
        info = C.c_int(0)
        self.h = C.c_void_p(self.L.hb_create(prec, C.c_longlong(nE), self.coords.ctypes.data_as(C.c_void_p),
                                             self.props.ctypes.data_as(C.c_void_p), mask.ctypes.data_as(C.c_void_p),
                                             n_partitions, C.byref(info)))
        assert info.value == 0, info.value
        self.n_partitions = self.L.hb_partitions(prec, self.h)

    def solve(self, F, out=None, want_lo=False):
        vec = F.dim() == 1
        F2 = F.reshape(self.n, -1).contiguous()
        U = torch.empty_like(F2) if out is None else out
        Ulo = torch.zeros_like(F2)
        self.L.hb_solve(self.prec, self.h, F2.shape[1], C.c_void_p(F2.data_ptr()), C.c_void_p(U.data_ptr()),
                        C.c_void_p(Ulo.data_ptr()))
        if want_lo:
            return U, Ulo
        return U.reshape(self.n) if vec else U

    def matvec(self, X, mass=False, out=None, lo=None, F=None):
        vec = X.dim() == 1
        X2 = X.reshape(self.n, -1).contiguous()
        Y = torch.empty_like(X2) if out is None else out
        self.L.hb_matvec(self.prec, self.h, 1 if mass else 0, X2.shape[1], C.c_void_p(X2.data_ptr()),
                         C.c_void_p(lo.data_ptr()) if lo is not None else None,
                         C.c_void_p(F.data_ptr()) if F is not None else None, C.c_void_p(Y.data_ptr()))
        return Y.reshape(self.n) if vec else Y

    def gram(self, X, Y):
        return X.t() @ Y

    def update(self, Z, Qm, out=None):
        r = Z @ Qm
        if out is not None:
            out.copy_(r)
            return out
        return r

    def small_eig(self, A, B):
        try:
            w, v = sla.eigh(A.numpy(), B.numpy())
            return torch.from_numpy(w), torch.from_numpy(v.T.copy()), torch.zeros(1, dtype=torch.int32)
        except Exception:
            return torch.zeros(A.shape[0], dtype=torch.float64), torch.eye(A.shape[0], dtype=torch.float64), \
                torch.ones(1, dtype=torch.int32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, default=100000)
    ap.add_argument('--prec', type=int, default=1)
    ap.add_argument('--modes', type=int, default=50)
    ap.add_argument('--parts', type=int, default=0)
    ap.add_argument('--no-modal', action='store_true')
    a = ap.parse_args()
    pk = sweep.beam_pack(a.n)
    coords, props = pk['coords'][0].numpy(), pk['props'][0].numpy()
    E, A, I, rho = props[0]
    Ltot = float(coords[-1, 0] - coords[0, 0])
    d_cf, r_cf, _ = bo.cantilever_closed_form(100.0, Ltot, E, I, rho, A, 1)
    t0 = time.time()
    hb = HostBeam(coords, props, [0, 1, 2], a.prec, a.parts)
    out = {'n': a.n, 'prec': a.prec, 'partitions': hb.n_partitions, 'create_s': time.time() - t0}
    f = pk['f_nodal'][0, 0].clone()
    t0 = time.time()
    u, ulo = hb.solve(f.reshape(-1, 1), want_lo=True)
    out['solve_s'] = time.time() - t0
    u = u.reshape(-1)
    r_hi = hb.matvec(u, F=f)
    r_dd = hb.matvec(u, F=f, lo=ulo.reshape(-1))
    kmax = float(hb.matvec(torch.ones_like(u)).abs().max())
    out['static'] = {'tip': float(u[-2]), 'closed': d_cf, 'rel_err_tip': abs(float(u[-2]) / d_cf - 1),
                     'rel_err_rot': abs(float(u[-1]) / r_cf - 1),
                     'scaled_res_hi': float(r_hi.abs().max()) / (kmax * float(u.abs().max())),
                     'scaled_res_dd': float(r_dd.abs().max()) / (kmax * float(u.abs().max()))}
    print(json.dumps(out, indent=1), flush=True)
    if a.no_modal:
        return
    t0 = time.time()
    mo = hb.modal(n_modes=a.modes, tol=1e-10, maxit=40, verbose=True)
    om = np.sqrt(np.maximum(mo['lam'].numpy(), 0))
    exact = bo.cantilever_spectrum(Ltot, E, I, rho, A, a.modes)
    out['modal'] = {'seconds': time.time() - t0, 'iterations': mo['iterations'], 'converged': mo['converged'],
                    'failed': mo['failed'], 'omega_first8': om[:8].tolist(), 'exact_first8': exact[:8].tolist(),
                    'max_rel_err_all': float(np.abs(om / exact - 1).max()),
                    'rel_err_first5': np.abs(om[:5] / exact[:5] - 1).tolist(),
                    'max_scaled_residual': float(mo['residual'].max())}
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
