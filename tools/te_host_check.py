#!/usr/bin/env python
"""DEVELOPMENT TOOL (no GPU needed): the two-ended modal engine (beamfe2_b200/csrc/bfe2_te.cuh) run on the CPU through
tools/te_host_check.cu against the 40-digit truth fixture (64 P20 variants, tests/golden/p20_truth.npz)."""
import ctypes as C
import os
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from beamfe2_b200 import sweep                  # noqa: E402
from oracle import beamfe2_oracle as orc        # noqa: E402
from tests import util                          # noqa: E402

SO = '/tmp/libte_host.so'


def build():
    subprocess.check_call(['g++', '-std=c++20', '-O2', '-pthread', '-fPIC', '-shared', '-DBFE2_TE_HOST', '-ffp-contract=off',
                           '-o', SO, '-x', 'c++', os.path.join(ROOT, 'tools', 'te_host_check.cu')])


def to_band(A, b):
    n = A.shape[0]
    out = np.zeros((b + 1, n))
    for d in range(b + 1):
        out[d, :n - d] = np.diagonal(A, -d)
    return out


def main():
    build()
    L = C.CDLL(SO)
    t = util.p20_truth()
    P = t['params']
    pk = {k: v.numpy() for k, v in sweep.p20_pack(torch.from_numpy(P)).items()}
    conn, sup = sweep.p20_topology()
    fixed = orc.positions_to_eliminate(sup)
    o = orc.solve_batch(conn, fixed, pk['coords'], pk['props'], pk['nodal_mass'], pk['f_nodal'], pk['r_local'], n_modes=1)
    keep = orc.free_positions(21, fixed)
    n, b, nm = 57, 5, 10
    wl = wv = 0.0
    t0 = time.time()
    for v in range(P.shape[0]):
        Kb = np.ascontiguousarray(to_band(o['Kc'][v], b))
        Mb = np.ascontiguousarray(to_band(o['Mc'][v], b))
        lam = np.zeros(n)
        phi = np.zeros((nm, n))
        rc = L.te_modal(n, b, Kb.ctypes.data_as(C.c_void_p), Mb.ctypes.data_as(C.c_void_p), nm,
                        lam.ctypes.data_as(C.c_void_p), phi.ctypes.data_as(C.c_void_p))
        el = np.abs(lam / t['lam'][v] - 1)
        full = np.zeros((nm, 63))
        full[:, keep] = phi
        ev = util.mode_err_clustered(full, t['phi'][v], t['lam'][v])
        G = phi @ o['Mc'][v] @ phi.T
        og = np.abs(G - np.eye(nm)).max()
        wl, wv = max(wl, el.max()), max(wv, np.nanmax(ev))
        if rc != 0 or el.max() > 2e-11 or np.nanmax(ev) > 2e-11 or og > 1e-11:
            print(v, t['kind'][v], 'rc', rc, 'lam err %.2e (index %d)' % (el.max(), el.argmax()), 'vec err %.2e' % np.nanmax(ev),
                  'orth %.1e' % og)
    print('worst lam %.2e  worst vec %.2e  (%.1f s)' % (wl, wv, time.time() - t0))


if __name__ == '__main__':
    main()
