"""ctypes stub over include/bfe2.h -- what a maintainer of jkbgbr/BeamFE2 would drop into the package as
`BeamFE2/_bfe2.py` to run its two solvers on libbfe2.so (option B).

It reads a *reference* `Structure` through the reference's own attributes only (file:line = BeamFE2/...):
nodes (Structure.py:154-157), beams and their i/j node IDs (helpers.py:28-31), positions_to_eliminate
(Structure.py:164-181), mass_matrix flags (HermitianBeam_2D.py:251-259), _mass_vector (Structure.py:289-313),
nodal_loads (Structure.py:76-97) and each beam's reduced_internal_loads_asvector (HermitianBeam_2D.py:149-162).

`pack_structure` is pure numpy (checked on the CPU against the reference's own load vector);
`solve_structure` needs a CUDA device: there is no CPU fallback.
"""
import ctypes as C
import os

import numpy as np

_LIB = os.environ.get('BFE2_LIB') or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                  'beamfe2_b200', 'libbfe2.so')
_L = None


def _lib():
    global _L
    if _L is None:
        L = C.CDLL(_LIB)
        vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
        L.bfe2_last_error.restype = C.c_char_p
        L.bfe2_plan_create.argtypes = [C.POINTER(vp), i32, i32, vp, i32, vp, vp]
        L.bfe2_plan_destroy.argtypes = [vp]
        L.bfe2_plan_destroy.restype = None
        L.bfe2_solve_fused.argtypes = [vp, i64, i32, i32, i32] + [vp] * 13
        _L = L
    return _L


def _chk(rc):
    if rc < 0:
        raise RuntimeError(_lib().bfe2_last_error().decode())


def pack_structure(structure):
    """reference Structure -> the arrays of include/bfe2.h (batch of one, one load case)."""
    nodes = sorted(structure.nodes, key=lambda n: n.ID)
    nN, nE = len(nodes), len(structure.beams)
    conn = np.array([[b.i.ID - 1, b.j.ID - 1] for b in structure.beams], dtype=np.int32)
    fixed = np.array(structure.positions_to_eliminate, dtype=np.int32)
    lumped = np.array([b.mass_matrix == 'lumped' for b in structure.beams], dtype=np.uint8)
    coords = np.array([[n.x, n.y] for n in nodes], dtype=np.float64)
    props = np.array([[b.E, b.A, b.I, 0.0 if b.rho is None else b.rho] for b in structure.beams], dtype=np.float64)   # rho is optional (:42-67)
    mv = np.zeros(3 * nN) if structure._mass_vector is None else np.asarray(structure._mass_vector, dtype=np.float64).ravel()
    mass = np.stack([mv[0::3], mv[1::3]], axis=1)
    r_local = np.zeros((nE, 6))
    for e, b in enumerate(structure.beams):
        r_local[e] = np.asarray(b.reduced_internal_loads_asvector, dtype=np.float64).ravel()
    # the reference's load vector = directly applied nodal dynams + T(alpha) r_local of every beam load
    # (Structure.py:244-287); libbfe2 takes the two parts separately (it needs r_local again for the end forces)
    f_nodal = np.zeros(3 * nN)
    if structure._load_vector is not None:
        f_nodal = np.asarray(structure._load_vector, dtype=np.float64).ravel().copy()
        for e, b in enumerate(structure.beams):
            if not np.any(r_local[e]):
                continue
            a = b.direction
            c, s = np.cos(a), np.sin(a)
            R = np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
            f_nodal[3 * conn[e, 0]:3 * conn[e, 0] + 3] -= R @ r_local[e, :3]
            f_nodal[3 * conn[e, 1]:3 * conn[e, 1] + 3] -= R @ r_local[e, 3:]
    return dict(conn=conn, fixed=fixed, lumped=lumped, coords=coords, props=props, nodal_mass=mass,
                f_nodal=f_nodal, r_local=r_local)


def solve_structure(structure, static=True, modal=False):
    """One reference Structure -> libbfe2 (batch of 1).  Returns dict(u, endforces, reactions, lam, phi, info)."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError('libbfe2 needs a CUDA device (no CPU fallback)')
    L = _lib()
    pk = pack_structure(structure)
    nN, nE = pk['coords'].shape[0], pk['props'].shape[0]
    n = 3 * nN - len(pk['fixed'])
    plan = C.c_void_p()
    _chk(L.bfe2_plan_create(C.byref(plan), nN, nE, pk['conn'].ctypes.data, len(pk['fixed']),
                            pk['fixed'].ctypes.data if len(pk['fixed']) else None, pk['lumped'].ctypes.data))
    try:
        def dev(a):
            return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device='cuda')
        coords, props, mass = dev(pk['coords'][None]), dev(pk['props'][None]), dev(pk['nodal_mass'][None])
        f_nodal, r_local = dev(pk['f_nodal'][None, None]), dev(pk['r_local'][None, None])
        Cn, nm = (1 if static else 0), (n if modal else 0)
        u = torch.empty((1, Cn, 3 * nN), dtype=torch.float64, device='cuda')
        ef = torch.empty((1, Cn, nE, 6), dtype=torch.float64, device='cuda')
        re = torch.empty((1, Cn, 3 * nN), dtype=torch.float64, device='cuda')
        lam = torch.empty((1, n), dtype=torch.float64, device='cuda')
        phi = torch.empty((1, nm, 3 * nN), dtype=torch.float64, device='cuda')
        info = torch.zeros(1, dtype=torch.int32, device='cuda')

        def p(t, on=True):
            return C.c_void_p(t.data_ptr()) if (on and t.numel()) else None
        _chk(L.bfe2_solve_fused(plan, 1, Cn, int(modal), nm, p(coords), p(props), p(mass), p(f_nodal, static),
                                p(r_local, static), p(u), p(ef), p(re), p(lam, modal), p(phi), p(info), None,
                                C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        torch.cuda.synchronize()
    finally:
        L.bfe2_plan_destroy(plan)
    return dict(u=u[0].cpu().numpy(), endforces=ef[0].cpu().numpy(), reactions=re[0].cpu().numpy(),
                lam=lam[0].cpu().numpy(), phi=phi[0].cpu().numpy(), info=int(info.item()))
