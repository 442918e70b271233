"""CPU oracle for the BeamFE2 hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy/scipy restatement of the reference's algorithm for the path
element K/M -> rotation -> global scatter-add -> supports -> static solve -> modal eigenproblem
-> end forces / reactions, vectorised over a leading batch axis.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module, and only as the checker / reported baseline.  The product path
(beamfe2_b200) never imports it and has no CPU fallback.

Parity status: PINNED.  tests/test_oracle_pinning.py checks this restatement against golden
vectors produced by the real reference (imported from /root/reference by oracle/make_golden.py
in the build container; the vectors are committed under tests/golden/) and against the
reference's own known-answer tests (cited per function below).

All citations are file:line relative to the reference repo root.

Packed model format shared with the product:
    conn      int   [nE, 2]    0-based node indices (i, j) of each beam, in beam-list order
    fixed     int   [nF]       sorted supported positions == Structure.positions_to_eliminate
    lumped    bool  [nE]       per-beam mass_matrix == 'lumped'
    coords    f64   [..., nN, 2]
    props     f64   [..., nE, 4]   (E, A, I, rho)
    nodal_mass f64  [..., nN, 2]   (mx, my) summed per node
    f_nodal   f64   [..., C, 3 nN] directly applied nodal loads per load case
    r_local   f64   [..., C, nE, 6] sum of the local fixed-end reaction vectors per beam
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla

DOF = 3
DOFNAMES = ('ux', 'uy', 'rotz')        # Structure.py:17
LOADNAMES = ('FX', 'FY', 'MZ')         # Structure.py:16
MASSNAMES = ('mx', 'my')               # Structure.py:18
VERY_LARGE_NUMBER = 10e40              # Structure.py:11

BEAM_LOAD_KINDS = ('uniform axial force', 'uniform perpendicular force',
                   'concentrated perpendicular force', 'concentrated axial force',
                   'concentrated moment')


# --------------------------------------------------------------------------- INT maps
def position_in_matrix(node_id: int, name: str) -> int:
    """Structure.py:379-393 -- (nodeID-1)*3 + index of the DOF (or load) name.  node_id is 1-based."""
    if name in DOFNAMES:
        return (node_id - 1) * DOF + DOFNAMES.index(name)
    if name in LOADNAMES:
        return (node_id - 1) * DOF + LOADNAMES.index(name)
    raise AssertionError(name)


def nodenr_dof_from_position(position: int):
    """Structure.py:370-377."""
    n, d = divmod(position, DOF)
    return n + 1, LOADNAMES[d]


def positions_to_eliminate(supports: dict) -> list:
    """Structure.py:164-181 -- sorted supported positions from {nodeID: [dofnames]}."""
    out = []
    for node_id, dofs in supports.items():
        for d in dofs:
            out.append(position_in_matrix(node_id, d))
    out.sort()
    return out


def block_starts(conn: np.ndarray) -> np.ndarray:
    """helpers.py:28-31 -- (_sti, _stj) = ((i.ID-1)*3, (j.ID-1)*3); conn is already 0-based."""
    return np.asarray(conn, dtype=np.int64) * DOF


def free_positions(n_nodes: int, fixed) -> np.ndarray:
    """Complement of positions_to_eliminate in ascending order (what np.delete leaves behind,
    Structure.py:183-191)."""
    mask = np.ones(n_nodes * DOF, dtype=bool)
    mask[np.asarray(fixed, dtype=np.int64)] = False
    return np.nonzero(mask)[0]


# --------------------------------------------------------------------------- element level
def length_direction(coords, conn):
    """HermitianBeam_2D.py:98-100 (l) and :114-120 (direction = arctan2(dy, dx))."""
    conn = np.asarray(conn)
    pi = coords[..., conn[:, 0], :]
    pj = coords[..., conn[:, 1], :]
    dx = pj[..., 0] - pi[..., 0]
    dy = pj[..., 1] - pi[..., 1]
    L = np.sqrt((pi[..., 0] - pj[..., 0]) ** 2 + (pi[..., 1] - pj[..., 1]) ** 2)
    alpha = np.arctan2(dy, dx)
    return L, alpha


def ke_local(E, A, I, L):
    """HermitianBeam_2D.py:284-297 -- local 6x6 stiffness, DOF order [ux_i,uy_i,rz_i,ux_j,uy_j,rz_j]."""
    EA = E * A
    EI = E * I
    k = np.zeros(np.shape(L) + (6, 6))
    k[..., 0, 0] = k[..., 3, 3] = EA / L
    k[..., 1, 1] = k[..., 4, 4] = 12 * EI / (L ** 3)
    v = 6 * EI / (L ** 2)
    k[..., 1, 2] = k[..., 2, 1] = k[..., 1, 5] = k[..., 5, 1] = v
    k[..., 2, 2] = k[..., 5, 5] = 4 * EI / L
    k[..., 2, 5] = k[..., 5, 2] = 2 * EI / L
    k[..., 0, 3] = k[..., 3, 0] = -1 * k[..., 0, 0]
    k[..., 1, 4] = k[..., 4, 1] = -1 * k[..., 1, 1]
    k[..., 2, 4] = k[..., 4, 2] = k[..., 4, 5] = k[..., 5, 4] = -1 * v
    return k


def me_local(rho, A, L, lumped=None):
    """HermitianBeam_2D.py:233-249 (consistent) and :225-231 (lumped, rotary = m*1e-10)."""
    L = np.asarray(L, dtype=np.float64)
    m = np.zeros(L.shape + (6, 6))
    m[..., 0, 0] = m[..., 3, 3] = 140
    m[..., 0, 3] = m[..., 3, 0] = 70
    m[..., 1, 1] = m[..., 4, 4] = 156
    m[..., 1, 2] = m[..., 2, 1] = 22 * L
    m[..., 1, 4] = m[..., 4, 1] = 54
    m[..., 1, 5] = m[..., 5, 1] = -13 * L
    m[..., 2, 2] = m[..., 5, 5] = 4 * L ** 2
    m[..., 2, 4] = m[..., 4, 2] = -1 * m[..., 1, 5]
    m[..., 2, 5] = m[..., 5, 2] = -3 * L ** 2
    m[..., 4, 5] = m[..., 5, 4] = -1 * m[..., 1, 2]
    m = (rho * A * L / 420)[..., None, None] * m
    if lumped is not None and np.any(lumped):
        ml = np.zeros(L.shape + (6, 6))
        mm = 0.5 * A * L * rho
        for d in (0, 1, 3, 4):
            ml[..., d, d] = mm
        ml[..., 2, 2] = ml[..., 5, 5] = mm * 1e-10
        sel = np.broadcast_to(np.asarray(lumped, dtype=bool), L.shape)[..., None, None]
        m = np.where(sel, ml, m)
    return m


def rotation3(alpha):
    """helpers.py:43-69 with blocks=1 -- [[c,-s,0],[s,c,0],[0,0,1]], c=cos(alpha), s=sin(alpha)."""
    alpha = np.asarray(alpha, dtype=np.float64)
    c = np.cos(alpha)
    s = np.sin(alpha)
    R = np.zeros(alpha.shape + (3, 3))
    R[..., 0, 0] = c
    R[..., 0, 1] = -s
    R[..., 1, 0] = s
    R[..., 1, 1] = c
    R[..., 2, 2] = 1
    return R


def transfer_matrix(alpha):
    """helpers.py:43-69 with blocks=2 -- T = diag(R, R)."""
    R = rotation3(alpha)
    T = np.zeros(R.shape[:-2] + (6, 6))
    T[..., :3, :3] = R
    T[..., 3:, 3:] = R
    return T


def matrix_in_global(T, X):
    """HermitianBeam_2D.py:173-175 -- T * X * T^T."""
    return T @ X @ np.swapaxes(T, -1, -2)


def element_km_global(coords, conn, props, lumped=None):
    """Per-element global 6x6 K and M (what compile_global_matrix feeds on, helpers.py:19-23)."""
    L, alpha = length_direction(coords, conn)
    E, A, I, rho = (props[..., k] for k in range(4))
    T = transfer_matrix(alpha)
    Kg = matrix_in_global(T, ke_local(E, A, I, L))
    Mg = matrix_in_global(T, me_local(rho, A, L, lumped))
    return Kg, Mg


# --------------------------------------------------------------------------- assembly
def compile_global_matrix(Xg, conn, n_nodes):
    """helpers.py:7-40 -- dense scatter-add of the four 3x3 blocks per beam, beams in list order."""
    Xg = np.asarray(Xg)
    n = n_nodes * DOF
    G = np.zeros(Xg.shape[:-3] + (n, n))
    st = block_starts(conn)
    for e in range(len(conn)):
        si, sj = int(st[e, 0]), int(st[e, 1])
        m = Xg[..., e, :, :]
        G[..., si:si + 3, si:si + 3] += m[..., :3, :3]
        G[..., sj:sj + 3, si:si + 3] += m[..., 3:, :3]
        G[..., si:si + 3, sj:sj + 3] += m[..., :3, 3:]
        G[..., sj:sj + 3, sj:sj + 3] += m[..., 3:, 3:]
    return G


def k_with_bc(K, fixed):
    """Structure.py:221-231 -- penalty 1e41 on every supported diagonal."""
    K = np.array(K, copy=True)
    for p in fixed:
        K[..., p, p] += VERY_LARGE_NUMBER
    return K


def m_with_masses(M, nodal_mass):
    """Structure.py:198-209, 289-313 -- M[k,k] += mass_vector[k]; mx->ux, my->uy, no rotary term."""
    M = np.array(M, copy=True)
    nN = nodal_mass.shape[-2]
    idx = np.arange(nN) * DOF
    M[..., idx, idx] += nodal_mass[..., :, 0]
    M[..., idx + 1, idx + 1] += nodal_mass[..., :, 1]
    return M


def condense(X, fixed):
    """Structure.py:183-191 -- delete supported rows and columns."""
    n = X.shape[-1]
    keep = free_positions(n // DOF, fixed)
    return X[..., keep[:, None], keep[None, :]]


# --------------------------------------------------------------------------- loads
def beam_load_reactions_local(kind, value, L, position=0.0):
    """Loads.py:202-208, 264-270, 333-342, 409-416, 474-484 -- local 1x6 fixed-end vectors."""
    L = np.asarray(L, dtype=np.float64)
    z = np.zeros_like(L)
    a = position
    b = 1 - position
    if kind == 'uniform axial force':
        r = [value * L / 2, z, z, value * L / 2, z, z]
    elif kind == 'uniform perpendicular force':
        r = [z, value * L / 2, value * (L ** 2) / 12, z, value * L / 2, -1 * value * (L ** 2) / 12.]
    elif kind == 'concentrated perpendicular force':
        r = [z, (3 - 2 * b) * (b ** 2) * value + z, a * (b ** 2) * value * L,
             z, (3 - 2 * a) * (a ** 2) * value + z, -1 * b * (a ** 2) * value * L]
    elif kind == 'concentrated axial force':
        r = [value / 2. + z, z, z, value / 2. + z, z, z]
    elif kind == 'concentrated moment':
        MR = 6 * a * b * value / L
        r = [z, -MR, 1 * (3 * b - 2) * b * value + z, z, MR, 1 * (3 * a - 2) * a * value + z]
    else:
        raise NotImplementedError(kind)
    return np.stack(r, axis=-1)


def load_vector(f_nodal, r_local, coords, conn):
    """Structure.py:244-287 + Loads.py:100-112 -- f = nodal loads + sum_e T(alpha_e) r_local_e,
    scattered to the beam's two nodes."""
    f = np.array(f_nodal, dtype=np.float64, copy=True)
    if r_local is None:
        return f
    _, alpha = length_direction(coords, conn)
    T = transfer_matrix(alpha)                       # [..., nE, 6, 6]
    if f.ndim == T.ndim - 1:                         # [..., C, n] vs [..., nE,6,6]: add the C axis
        T = T[..., None, :, :, :]
    rg = np.einsum('...ij,...j->...i', T, r_local)   # [..., C, nE, 6]
    st = block_starts(conn)
    for e in range(len(conn)):
        si, sj = int(st[e, 0]), int(st[e, 1])
        f[..., si:si + 3] += rg[..., e, :3]
        f[..., sj:sj + 3] += rg[..., e, 3:]
    return f


# --------------------------------------------------------------------------- solves
def linear_static_penalty(Kbc, f):
    """Solver.py:83 -- disps = inv(K_with_BC) * load_vector (dense inverse, penalty supports)."""
    return np.einsum('...ij,...cj->...ci', np.linalg.inv(Kbc), f)


def linear_static_condensed(K, f, fixed):
    """north_star comparison target: numpy.linalg.solve on the condensed system, zeros at supports."""
    n = K.shape[-1]
    keep = free_positions(n // DOF, fixed)
    Kc = K[..., keep[:, None], keep[None, :]]
    fc = f[..., keep]
    uc = np.linalg.solve(Kc, np.swapaxes(fc, -1, -2))
    u = np.zeros_like(f)
    u[..., keep] = np.swapaxes(uc, -1, -2)
    return u


def modal_eigh(Kc, Mc):
    """Solver.py:39 -- scipy.linalg.eigh(K, M): ascending eigenvalues, Phi^T M Phi = I.
    Loops over the batch (scipy.linalg.eigh has no batched generalized driver)."""
    Kc = np.asarray(Kc)
    Mc = np.asarray(Mc)
    if Kc.ndim == 2:
        return sla.eigh(Kc, Mc)
    lead = Kc.shape[:-2]
    n = Kc.shape[-1]
    Kf = Kc.reshape((-1, n, n))
    Mf = Mc.reshape((-1, n, n))
    lam = np.empty((Kf.shape[0], n))
    phi = np.empty((Kf.shape[0], n, n))
    for b in range(Kf.shape[0]):
        lam[b], phi[b] = sla.eigh(Kf[b], Mf[b])
    return lam.reshape(lead + (n,)), phi.reshape(lead + (n, n))


def expand_shapes(phi, n_nodes, fixed):
    """Solver.py:54-65 -- one column per eigenvector, zeros re-inserted at the supported positions.
    Returns [..., n_free(modes), sumdof]."""
    keep = free_positions(n_nodes, fixed)
    out = np.zeros(phi.shape[:-2] + (phi.shape[-1], n_nodes * DOF))
    out[..., :, keep] = np.swapaxes(phi, -1, -2)
    return out


def circular_frequencies(lam):
    """Solver.py:42 -- sqrt of the strictly positive eigenvalues only."""
    return [float(np.sqrt(x)) for x in np.asarray(lam).ravel() if x > 0]


# --------------------------------------------------------------------------- post-processing
def element_displacements_local(u, coords, conn):
    """Results.py:48-79 -- [R(-alpha) u_i ; R(-alpha) u_j] per beam.  u: [..., C, sumdof]."""
    _, alpha = length_direction(coords, conn)
    Rm = rotation3(-alpha)                            # [..., nE, 3, 3]
    if u.ndim == Rm.ndim - 1:
        Rm = Rm[..., None, :, :, :]
    st = block_starts(conn)
    ui = np.stack([u[..., st[e, 0]:st[e, 0] + 3] for e in range(len(conn))], axis=-2)
    uj = np.stack([u[..., st[e, 1]:st[e, 1] + 3] for e in range(len(conn))], axis=-2)
    li = np.einsum('...ij,...j->...i', Rm, ui)
    lj = np.einsum('...ij,...j->...i', Rm, uj)
    return np.concatenate([li, lj], axis=-1)          # [..., C, nE, 6]


def element_end_forces(u, coords, conn, props, r_local=None):
    """HermitianBeam_2D.py:336-351 -- Ke * u_local - sum of local fixed-end vectors."""
    L, _ = length_direction(coords, conn)
    E, A, I = props[..., 0], props[..., 1], props[..., 2]
    Ke = ke_local(E, A, I, L)
    ul = element_displacements_local(u, coords, conn)
    if ul.ndim == Ke.ndim:
        Ke = Ke[..., None, :, :, :]
    q = np.einsum('...ij,...j->...i', Ke, ul)
    if r_local is not None:
        q = q - r_local
    return q


def reaction_forces_asvector(endf, f_nodal, coords, conn):
    """Results.py:92-127 -- sum over beams of R(alpha) * end forces at each node, minus the
    directly applied nodal loads."""
    _, alpha = length_direction(coords, conn)
    R = rotation3(alpha)
    if endf.ndim == R.ndim:
        R = R[..., None, :, :, :]
    gi = np.einsum('...ij,...j->...i', R, endf[..., :3])
    gj = np.einsum('...ij,...j->...i', R, endf[..., 3:])
    out = np.zeros_like(f_nodal, dtype=np.float64)
    st = block_starts(conn)
    for e in range(len(conn)):
        out[..., st[e, 0]:st[e, 0] + 3] += gi[..., e, :]
        out[..., st[e, 1]:st[e, 1] + 3] += gj[..., e, :]
    return out - f_nodal


def internal_action_baseline(endf_e, action, pos):
    """HermitianBeam_2D.py:398-431 -- the end-force baseline only (v1, -v2 interpolation for the
    moment, -v1 for axial/shear); the load add-on curves are host-side queries."""
    ndx = ('axial', 'shear', 'moment').index(action)
    v1 = endf_e[..., ndx]
    v2 = endf_e[..., ndx + 3]
    if action == 'moment':
        return v1 + (-v2 - v1) * pos
    return -v1


def internal_action_field(endf, L, q_axial=None, q_perp=None, npts=11):
    """HermitianBeam_2D.py:398-431 + Loads.py:214-225, 277-289 -- N, V, M at npts equidistant positions for
    uniform beam loads.  endf [..., 6], L [...], q_* [...] -> [..., 3, npts]."""
    xi = np.linspace(0.0, 1.0, npts) if npts > 1 else np.zeros(1)
    L = np.asarray(L)[..., None]
    qa = 0.0 if q_axial is None else np.asarray(q_axial)[..., None]
    qp = 0.0 if q_perp is None else np.asarray(q_perp)[..., None]
    n1, v1, m1, m2 = (endf[..., k][..., None] for k in (0, 1, 2, 5))
    ax = -n1 + (-xi * qa * L)
    sh = -v1 + (xi * L * -qp)
    mo = (m1 + (-m2 - m1) * xi) + ((xi * (1 - xi)) * qp * L ** 2) / 2.
    return np.stack([ax + 0 * xi, sh + 0 * xi, mo], axis=-2)


def participation_factors(phi, M):
    """Examples/modal_masses.py:40-49 -- Gamma = phi^T M r / (phi^T M phi) for r = 1 on the ux (resp. uy) DOFs.
    phi [..., m, sumdof], M [..., sumdof, sumdof] -> [..., m, 2]."""
    n = M.shape[-1]
    out = []
    for k in (0, 1):
        r = np.zeros(n)
        r[k::3] = 1.0
        num = np.einsum('...mi,...ij,j->...m', phi, M, r)
        den = np.einsum('...mi,...ij,...mj->...m', phi, M, phi)
        out.append(num / den)
    return np.stack(out, axis=-1)


# --------------------------------------------------------------------------- whole path
def solve_batch(conn, fixed, coords, props, nodal_mass=None, f_nodal=None, r_local=None,
                lumped=None, n_modes=None, static='condensed'):
    """The whole hot path on packed inputs; returns a dict of numpy arrays.

    static='condensed' -> numpy.linalg.solve on free DOFs (north_star target);
    static='penalty'   -> the reference's inv(K + 1e41) route (Solver.py:83).
    """
    conn = np.asarray(conn)
    nN = coords.shape[-2]
    Kg, Mg = element_km_global(coords, conn, props, lumped)
    K = compile_global_matrix(Kg, conn, nN)
    out = {'K': K}
    if f_nodal is not None:
        f = load_vector(f_nodal, r_local, coords, conn)
        if static == 'penalty':
            u = linear_static_penalty(k_with_bc(K, fixed), f)
        else:
            u = linear_static_condensed(K, f, fixed)
        endf = element_end_forces(u, coords, conn, props, r_local)
        out.update(f=f, u=u, endforces=endf,
                   reactions=reaction_forces_asvector(endf, f_nodal, coords, conn))
    if nodal_mass is not None or n_modes is not None:
        M = compile_global_matrix(Mg, conn, nN)
        if nodal_mass is not None:
            M = m_with_masses(M, nodal_mass)
        Kc = condense(K, fixed)
        Mc = condense(M, fixed)
        lam, phi = modal_eigh(Kc, Mc)
        shapes = expand_shapes(phi, nN, fixed)
        if n_modes is not None:
            shapes = shapes[..., :n_modes, :]
        out.update(M=M, Kc=Kc, Mc=Mc, lam=lam, phi=shapes)
    return out
