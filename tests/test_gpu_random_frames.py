"""Randomised general frames (not chains): random connected topology on a jittered grid, random node numbering
(wide bands -> warp band-Cholesky and generic substitution paths), several beams per node, random supports,
nodal masses, nodal + uniform beam loads, n_free from 3 to 116 (both Jacobi variants, odd and even n).
Everything is compared with the CPU oracle (numpy.linalg.solve / scipy.linalg.eigh on the same inputs)."""
import numpy as np
import pytest
import torch

from beamfe2_b200 import batch
from beamfe2_b200.plan import Plan
from oracle import beamfe2_oracle as orc
from tests import util

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
EPS = np.finfo(float).eps


def random_frame(rng, n_nodes, extra_edges, n_clamped, n_pinned):
    side = int(np.ceil(np.sqrt(n_nodes)))
    pts = [(i, j) for i in range(side) for j in range(side)][:n_nodes]
    xy = np.array(pts, dtype=float) * 1000.0 + rng.uniform(-200, 200, (n_nodes, 2))
    perm = rng.permutation(n_nodes)                       # random numbering
    xy = xy[perm]
    edges = set()
    order = rng.permutation(n_nodes)
    for k in range(1, n_nodes):                           # random spanning tree: connected
        a, b = int(order[k]), int(order[rng.integers(0, k)])
        edges.add((min(a, b), max(a, b)))
    while len(edges) < n_nodes - 1 + extra_edges:
        a, b = (int(v) for v in rng.integers(0, n_nodes, 2))
        if a != b:
            edges.add((min(a, b), max(a, b)))
    conn = np.array(sorted(edges), dtype=np.int32)
    conn = conn[rng.permutation(len(conn))]
    flip = rng.uniform(size=len(conn)) < 0.5              # beam direction i -> j either way
    conn[flip] = conn[flip][:, ::-1]
    nodes = rng.permutation(n_nodes)
    fixed = []
    for k in nodes[:n_clamped]:
        fixed += [3 * k, 3 * k + 1, 3 * k + 2]
    for k in nodes[n_clamped:n_clamped + n_pinned]:
        fixed += [3 * k, 3 * k + 1]
    return xy, conn, sorted(int(f) for f in fixed)


CASES = [(2, 0, 1, 0), (3, 0, 1, 0), (4, 1, 1, 1), (6, 2, 1, 1), (9, 3, 2, 0), (12, 4, 1, 2), (16, 6, 2, 1),
         (20, 8, 1, 0), (21, 5, 1, 1), (25, 10, 2, 2), (30, 12, 2, 0), (36, 14, 2, 3), (40, 16, 1, 1)]


@pytest.mark.parametrize('reorder', [False, True])
@pytest.mark.parametrize('n_nodes,extra,n_cl,n_pin', CASES)
def test_random_frame_static_and_modal(n_nodes, extra, n_cl, n_pin, reorder):
    rng = np.random.default_rng(1000 * n_nodes + extra)
    xy, conn, fixed = random_frame(rng, n_nodes, extra, n_cl, n_pin)
    nE = len(conn)
    n_free = 3 * n_nodes - len(fixed)
    if n_free > 116:
        pytest.skip('beyond the shared-memory Jacobi path')
    B, C = 5, 2
    coords = xy[None] + rng.uniform(-20, 20, (B, n_nodes, 2))
    props = np.stack([rng.uniform(1.9e5, 2.2e5, (B, nE)), rng.uniform(2e3, 2e4, (B, nE)),
                      rng.uniform(1e6, 1e8, (B, nE)), rng.uniform(7e-9, 8.5e-9, (B, nE))], axis=-1)
    lumped = rng.uniform(size=nE) < 0.0                   # consistent mass (the lumped matrix is ill-posed)
    mass = rng.uniform(0, 1e-3, (B, n_nodes, 2)) * (rng.uniform(size=(1, n_nodes, 1)) < 0.3)
    f_nodal = rng.uniform(-1e4, 1e4, (B, C, 3 * n_nodes)) * (rng.uniform(size=(1, C, 3 * n_nodes)) < 0.2)
    f_nodal[:, :, 0] += 1.0                               # never an all-zero case
    L, _ = orc.length_direction(coords, conn)
    q = rng.uniform(-5, 5, (B, C, nE)) * (rng.uniform(size=(1, C, nE)) < 0.4)
    Lc = np.broadcast_to(L[:, None, :], q.shape)
    r_local = np.zeros((B, C, nE, 6))                     # Loads.py:264-270
    r_local[..., 1] = q * Lc / 2
    r_local[..., 2] = q * (Lc ** 2) / 12
    r_local[..., 4] = q * Lc / 2
    r_local[..., 5] = -1 * q * (Lc ** 2) / 12.
    plan = Plan(n_nodes, conn, fixed, lumped, reorder=reorder)     # reorder: internal RCM numbering of the free DOFs
    assert plan.n_free == n_free
    keep_sorted = orc.free_positions(n_nodes, fixed)
    assert sorted(plan.pos_of_free.tolist()) == keep_sorted.tolist()
    perm = np.searchsorted(keep_sorted, plan.pos_of_free)           # internal free index -> ascending free index
    if not reorder:
        assert np.array_equal(perm, np.arange(n_free))
    T = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=DEV)     # noqa: E731
    if plan.smem_bytes(C) > 227 * 1024:                   # wide band AND large n: loud refusal, not a wrong answer
        with pytest.raises(Exception, match='shared memory'):
            batch.solve_fused(plan, T(coords), T(props), T(mass), T(f_nodal), T(r_local), n_modes=4)
        return
    out = batch.solve_fused(plan, T(coords), T(props), T(mass), T(f_nodal), T(r_local), n_modes=min(n_free, 8))
    assert (out['info'] == 0).all(), out['info']
    o = orc.solve_batch(conn, fixed, coords, props, mass, f_nodal, r_local, lumped, n_modes=min(n_free, 8))
    # band width claim and unfused assembly
    Kb, Mb = batch.assemble_banded(plan, T(coords), T(props), T(mass))
    for b in range(B):
        Kp, Mp = o['Kc'][b][np.ix_(perm, perm)], o['Mc'][b][np.ix_(perm, perm)]
        assert util.relmax(np.tril(plan.band_to_dense(Kb[b].cpu().numpy())), np.tril(Kp)) < 1e-14
        assert util.relmax(np.tril(plan.band_to_dense(Mb[b].cpu().numpy())), np.tril(Mp)) < 1e-14
    condK = np.linalg.cond(o['Kc']).max()
    tol_s = max(1e-10, 50 * EPS * condK)                  # both solvers are backward stable: forward error ~ eps cond
    for b in range(B):
        assert util.relmax(out['u'][b].cpu().numpy(), o['u'][b]) < tol_s
        assert util.relmax(out['endforces'][b].cpu().numpy(), o['endforces'][b]) < tol_s
        assert util.relmax(out['reactions'][b].cpu().numpy(), o['reactions'][b]) < tol_s
    # equilibrium at the free DOFs: reactions vanish there
    free = torch.as_tensor(keep_sorted.astype(np.int64), device=DEV)
    react = out['reactions']
    assert (react[:, :, free].abs().amax() / react.abs().amax()).item() < max(1e-9, 50 * EPS * condK)
    lam = out['lam'].cpu().numpy()
    bound = 1e-10 + 64 * EPS * o['lam'][:, -1:] / o['lam']
    assert (np.abs(lam / o['lam'] - 1) < bound).all()
    phi = out['phi'].cpu().numpy()
    keep = keep_sorted
    for b in range(B):
        Pf = phi[b][:, keep]
        assert np.abs(Pf @ o['Mc'][b] @ Pf.T - np.eye(Pf.shape[0])).max() < 1e-9
        res = o['Kc'][b] @ Pf.T - (o['Mc'][b] @ Pf.T) * lam[b, None, :Pf.shape[0]]
        assert (np.abs(res).max(axis=0) / (np.abs(o['Kc'][b]).max() * np.abs(Pf).max(axis=1))).max() < 1e-11
    gap = np.minimum(np.diff(o['lam'], axis=1, prepend=-np.inf), np.diff(o['lam'], axis=1, append=np.inf))[:, :phi.shape[1]]
    tol_v = 1e-10 + 256 * EPS * o['lam'][:, -1:] / gap
    assert (util.mode_err(phi, o['phi']) < tol_v).all()


def test_first_modes_of_a_frame_beyond_the_jacobi_path():
    """A 3-storey, 5-bay frame with 2 elements per member (n_free = 189 > 116): the lowest modes by subspace
    iteration (frame_modes.FrameModes) against scipy.linalg.eigh on the oracle's dense K, M; and the drop-in
    ModalSolver taking that route."""
    import scipy.linalg as sla
    from beamfe2_b200.frame_modes import FrameModes

    def T(a):
        return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=DEV)
    from beamfe2_b200 import HermitianBeam_2D as HB, Node, Structure
    storeys, bays, H, Lb = 3, 5, 3000.0, 5000.0
    nodes, idx = [], {}

    def node(x, y):
        key = (round(x, 6), round(y, 6))
        if key not in idx:
            idx[key] = len(nodes)
            nodes.append((x, y))
        return idx[key]

    conn, kinds = [], []
    for b in range(bays + 1):                                    # columns, two elements per storey
        for s_ in range(storeys):
            ys = [s_ * H, s_ * H + H / 2, (s_ + 1) * H]
            for a, c in zip(ys[:-1], ys[1:]):
                conn.append((node(b * Lb, a), node(b * Lb, c)))
                kinds.append(0)
    for s_ in range(1, storeys + 1):                             # girders, two elements per bay
        for b in range(bays):
            xs = [b * Lb, b * Lb + Lb / 2, (b + 1) * Lb]
            for a, c in zip(xs[:-1], xs[1:]):
                conn.append((node(a, s_ * H), node(c, s_ * H)))
                kinds.append(1)
    xy = np.array(nodes)
    conn = np.array(conn, dtype=np.int32)
    props = np.array([[2.1e5, 300.0 * 300.0, 300.0 ** 4 / 12, 7.85e-9] if k == 0 else
                      [2.1e5, 200.0 * 500.0, 200.0 * 500.0 ** 3 / 12, 7.85e-9] for k in kinds])
    fixed = sorted(3 * i + d for i, (x, y) in enumerate(nodes) if y == 0.0 for d in range(3))
    plan = Plan(len(nodes), conn, fixed, reorder=True)
    assert plan.n_free > 116
    mass = np.zeros((len(nodes), 2))
    mass[[i for i, (x, y) in enumerate(nodes) if y > 0 and abs(y / H - round(y / H)) < 1e-9], :] = 2.0
    fm = FrameModes(plan, T(xy), T(props), T(mass))
    out = fm.modes(n_modes=20)
    assert out['converged']
    o = orc.solve_batch(conn, fixed, xy, props, nodal_mass=mass)
    w = sla.eigh(o['Kc'], o['Mc'], eigvals_only=True)
    lam = out['lam'].cpu().numpy()
    assert np.abs(lam / w[:20] - 1).max() < 1e-8
    keep = np.setdiff1d(np.arange(plan.sumdof), fixed)           # the oracle's condensed (ascending) numbering
    phi = out['phi'].cpu().numpy()[:, keep]
    assert np.abs(phi @ o['Mc'] @ phi.T - np.eye(20)).max() < 1e-8
    assert float(out['residual'].max().item()) < 1e-9
    # drop-in route
    ns = [Node.Node(ID=i + 1, coords=tuple(p)) for i, p in enumerate(nodes)]
    beams = [HB.HermitianBeam2D(ID=e + 1, E=props[e, 0], A=props[e, 1], I=props[e, 2], rho=props[e, 3],
                                i=ns[conn[e, 0]], j=ns[conn[e, 1]]) for e in range(len(conn))]
    sup = {i + 1: ['ux', 'uy', 'rotz'] for i, (x, y) in enumerate(nodes) if y == 0.0}
    st = Structure.Structure(beams=beams, supports=sup)
    for i, (x, y) in enumerate(nodes):
        if mass[i, 0] > 0:
            st.add_nodal_mass(nodeID=i + 1, mass={'mx': 2.0, 'my': 2.0})
    assert st.solver['modal'].solve()
    om = np.array(st.results['modal'].circular_frequencies[:20])
    assert np.abs(om / np.sqrt(w[:20]) - 1).max() < 1e-8


def test_first_modes_of_many_frames_in_one_batch():
    """BatchedFrameModes: 24 variants (sections, moduli, masses, node jitter) of a 3-storey, 4-bay frame with
    n_free = 150 > 116 in ONE sequence of launches; every variant's 12 lowest eigenpairs against scipy.linalg.eigh on
    the oracle's dense matrices, and against the single-structure FrameModes."""
    import scipy.linalg as sla
    from beamfe2_b200.frame_modes import BatchedFrameModes, FrameModes
    storeys, bays, H, Lb = 3, 4, 3000.0, 5000.0
    nodes, idx = [], {}

    def node(x, y):
        key = (round(x, 6), round(y, 6))
        if key not in idx:
            idx[key] = len(nodes)
            nodes.append((x, y))
        return idx[key]
    conn, kinds = [], []
    for b in range(bays + 1):
        for s_ in range(storeys):
            ys = [s_ * H, s_ * H + H / 2, (s_ + 1) * H]
            for a, c in zip(ys[:-1], ys[1:]):
                conn.append((node(b * Lb, a), node(b * Lb, c)))
                kinds.append(0)
    for s_ in range(1, storeys + 1):
        for b in range(bays):
            xs = [b * Lb, b * Lb + Lb / 2, (b + 1) * Lb]
            for a, c in zip(xs[:-1], xs[1:]):
                conn.append((node(a, s_ * H), node(c, s_ * H)))
                kinds.append(1)
    xy = np.array(nodes)
    conn = np.array(conn, dtype=np.int32)
    kinds = np.array(kinds)
    fixed = sorted(3 * i + d for i, (x, y) in enumerate(nodes) if y == 0.0 for d in range(3))
    plan = Plan(len(nodes), conn, fixed, reorder=True)
    assert plan.n_free > 116
    B = 24
    rng = np.random.default_rng(5)
    coords = xy[None] + rng.uniform(-50, 50, (B,) + xy.shape) * (xy[None, :, 1:2] > 0)      # supports stay put
    base = np.where(kinds[:, None] == 0, [[2.1e5, 300.0 * 300.0, 300.0 ** 4 / 12, 7.85e-9]],
                    [[2.1e5, 200.0 * 500.0, 200.0 * 500.0 ** 3 / 12, 7.85e-9]])
    props = base[None] * rng.uniform(0.7, 1.4, (B, len(conn), 1))
    mass = np.zeros((B, len(nodes), 2))
    floor = [i for i, (x, y) in enumerate(nodes) if y > 0 and abs(y / H - round(y / H)) < 1e-9]
    mass[:, floor, :] = rng.uniform(0.5, 4.0, (B, 1, 1))

    def T(a):
        return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=DEV)
    bf = BatchedFrameModes(plan, T(coords), T(props), T(mass))
    out = bf.modes(n_modes=12)
    assert bool(out['converged'].all()) and out['iterations'] < 40
    assert float(out['residual'].max().item()) < 1e-9
    lam = out['lam'].cpu().numpy()
    keep = np.setdiff1d(np.arange(plan.sumdof), fixed)
    for b in range(0, B, 5):
        o = orc.solve_batch(conn, fixed, coords[b], props[b], nodal_mass=mass[b])
        w = sla.eigh(o['Kc'], o['Mc'], eigvals_only=True)
        assert np.abs(lam[b] / w[:12] - 1).max() < 1e-8, b
        phi = out['phi'][b].cpu().numpy()[:, keep]
        assert np.abs(phi @ o['Mc'] @ phi.T - np.eye(12)).max() < 1e-8
    one = FrameModes(plan, T(coords[3]), T(props[3]), T(mass[3])).modes(n_modes=12)
    assert np.abs(lam[3] / one['lam'].cpu().numpy() - 1).max() < 1e-9


@pytest.mark.parametrize('n_nodes,extra,C,B,reorder', [
    (4, 0, 1, 1, False), (6, 1, 3, 3, True), (9, 2, 5, 45, True), (12, 3, 17, 7, True),
    (16, 4, 2, 1000, True), (20, 0, 3, 33, True), (25, 6, 4, 5, False), (30, 3, 3, 129, True)])
def test_static_only_kernels_on_general_frames(n_nodes, extra, C, B, reorder):
    """The static-only entry points (fused from coordinates, and K2 -> K3 from the band in HBM) on general
    frames: narrow bands run the 16-lanes-per-structure kernel (5- and 8-wide instantiations, with and without
    the early right-hand-side path, more load cases than lanes), wide bands the block / warp fallbacks; ragged
    batch sizes; optional outputs."""
    rng = np.random.default_rng(77 * n_nodes + C)
    xy, conn, fixed = random_frame(rng, n_nodes, extra, 1, 1)
    _check_static_only(rng, xy, conn, fixed, C, B, reorder)


@pytest.mark.parametrize('n_nodes,skip,C,B', [(3, 0, 1, 1), (8, 0, 3, 45), (21, 0, 3, 1000), (21, 0, 17, 7),
                                              (12, 1, 1, 3), (20, 1, 3, 130), (15, 1, 20, 33), (40, 0, 2, 9)])
def test_static_only_kernels_on_narrow_bands(n_nodes, skip, C, B):
    """Chains (half bandwidth 5) and chains with extra beams from node k to k + 2 (half bandwidth 8): the two
    instantiations of the 16-lanes-per-structure kernel, 1 .. 20 load cases, ragged batches."""
    rng = np.random.default_rng(5 * n_nodes + C + skip)
    xy = np.stack([np.arange(n_nodes) * 700.0, 300.0 * np.sin(np.arange(n_nodes))], axis=1)
    conn = [(k, k + 1) for k in range(n_nodes - 1)]
    if skip:
        conn += [(k, k + 2) for k in range(0, n_nodes - 2, 2)]
    conn = np.array(conn, dtype=np.int32)
    fixed = [0, 1, 2] + ([3 * (n_nodes - 1) + 1] if n_nodes > 4 else [])
    plan = Plan(n_nodes, conn, fixed)
    assert plan.hbw == (8 if skip else 5)
    _check_static_only(rng, xy, conn, fixed, C, B, False)


def _check_static_only(rng, xy, conn, fixed, C, B, reorder):
    n_nodes = len(xy)
    nE = len(conn)
    coords = xy[None] + rng.uniform(-20, 20, (B, n_nodes, 2))
    props = np.stack([rng.uniform(1.9e5, 2.2e5, (B, nE)), rng.uniform(2e3, 2e4, (B, nE)),
                      rng.uniform(1e6, 1e8, (B, nE)), rng.uniform(7e-9, 8.5e-9, (B, nE))], axis=-1)
    f_nodal = rng.uniform(-1e4, 1e4, (B, C, 3 * n_nodes)) * (rng.uniform(size=(1, C, 3 * n_nodes)) < 0.3)
    f_nodal[:, :, 0] += 1.0
    L, _ = orc.length_direction(coords, conn)
    q = rng.uniform(-5, 5, (B, C, nE)) * (rng.uniform(size=(1, C, nE)) < 0.5)
    Lc = np.broadcast_to(L[:, None, :], q.shape)
    r_local = np.zeros((B, C, nE, 6))
    r_local[..., 1] = q * Lc / 2
    r_local[..., 2] = q * (Lc ** 2) / 12
    r_local[..., 4] = q * Lc / 2
    r_local[..., 5] = -1 * q * (Lc ** 2) / 12.
    plan = Plan(n_nodes, conn, fixed, None, reorder=reorder)
    T = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=DEV)     # noqa: E731
    nb = min(B, 6)                                                     # oracle on the first and last few variants
    pick = np.r_[0:nb // 2, B - (nb - nb // 2):B] if B > nb else np.arange(B)
    o = orc.solve_batch(conn, fixed, coords[pick], props[pick], None, f_nodal[pick], r_local[pick])
    condK = np.linalg.cond(o['K'][:, orc.free_positions(n_nodes, fixed)][:, :, orc.free_positions(n_nodes, fixed)]).max()
    tol = max(1e-10, 50 * EPS * condK)
    fused = batch.solve_fused(plan, T(coords), T(props), None, T(f_nodal), T(r_local), modal=False)
    Kb, _ = batch.assemble_banded(plan, T(coords), T(props), None)
    unfused = batch.static_solve(plan, Kb, T(coords), T(props), T(f_nodal), T(r_local))
    for out in (fused, unfused):
        assert (out['info'] == 0).all()
        for k, b in enumerate(pick):
            for key in ('u', 'endforces', 'reactions'):
                assert util.relmax(out[key][b].cpu().numpy(), o[key][k]) < tol, (key, int(b))
    # optional outputs and no beam loads: displacements do not depend on what else is requested
    u_only = batch.solve_fused(plan, T(coords), T(props), None, T(f_nodal), T(r_local), modal=False,
                               want_endforces=False, want_reactions=False)
    assert torch.equal(u_only['u'], fused['u']) and u_only['endforces'] is None and u_only['reactions'] is None
    o0 = orc.solve_batch(conn, fixed, coords[pick], props[pick], None, f_nodal[pick], np.zeros_like(r_local[pick]))
    nob = batch.solve_fused(plan, T(coords), T(props), None, T(f_nodal), None, modal=False)
    for k, b in enumerate(pick):
        assert util.relmax(nob['u'][b].cpu().numpy(), o0['u'][k]) < tol
        assert util.relmax(nob['reactions'][b].cpu().numpy(), o0['reactions'][k]) < tol


def test_structures_beyond_shared_memory_use_the_hbm_workspace_path():
    """A 20 x 20 grid frame (400 nodes, ~1100 free DOFs, half bandwidth ~60 after RCM: 600 KB of band) does not
    fit the shared memory of an SM: both static entry points run the HBM-workspace kernel, and the first-k modal
    route (which applies K^-1 through bfe2_static_solve) works on top of it."""
    import scipy.linalg as sla
    from beamfe2_b200.frame_modes import FrameModes
    rng = np.random.default_rng(2020)
    side = 20
    xy = np.array([(1000.0 * i, 1000.0 * j) for j in range(side) for i in range(side)]) + rng.uniform(-50, 50, (side * side, 2))
    conn = []
    for j in range(side):
        for i in range(side):
            k = j * side + i
            if i + 1 < side:
                conn.append((k, k + 1))
            if j + 1 < side:
                conn.append((k, k + side))
    conn = np.array(conn, dtype=np.int32)
    nN, nE = side * side, len(conn)
    fixed = [3 * k + d for k in range(side) for d in range(3)]                  # bottom row clamped
    plan = Plan(nN, conn, fixed, reorder=True)
    assert plan.smem_bytes(2) > 227 * 1024
    B, C = 2, 2
    coords = xy[None] + rng.uniform(-5, 5, (B, nN, 2))
    props = np.stack([rng.uniform(1.9e5, 2.2e5, (B, nE)), rng.uniform(2e3, 2e4, (B, nE)),
                      rng.uniform(1e6, 1e8, (B, nE)), rng.uniform(7e-9, 8.5e-9, (B, nE))], axis=-1)
    f_nodal = rng.uniform(-1e4, 1e4, (B, C, 3 * nN)) * (rng.uniform(size=(1, C, 3 * nN)) < 0.1)
    L, _ = orc.length_direction(coords, conn)
    q = rng.uniform(-5, 5, (B, C, nE)) * (rng.uniform(size=(1, C, nE)) < 0.3)
    Lc = np.broadcast_to(L[:, None, :], q.shape)
    r_local = np.zeros((B, C, nE, 6))
    r_local[..., 1] = q * Lc / 2
    r_local[..., 2] = q * (Lc ** 2) / 12
    r_local[..., 4] = q * Lc / 2
    r_local[..., 5] = -1 * q * (Lc ** 2) / 12.
    T = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=DEV)     # noqa: E731
    o = orc.solve_batch(conn, fixed, coords, props, np.zeros((B, nN, 2)), f_nodal, r_local, n_modes=1)
    condK = np.linalg.cond(o['Kc']).max()
    tol = max(1e-10, 50 * EPS * condK)
    fused = batch.solve_fused(plan, T(coords), T(props), None, T(f_nodal), T(r_local), modal=False)
    Kb, _ = batch.assemble_banded(plan, T(coords), T(props), None)
    unfused = batch.static_solve(plan, Kb, T(coords), T(props), T(f_nodal), T(r_local))
    for out in (fused, unfused):
        assert (out['info'] == 0).all()
        for key in ('u', 'endforces', 'reactions'):
            assert util.relmax(out[key].cpu().numpy(), o[key]) < tol, key
    fm = FrameModes(plan, T(coords[0]), T(props[0]), None)
    res = fm.modes(n_modes=8)
    w = sla.eigh(o['Kc'][0], o['Mc'][0], eigvals_only=True, subset_by_index=[0, 7])
    assert np.abs(res['lam'].cpu().numpy() / w - 1).max() < 1e-7
