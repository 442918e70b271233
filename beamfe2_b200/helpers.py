"""Module-level helpers with the reference's names (BeamFE2/helpers.py): `compile_global_matrix` (:7-40) runs the
CUDA assembly kernel (K2) on the uncondensed topology plan, `transfer_matrix` (:43-69) and `np_matrix_tolist`
(:72-78) are the small host utilities user scripts import alongside it."""
import itertools
import math

import numpy as np


def compile_global_matrix(beams, stiffness=False, mass=False, geometrical=False):
    """Dense global K (stiffness=True) or M (mass=True) of a list of drop-in beams, assembled on the GPU in
    beam-list order with the block positions (ID - 1) * 3 of the reference."""
    assert stiffness or mass or geometrical
    import torch
    from . import batch
    from .plan import Plan
    nodes = sorted(set(itertools.chain.from_iterable(b.nodes for b in beams)), key=lambda n: n.ID)
    if [n.ID for n in nodes] != list(range(1, len(nodes) + 1)):
        raise Exception('Node numbering is not ok: IDs must be 1..N without gaps')
    conn = np.array([[b.i.ID - 1, b.j.ID - 1] for b in beams], dtype=np.int32)
    lumped = np.array([b.mass_matrix == 'lumped' for b in beams], dtype=np.uint8)
    plan = Plan(len(nodes), conn, [], lumped)

    def dev(a):
        return torch.as_tensor(np.ascontiguousarray(a)[None], dtype=torch.float64, device='cuda')
    coords = dev([[float(n.x), float(n.y)] for n in nodes])
    props = dev([[float(b.E), float(b.A), float(b.I), float(b.rho if b.rho is not None else 0.0)] for b in beams])
    if geometrical and not (stiffness or mass):
        # helpers.py:26-27 sums b.Ke_geom, i.e. _Ke_geom() with its default N = -1 for every beam
        axial = torch.full((1, len(beams)), -1.0, dtype=torch.float64, device='cuda')
        band = batch.assemble_geometric(plan, coords, axial)[0].cpu().numpy()
        return np.matrix(plan.band_to_dense(band))
    Kb, Mb = batch.assemble_banded(plan, coords, props, None, want_mass=bool(mass))
    torch.cuda.synchronize()
    band = (Kb if stiffness else Mb)[0].cpu().numpy()
    return np.matrix(plan.band_to_dense(band))


def transfer_matrix(alpha, asdegree=False, blocks=2, blocksize=3):
    """Block-diagonal rotation matrix diag(R, ..., R), R = [[c, -s(, 0)], [s, c(, 0)](, [0, 0, 1])]."""
    if asdegree:
        alpha = math.radians(alpha)
    c, s = math.cos(alpha), math.sin(alpha)
    if blocksize == 3:
        R = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=float)
    elif blocksize == 2:
        R = np.array([[c, -s], [s, c]], dtype=float)
    else:
        raise Exception('not implemented, dof should be 2 or 3')
    return np.matrix(np.kron(np.eye(blocks), R))


def np_matrix_tolist(mtrx):
    return list(itertools.chain.from_iterable(mtrx.tolist()))
