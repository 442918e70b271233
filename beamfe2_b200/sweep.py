"""Synthetic structure generators for the named configs.

P20  = 20-element portal frame (7 + 6 + 7 elements numbered as a chain, 21 nodes, 63 DOFs, both feet
       clamped -> positions_to_eliminate = [0,1,2,60,61,62], 57 free DOFs, half-bandwidth 5,
       3 load cases) -- configs 3 and 5.
BEAM = one straight cantilever discretised in n equal elements -- config 4.

The raw per-variant parameters are drawn on the host with numpy (so that the stream is the one the
survey names: numpy.random.default_rng(20261019), or default_rng([20261019, shard]) per shard);
turning them into the packed coords / props / masses / load tensors is a handful of torch
broadcasts that run on whatever device the parameters live on.
"""
from __future__ import annotations

import numpy as np
import torch

SEED = 20261019
P20_PARAM_NAMES = ('H', 'L', 'h_c', 'b_c', 'h_g', 'b_g', 'E', 'rho', 'm', 's')
P20_PARAM_RANGES = ((3000., 6000.), (4000., 10000.), (150., 400.), (150., 400.), (200., 600.),
                    (150., 300.), (1.9e5, 2.2e5), (7.0e-9, 8.5e-9), (0., 5.), (0.5, 2.))
P20_N_NODES = 21
P20_N_ELEM = 20
P20_N_LOADCASES = 3


def p20_topology():
    """(conn [20,2] int32 0-based, supports {nodeID: dofnames}) for the chain-numbered portal."""
    conn = np.stack([np.arange(0, 20), np.arange(1, 21)], axis=1).astype(np.int32)
    supports = {1: ['ux', 'uy', 'rotz'], 21: ['ux', 'uy', 'rotz']}
    return conn, supports


def p20_draw_params(B: int, shard=None) -> np.ndarray:
    """[B, 10] raw parameters, each column drawn as one [B] uniform array in P20_PARAM_NAMES order."""
    rng = np.random.default_rng(SEED if shard is None else [SEED, int(shard)])
    cols = [rng.uniform(lo, hi, size=B) for (lo, hi) in P20_PARAM_RANGES]
    return np.stack(cols, axis=1)


def p20_pack(params: torch.Tensor):
    """Raw [B,10] parameters -> packed per-variant tensors on params.device.

    Returns dict(coords [B,21,2], props [B,20,4], nodal_mass [B,21,2], f_nodal [B,3,63],
    r_local [B,3,20,6]).  Load cases: LC1 FX = s*1e4 at node 8; LC2 uniform perpendicular
    q = -s*10 on the 6 girder elements (local fixed-end vector [0, qL/2, qL^2/12, 0, qL/2, -qL^2/12],
    reference Loads.py:264-270); LC3 = LC1 + LC2 + MZ = s*1e6 at node 14.
    """
    p = params.to(torch.float64)
    B = p.shape[0]
    dev = p.device
    H, L, h_c, b_c, h_g, b_g, E, rho, m, s = (p[:, k] for k in range(10))
    k7 = torch.arange(8, dtype=torch.float64, device=dev)
    k6 = torch.arange(7, dtype=torch.float64, device=dev)
    coords = torch.zeros((B, 21, 2), dtype=torch.float64, device=dev)
    coords[:, 0:8, 1] = H[:, None] * k7[None, :] / 7.0               # left column, nodes 1..8
    coords[:, 7:14, 0] = L[:, None] * k6[None, :] / 6.0              # girder, nodes 8..14
    coords[:, 7:14, 1] = H[:, None]
    coords[:, 13:21, 0] = L[:, None]                             # right column, nodes 14..21
    coords[:, 13:21, 1] = H[:, None] - H[:, None] * k7[None, :] / 7.0
    A_c = h_c * b_c
    I_c = h_c * h_c * h_c * b_c / 12.0
    A_g = h_g * b_g
    I_g = h_g * h_g * h_g * b_g / 12.0
    props = torch.empty((B, 20, 4), dtype=torch.float64, device=dev)
    props[:, :, 0] = E[:, None]
    props[:, :, 3] = rho[:, None]
    props[:, 0:7, 1] = A_c[:, None]
    props[:, 0:7, 2] = I_c[:, None]
    props[:, 7:13, 1] = A_g[:, None]
    props[:, 7:13, 2] = I_g[:, None]
    props[:, 13:20, 1] = A_c[:, None]
    props[:, 13:20, 2] = I_c[:, None]
    mass = torch.zeros((B, 21, 2), dtype=torch.float64, device=dev)
    mass[:, 7:14, :] = m[:, None, None]
    f_nodal = torch.zeros((B, 3, 63), dtype=torch.float64, device=dev)
    f_nodal[:, 0, 7 * 3 + 0] = s * 1e4
    f_nodal[:, 2, 7 * 3 + 0] = s * 1e4
    f_nodal[:, 2, 13 * 3 + 2] = s * 1e6
    r_local = torch.zeros((B, 3, 20, 6), dtype=torch.float64, device=dev)
    # girder element length exactly as the reference computes it from node coordinates
    gx = coords[:, 7:14, 0]
    gy = coords[:, 7:14, 1]
    Lg = torch.sqrt((gx[:, :-1] - gx[:, 1:]) ** 2 + (gy[:, :-1] - gy[:, 1:]) ** 2)   # [B,6]
    q = (-s * 10.0)[:, None]
    r = torch.zeros((B, 6, 6), dtype=torch.float64, device=dev)
    r[:, :, 1] = q * Lg / 2
    r[:, :, 2] = q * (Lg ** 2) / 12
    r[:, :, 4] = q * Lg / 2
    r[:, :, 5] = -1 * q * (Lg ** 2) / 12.
    r_local[:, 1, 7:13, :] = r
    r_local[:, 2, 7:13, :] = r
    return dict(coords=coords, props=props, nodal_mass=mass, f_nodal=f_nodal, r_local=r_local)


def p20_model_dict(params_row):
    """One variant as the plain 'model' dict understood by oracle/ref_harness.py (tests only)."""
    H, L, h_c, b_c, h_g, b_g, E, rho, m, s = (float(x) for x in params_row)
    nodes = [(0.0, H * k / 7) for k in range(7)]                 # nodes 1..7
    nodes += [(L * k / 6, H) for k in range(0, 6)]               # nodes 8..13 (node 8 = (0, H))
    nodes += [(L, H - H * k / 7) for k in range(0, 8)]           # nodes 14..21 (node 14 = (L, H))
    beams = []
    for e in range(20):
        if 7 <= e < 13:
            A, I = h_g * b_g, h_g * h_g * h_g * b_g / 12.0
        else:
            A, I = h_c * b_c, h_c * h_c * h_c * b_c / 12.0
        beams.append(dict(i=e + 1, j=e + 2, E=E, A=A, I=I, rho=rho))
    q = -s * 10.0
    girder = [(e, 'uniform perpendicular force', dict(value=q)) for e in range(7, 13)]
    lcs = [dict(nodal=[(8, {'FX': s * 1e4})]),
           dict(beam=girder),
           dict(nodal=[(8, {'FX': s * 1e4}), (14, {'MZ': s * 1e6})], beam=girder)]
    masses = [(k, {'mx': m, 'my': m}) for k in range(8, 15)]
    return dict(nodes=nodes, beams=beams, supports={1: ['ux', 'uy', 'rotz'], 21: ['ux', 'uy', 'rotz']},
                masses=masses, load_cases=lcs)


# ----------------------------------------------------------------------------- config 4
def beam_topology(n_elem: int):
    conn = np.stack([np.arange(0, n_elem), np.arange(1, n_elem + 1)], axis=1).astype(np.int32)
    supports = {1: ['ux', 'uy', 'rotz']}
    return conn, supports


def beam_pack(n_elem: int, length=1000.0, A=30.0, I=22.5, E=2.1e5, rho=7.85e-9, tip_fy=100.0,
              device='cpu'):
    """Config 4: straight cantilever along +x, node 1 clamped, tip FY (one structure, batch of 1)."""
    x = torch.arange(n_elem + 1, dtype=torch.float64, device=device) * (length / n_elem)
    coords = torch.zeros((1, n_elem + 1, 2), dtype=torch.float64, device=device)
    coords[0, :, 0] = x
    props = torch.empty((1, n_elem, 4), dtype=torch.float64, device=device)
    props[..., 0] = E
    props[..., 1] = A
    props[..., 2] = I
    props[..., 3] = rho
    f = torch.zeros((1, 1, 3 * (n_elem + 1)), dtype=torch.float64, device=device)
    f[0, 0, 3 * n_elem + 1] = tip_fy
    return dict(coords=coords, props=props, f_nodal=f)
