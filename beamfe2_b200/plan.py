"""Topology plan: the integer side of the hot path (DOF map, support elimination, band-slot scatter
plan), built once per topology by libbfe2 (C++) and shared by every variant of a batch.

Mirrors Structure.position_in_matrix / positions_to_eliminate / condense (reference
BeamFE2/Structure.py:164-191, 379-393) and the block indexing of helpers.compile_global_matrix
(BeamFE2/helpers.py:28-39); tests compare these arrays bit-exactly with the reference.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

DOF = 3
DOFNAMES = ('ux', 'uy', 'rotz')
LOADNAMES = ('FX', 'FY', 'MZ')
MASSNAMES = ('mx', 'my')


def position_in_matrix(nodeID: int, DOF: str = None, dynam: str = None) -> int:
    """Same contract as Structure.position_in_matrix (Structure.py:379-393); nodeID is 1-based."""
    if DOF is not None:
        assert DOF in DOFNAMES
        return (nodeID - 1) * 3 + DOFNAMES.index(DOF)
    if dynam is not None:
        assert dynam in LOADNAMES
        return (nodeID - 1) * 3 + LOADNAMES.index(dynam)
    return None


def positions_to_eliminate(supports: dict) -> list:
    """Same contract as Structure.positions_to_eliminate (Structure.py:164-181)."""
    out = [position_in_matrix(nodeID=k, DOF=d) for k, dofs in supports.items() for d in dofs]
    out.sort()
    return out


class Plan:
    """Owns a bfe2_plan handle.  conn: [nE,2] 0-based node indices in beam-list order."""

    def __init__(self, n_nodes: int, conn, fixed_positions, lumped=None, reorder: bool = False):
        L = _lib.lib()
        conn = np.ascontiguousarray(np.asarray(conn, dtype=np.int32).reshape(-1, 2))
        fixed = np.ascontiguousarray(np.asarray(list(fixed_positions), dtype=np.int32).ravel())
        lump = None
        if lumped is not None:
            lump = np.ascontiguousarray(np.asarray(lumped, dtype=np.uint8).ravel())
            assert lump.shape[0] == conn.shape[0]
        h = C.c_void_p()
        _lib.check(L.bfe2_plan_create_ex(C.byref(h), int(n_nodes), int(conn.shape[0]), _lib.ptr(conn),
                                         int(fixed.shape[0]), _lib.ptr(fixed) if fixed.size else None,
                                         _lib.ptr(lump), 1 if reorder else 0))
        self._h = h
        d = [C.c_int32() for _ in range(6)]
        _lib.check(L.bfe2_plan_dims(h, *[C.byref(x) for x in d]))
        self.n_nodes, self.n_elem, self.sumdof, self.n_free, self.hbw, self.n_contrib = (x.value for x in d)
        self.conn = conn
        self.lumped = lump

    @classmethod
    def from_supports(cls, n_nodes, conn, supports: dict, lumped=None, reorder: bool = False):
        return cls(n_nodes, conn, positions_to_eliminate(supports), lumped, reorder)

    @property
    def handle(self):
        return self._h

    def array(self, which: int) -> np.ndarray:
        cap = max(self.n_contrib, (self.hbw + 1) * self.n_free + 1, self.sumdof + 1, 2 * self.n_elem) + 8
        out = np.empty(cap, dtype=np.int32)
        k = _lib.check(_lib.lib().bfe2_plan_get_array(self._h, which, _lib.ptr(out), cap))
        return out[:k].copy()

    @property
    def fixed(self):
        return self.array(_lib.ARR_FIXED)

    @property
    def pos_of_free(self):
        return self.array(_lib.ARR_POS_OF_FREE)

    @property
    def free_of_pos(self):
        return self.array(_lib.ARR_FREE_OF_POS)

    def smem_bytes(self, n_loadcases: int) -> int:
        return int(_lib.lib().bfe2_plan_smem_bytes(self._h, int(n_loadcases)))

    def band_to_dense(self, Xb):
        """[..., hbw+1, n] lower band -> dense symmetric [..., n, n] (numpy; test/diagnostic helper)."""
        Xb = np.asarray(Xb)
        n, b = self.n_free, self.hbw
        out = np.zeros(Xb.shape[:-2] + (n, n))
        for d in range(b + 1):
            idx = np.arange(n - d)
            out[..., idx + d, idx] = Xb[..., d, :n - d]
            out[..., idx, idx + d] = Xb[..., d, :n - d]
        return out

    def __del__(self):
        try:
            if getattr(self, '_h', None):
                _lib.lib().bfe2_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass
