"""CPU-only checks of the C-ABI library and of the integer topology plan (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest

from beamfe2_b200 import _lib, plan as planmod, sweep
from beamfe2_b200.plan import Plan
from oracle import beamfe2_oracle as orc
from tests import util

ROOT = util.ROOT


def header_symbols():
    txt = open(os.path.join(ROOT, 'include', 'bfe2.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(bfe2_[a-z0-9_]+)\s*\(', txt)))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    syms = header_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(L, s), 'libbfe2.so does not export %s' % s
    assert sorted(_lib.SYMBOLS) == syms
    assert L.bfe2_version() == 100
    assert isinstance(L.bfe2_last_error(), bytes)


def test_bad_arguments_report_errors():
    L = _lib.lib()
    h = ctypes.c_void_p()
    conn = np.array([[0, 0]], dtype=np.int32)
    assert L.bfe2_plan_create(ctypes.byref(h), 2, 1, _lib.ptr(conn), 0, None, None) < 0
    assert b'invalid node' in L.bfe2_last_error()
    conn = np.array([[0, 1]], dtype=np.int32)
    fixed = np.array([0, 0], dtype=np.int32)
    assert L.bfe2_plan_create(ctypes.byref(h), 2, 1, _lib.ptr(conn), 2, _lib.ptr(fixed), None) < 0
    # a node that no beam touches == Structure.node_numbering_is_ok failing (Structure.py:334-341)
    assert L.bfe2_plan_create(ctypes.byref(h), 3, 1, _lib.ptr(conn), 0, None, None) < 0
    with pytest.raises(_lib.Bfe2Error):
        Plan(2, [[0, 5]], [0, 1, 2])


def test_position_maps_match_reference_contract():
    # Structure.position_in_matrix (Structure.py:379-393)
    assert planmod.position_in_matrix(nodeID=1, DOF='ux') == 0
    assert planmod.position_in_matrix(nodeID=21, DOF='rotz') == 62
    assert planmod.position_in_matrix(nodeID=8, dynam='FX') == 21
    for node in range(1, 30):
        for k, (d, f) in enumerate(zip(orc.DOFNAMES, orc.LOADNAMES)):
            assert planmod.position_in_matrix(nodeID=node, DOF=d) == orc.position_in_matrix(node, d) == (node - 1) * 3 + k
            assert planmod.position_in_matrix(nodeID=node, dynam=f) == orc.position_in_matrix(node, f)
    conn, sup = sweep.p20_topology()
    assert planmod.positions_to_eliminate(sup) == [0, 1, 2, 60, 61, 62] == orc.positions_to_eliminate(sup)


@pytest.mark.parametrize('case', sorted(util.golden_models().keys()))
def test_plan_integer_maps_bit_exact(case):
    """positions_to_eliminate / free map / block starts equal the reference's (golden) integers, and the
    band-slot scatter plan reproduces condense(compile_global_matrix(...)) EXACTLY when emulated on
    the CPU with the oracle's element matrices (same beam-order summation)."""
    model = util.golden_models()[case]
    g = util.golden(case)
    pm = util.pack(model)
    nN = len(model['nodes'])
    if 3 * nN - len(pm['fixed']) == 0:
        pytest.skip('fully supported structure: nothing to assemble')
    p = Plan.from_supports(nN, pm['conn'], model['supports'], pm['lumped'])
    assert np.array_equal(p.fixed, g['positions_to_eliminate'].astype(np.int32))
    assert p.sumdof == int(g['sumdof'])
    keep = orc.free_positions(nN, pm['fixed'])
    assert np.array_equal(p.pos_of_free, keep.astype(np.int32))
    fop = p.free_of_pos
    assert (fop[pm['fixed']] == -1).all() and np.array_equal(fop[keep], np.arange(len(keep)))
    assert np.array_equal(p.array(_lib.ARR_CONN).reshape(-1, 2) * 3, orc.block_starts(pm['conn']))
    # emulate stage_assemble with the plan
    Kg = g['Ke_global']                               # the reference's own element matrices
    sp, ct = p.array(_lib.ARR_SLOT_PTR), p.array(_lib.ARR_CONTRIB)
    n, b = p.n_free, p.hbw
    assert sp.shape[0] == (b + 1) * n + 1 and sp[-1] == ct.shape[0] == p.n_contrib
    Kb = np.zeros((b + 1, n))
    kflat = Kg.reshape(-1)
    for slot in range((b + 1) * n):
        acc = 0.0
        for q in range(sp[slot], sp[slot + 1]):
            acc += kflat[ct[q]]
        Kb[slot // n, slot % n] = acc
        e = ct[sp[slot]:sp[slot + 1]] // 36
        assert (np.diff(e) > 0).all()                   # beam-list order inside a slot
    Kc = orc.condense(np.asarray(g['K']), pm['fixed'])
    # bit-exact on the lower triangle the band stores (the reference's T K T^T is symmetric only to
    # 1 ulp for inclined members), including the bandwidth claim
    assert np.array_equal(np.tril(p.band_to_dense(Kb)), np.tril(Kc))
    # half bandwidth is tight
    i, j = np.nonzero(Kc)
    assert np.abs(i - j).max() <= b


def test_p20_plan_dimensions():
    conn, sup = sweep.p20_topology()
    p = Plan.from_supports(21, conn, sup)
    assert (p.n_nodes, p.n_elem, p.sumdof, p.n_free, p.hbw) == (21, 20, 63, 57, 5)
    assert p.smem_bytes(3) < 40 * 1024                  # >= 5 structures resident per SM
    np_ = p.array(_lib.ARR_NODE_PTR)
    ni = p.array(_lib.ARR_NODE_INC)
    assert np_[-1] == 40 and sorted(ni.tolist()) == list(range(40))


def test_compute_entry_points_fail_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('has a GPU')
    L = _lib.lib()
    assert L.bfe2_device_count() == 0
    conn, sup = sweep.p20_topology()
    p = Plan.from_supports(21, conn, sup)
    z = np.zeros(4096)
    info = np.zeros(4, dtype=np.int32)
    rc = L.bfe2_solve_fused(p.handle, 1, 0, 1, 0, _lib.ptr(z), _lib.ptr(z), None, None, None, None, None, None,
                            _lib.ptr(z), None, _lib.ptr(info), None, None)
    assert rc == -100 and b'failed' in L.bfe2_last_error()   # loud failure, no CPU fallback


def test_sweep_generator_is_deterministic():
    import torch
    a = sweep.p20_draw_params(16)
    b = sweep.p20_draw_params(16)
    assert np.array_equal(a, b)
    assert not np.array_equal(a, sweep.p20_draw_params(16, shard=1))
    lo = np.array([r[0] for r in sweep.P20_PARAM_RANGES])
    hi = np.array([r[1] for r in sweep.P20_PARAM_RANGES])
    assert (a >= lo).all() and (a <= hi).all()
    pk = sweep.p20_pack(torch.from_numpy(a))
    for k in range(3):
        pm = util.pack(sweep.p20_model_dict(a[k]))
        for key in ('coords', 'props', 'nodal_mass', 'f_nodal', 'r_local'):
            assert np.array_equal(pm[key], pk[key][k].numpy()), key


def test_pattern_layout_and_argument_checks():
    """bfe2_pattern_* (packed records) is host-side bookkeeping: widths / offsets of the records and the argument
    checks can be pinned without a GPU; the packed solve itself fails loudly without one."""
    import ctypes as C
    import torch
    from beamfe2_b200 import batch
    L = _lib.lib()
    conn, sup = sweep.p20_topology()
    p = Plan.from_supports(21, conn, sup)
    io = batch.PackedIO(p, 3, 10, True, f_index=[21, 147, 167], r_index=list(range(0, 48)))
    assert (io.w_in, io.w_out) == (42 + 80 + 42 + 3 + 48, 3 * 57 + 360 + 18 + 57 + 570 + 1)
    assert io.in_off == [0, 42, 122, 164, 167] and io.out_off == [0, 171, 531, 549, 606, 1176]
    assert io.n_fixed == 6 and io.out_shapes['reactions'] == (3, 6)
    dense = batch.PackedIO(p, 3, 10)                      # no pattern given: every load entry travels
    assert dense.w_in == 42 + 80 + 42 + 189 + 360
    static_only = batch.PackedIO(p, 2, 0, modal=False, f_index=[0], r_index=[])
    assert static_only.w_out == 2 * 57 + 240 + 12 + 1
    # views: strided windows of the record, info / sweeps share one float64 slot
    rec = torch.zeros((5, io.w_out), dtype=torch.float64)
    v = io.out_views(rec)
    v['info'][:] = torch.arange(5, dtype=torch.int32)
    v['sweeps'][:] = 7
    v['lam'][:, 0] = 3.5
    assert rec[2, io.out_off[3]] == 3.5 and v['u'].shape == (5, 3, 57) and v['phi'].shape == (5, 10, 57)
    assert rec.view(torch.int32)[4, 2 * io.out_off[5]] == 4 and rec.view(torch.int32)[4, 2 * io.out_off[5] + 1] == 7
    # argument checks
    h = C.c_void_p()
    bad = np.array([3 * 63], dtype=np.int32)             # one past the last (lc, position)
    assert L.bfe2_pattern_create(C.byref(h), p.handle, 3, 1, 10, 1, _lib.ptr(bad), 0, None) == -1
    assert b'f_index' in L.bfe2_last_error()
    assert L.bfe2_pattern_create(C.byref(h), p.handle, 0, 0, 0, 0, None, 0, None) == -1          # nothing to do
    assert L.bfe2_pattern_create(C.byref(h), p.handle, 3, 1, 58, 0, None, 0, None) == -1         # more shapes than DOFs
    assert L.bfe2_set_modal_engine(7) == -1 and L.bfe2_get_modal_engine() == 0
    if not torch.cuda.is_available():
        z = np.zeros(io.w_in + io.w_out)
        assert L.bfe2_solve_fused_packed(p.handle, io._h, 1, _lib.ptr(z), _lib.ptr(z), None) == -100
