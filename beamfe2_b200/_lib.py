"""ctypes binding of libbfe2.so (include/bfe2.h).  Fails loudly when the library is missing:
there is no CPU fallback and no alternative backend."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('BFE2_LIB', os.path.join(_HERE, 'libbfe2.so'))   # BFE2_LIB: A/B builds only

# every symbol include/bfe2.h declares (tests check the library exports exactly these)
SYMBOLS = (
    'bfe2_version', 'bfe2_last_error', 'bfe2_device_count',
    'bfe2_plan_create', 'bfe2_plan_create_ex', 'bfe2_plan_destroy', 'bfe2_plan_dims', 'bfe2_plan_get_array',
    'bfe2_plan_smem_bytes', 'bfe2_workspace_bytes',
    'bfe2_element_km', 'bfe2_assemble_banded', 'bfe2_static_solve', 'bfe2_modal_solve',
    'bfe2_solve_fused', 'bfe2_set_modal_engine', 'bfe2_get_modal_engine', 'bfe2_static_tile_launches',
    'bfe2_pattern_create', 'bfe2_pattern_destroy', 'bfe2_pattern_layout', 'bfe2_solve_fused_packed',
    'bfe2_plan_create_dense', 'bfe2_sym_eig_dense',
    'bfe2_beam_create', 'bfe2_beam_create_ex', 'bfe2_beam_destroy', 'bfe2_beam_dims', 'bfe2_beam_solve',
    'bfe2_beam_solve_dd', 'bfe2_beam_matvec', 'bfe2_beam_residual', 'bfe2_beam_export_blocks',
    'bfe2_gram_workspace_bytes', 'bfe2_gram64', 'bfe2_update64',
    'bfe2_internal_actions', 'bfe2_internal_actions_ex', 'bfe2_modal_participation', 'bfe2_assemble_geometric', 'bfe2_band_matvec',
    'bfe2_band_matvec_batched', 'bfe2_gram64_batched', 'bfe2_update64_batched',
)

ARR_CONN, ARR_FREE_OF_POS, ARR_POS_OF_FREE, ARR_FIXED, ARR_SLOT_PTR, ARR_CONTRIB, ARR_NODE_PTR, \
    ARR_NODE_INC = range(8)

_lib = None


class Bfe2Error(RuntimeError):
    pass


def build(verbose=False):
    """Compile libbfe2.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    import subprocess
    csrc = os.path.join(_HERE, 'csrc')
    r = subprocess.run(['make', '-j4', '-C', csrc], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout)
        print(r.stderr)
    if r.returncode != 0:
        raise Bfe2Error('building libbfe2.so failed')
    return LIB_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Bfe2Error('%s not found -- run `python -c "import __graft_entry__ as g; g.build()"` '
                        '(there is no CPU fallback)' % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t
    L.bfe2_version.restype = C.c_int
    L.bfe2_last_error.restype = C.c_char_p
    L.bfe2_device_count.restype = C.c_int
    L.bfe2_plan_create.restype = C.c_int
    L.bfe2_plan_create.argtypes = [C.POINTER(vp), i32, i32, vp, i32, vp, vp]
    L.bfe2_plan_create_ex.restype = C.c_int
    L.bfe2_plan_create_ex.argtypes = [C.POINTER(vp), i32, i32, vp, i32, vp, vp, i32]
    L.bfe2_plan_destroy.restype = None
    L.bfe2_plan_destroy.argtypes = [vp]
    L.bfe2_plan_dims.restype = C.c_int
    L.bfe2_plan_dims.argtypes = [vp] + [C.POINTER(i32)] * 6
    L.bfe2_plan_get_array.restype = C.c_int
    L.bfe2_plan_get_array.argtypes = [vp, C.c_int, vp, i64]
    L.bfe2_plan_smem_bytes.restype = sz
    L.bfe2_plan_smem_bytes.argtypes = [vp, i32]
    L.bfe2_workspace_bytes.restype = sz
    L.bfe2_workspace_bytes.argtypes = [vp, i64]
    L.bfe2_element_km.restype = C.c_int
    L.bfe2_element_km.argtypes = [i64] + [vp] * 12
    L.bfe2_assemble_banded.restype = C.c_int
    L.bfe2_assemble_banded.argtypes = [vp, i64] + [vp] * 6
    L.bfe2_static_solve.restype = C.c_int
    L.bfe2_static_solve.argtypes = [vp, i64, i32] + [vp] * 10
    L.bfe2_modal_solve.restype = C.c_int
    L.bfe2_modal_solve.argtypes = [vp, i64, i32] + [vp] * 7
    L.bfe2_solve_fused.restype = C.c_int
    L.bfe2_solve_fused.argtypes = [vp, i64, i32, i32, i32] + [vp] * 13
    L.bfe2_set_modal_engine.restype = C.c_int
    L.bfe2_set_modal_engine.argtypes = [i32]
    L.bfe2_get_modal_engine.restype = C.c_int
    L.bfe2_static_tile_launches.restype = C.c_int64
    L.bfe2_static_tile_launches.argtypes = []
    L.bfe2_pattern_create.restype = C.c_int
    L.bfe2_pattern_create.argtypes = [C.POINTER(vp), vp, i32, i32, i32, i32, vp, i32, vp]
    L.bfe2_pattern_destroy.restype = None
    L.bfe2_pattern_destroy.argtypes = [vp]
    L.bfe2_pattern_layout.restype = C.c_int
    L.bfe2_pattern_layout.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), vp, vp]
    L.bfe2_solve_fused_packed.restype = C.c_int
    L.bfe2_solve_fused_packed.argtypes = [vp, vp, i64, vp, vp, vp]
    L.bfe2_plan_create_dense.restype = C.c_int
    L.bfe2_plan_create_dense.argtypes = [C.POINTER(vp), i32]
    L.bfe2_sym_eig_dense.restype = C.c_int
    L.bfe2_sym_eig_dense.argtypes = [vp, i64] + [vp] * 8
    L.bfe2_beam_create.restype = C.c_int
    L.bfe2_beam_create.argtypes = [C.POINTER(vp), i64, vp, vp, vp, i32, vp]
    L.bfe2_beam_create_ex.restype = C.c_int
    L.bfe2_beam_create_ex.argtypes = [C.POINTER(vp), i64, vp, vp, vp, i32, i32, vp]
    L.bfe2_beam_solve_dd.restype = C.c_int
    L.bfe2_beam_solve_dd.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    L.bfe2_beam_residual.restype = C.c_int
    L.bfe2_beam_residual.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    L.bfe2_beam_export_blocks.restype = C.c_int
    L.bfe2_beam_export_blocks.argtypes = [vp, i32, vp, vp]
    L.bfe2_beam_destroy.restype = None
    L.bfe2_beam_destroy.argtypes = [vp]
    L.bfe2_beam_dims.restype = C.c_int
    L.bfe2_beam_dims.argtypes = [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i32)]
    L.bfe2_beam_solve.restype = C.c_int
    L.bfe2_beam_solve.argtypes = [vp, i32, vp, vp, vp]
    L.bfe2_beam_matvec.restype = C.c_int
    L.bfe2_beam_matvec.argtypes = [vp, i32, i32, vp, vp, vp]
    L.bfe2_gram_workspace_bytes.restype = sz
    L.bfe2_gram_workspace_bytes.argtypes = [i64]
    L.bfe2_gram64.restype = C.c_int
    L.bfe2_gram64.argtypes = [i64, vp, vp, vp, vp, vp]
    L.bfe2_update64.restype = C.c_int
    L.bfe2_update64.argtypes = [i64, vp, vp, vp, vp]
    L.bfe2_internal_actions.restype = C.c_int
    L.bfe2_internal_actions.argtypes = [vp, i64, i32, i32] + [vp] * 7
    L.bfe2_internal_actions_ex.restype = C.c_int
    L.bfe2_internal_actions_ex.argtypes = [vp, i64, i32, i32] + [vp] * 4 + [i32] + [vp] * 4
    L.bfe2_band_matvec_batched.restype = C.c_int
    L.bfe2_band_matvec_batched.argtypes = [i32, i32, i32, i64, vp, vp, vp, vp]
    L.bfe2_gram64_batched.restype = C.c_int
    L.bfe2_gram64_batched.argtypes = [i64, i64, vp, vp, vp, vp]
    L.bfe2_update64_batched.restype = C.c_int
    L.bfe2_update64_batched.argtypes = [i64, i64, vp, vp, vp, vp]
    L.bfe2_band_matvec.restype = C.c_int
    L.bfe2_band_matvec.argtypes = [i32, i32, i32, vp, vp, vp, vp]
    L.bfe2_assemble_geometric.restype = C.c_int
    L.bfe2_assemble_geometric.argtypes = [vp, i64, vp, vp, vp, vp]
    L.bfe2_modal_participation.restype = C.c_int
    L.bfe2_modal_participation.argtypes = [vp, i64, i32] + [vp] * 4
    _lib = L
    return L


def check(rc):
    if rc < 0:
        raise Bfe2Error('libbfe2: %s (code %d)' % (lib().bfe2_last_error().decode(), rc))
    return rc


def ptr(t):
    """Device (or host, for plan creation) pointer of a tensor / ndarray, None -> NULL."""
    if t is None:
        return None
    if hasattr(t, 'data_ptr'):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(t.ctypes.data)
