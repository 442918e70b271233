"""Shared helpers for the test-suite: golden fixtures, packing, comparison metrics."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, 'tests', 'golden')

_cache = {}


def golden_models():
    if 'models' not in _cache:
        with open(os.path.join(GOLD, 'ref_cases.json')) as f:
            raw = json.load(f)
        for m in raw.values():
            m['supports'] = {int(k): v for k, v in m['supports'].items()}
            m['nodes'] = [tuple(p) for p in m['nodes']]
            m['masses'] = [(int(a), b) for a, b in m.get('masses', [])]
            for lc in m.get('load_cases', []):
                lc['nodal'] = [(int(a), b) for a, b in lc.get('nodal', [])]
                lc['beam'] = [(int(a), b, c) for a, b, c in lc.get('beam', [])]
        _cache['models'] = raw
    return _cache['models']


def golden_arrays():
    if 'arrays' not in _cache:
        _cache['arrays'] = dict(np.load(os.path.join(GOLD, 'ref_cases.npz')))
    return _cache['arrays']


def golden(case):
    pre = case + '/'
    return {k[len(pre):]: v for k, v in golden_arrays().items() if k.startswith(pre)}


def p20_truth():
    if 'truth' not in _cache:
        _cache['truth'] = dict(np.load(os.path.join(GOLD, 'p20_truth.npz')))
    return _cache['truth']


def pack(model):
    from oracle import ref_harness as rh
    return rh.pack_model(model)


def relmax(a, b):
    """max |a-b| / max |b|  (relative to the largest entry of the reference array)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / (den if den > 0 else 1.0))


def align_sign(phi, ref):
    """Flip each mode of phi [..., m, n] to the sign of ref (modes are defined up to sign)."""
    s = np.sign(np.sum(phi * ref, axis=-1, keepdims=True))
    s[s == 0] = 1
    return phi * s


def mode_err(phi, ref):
    """Per-mode max abs error relative to the mode's largest entry, after sign alignment."""
    p = align_sign(phi, ref)
    return np.abs(p - ref).max(axis=-1) / np.abs(ref).max(axis=-1)
