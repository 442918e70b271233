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


def mode_err_clustered(phi, ref, lam, rel_gap=1e-3):
    """Like mode_err, but modes whose eigenvalues are closer than rel_gap (relative) are compared as a SUBSPACE
    (SURVEY.md 7.3-1): inside such a cluster the individual vectors are defined only up to a rotation that depends on
    rounding, so each computed mode is measured by its distance from the span of the reference cluster (least-squares
    fit; relative to the mode's largest entry).  phi, ref: [m, n]; lam: [>= m + 1] ascending eigenvalues (the one
    after the last mode decides whether that mode is in a cluster)."""
    phi, ref, lam = np.asarray(phi), np.asarray(ref), np.asarray(lam)
    m = ref.shape[0]
    err = np.zeros(m)
    close = np.diff(lam[:m + 1]) / lam[1:m + 1] < rel_gap            # close[k]: modes k and k+1 are a pair
    k = 0
    while k < m:
        j = k
        while j < m and close[j]:
            j += 1
        j = min(j, m - 1)
        if j == k and not (k == m - 1 and close[k]):
            err[k] = mode_err(phi[k], ref[k])
        elif j == k:                                               # last mode pairs with one outside the stored set
            err[k] = np.nan
        else:
            R = ref[k:j + 1]
            for i in range(k, j + 1):
                c = np.linalg.lstsq(R.T, phi[i], rcond=None)[0]
                err[i] = np.abs(phi[i] - R.T @ c).max() / np.abs(phi[i]).max()
                if j == m - 1 and close[j]:                        # the cluster continues beyond the stored modes
                    err[i] = np.nan
        k = j + 1
    return err


def p20_truth_batch(device=None):
    """(truth dict, packed tensors) of the truth fixture's own parameter rows."""
    import torch
    from beamfe2_b200 import sweep
    t = p20_truth()
    P = torch.from_numpy(t['params'])
    return t, sweep.p20_pack(P if device is None else P.to(device))


def refined_eigenvalues(Kc, Mc, V):
    """Rayleigh quotients v^T K v / v^T M v of (scipy's) eigenvectors V [B, m, n] in extended precision: second-order
    accurate in the vector error, i.e. far tighter than LAPACK's own eigenvalues (absolute accuracy eps * lam_max)."""
    K, M, V = Kc.astype(np.longdouble), Mc.astype(np.longdouble), V.astype(np.longdouble)
    rq = np.einsum('bmi,bij,bmj->bm', V, K, V) / np.einsum('bmi,bij,bmj->bm', V, M, V)
    return rq.astype(np.float64)


def truth_eigenvalues(Kc, Mc, dps=60):
    """All eigenvalues (ascending, as floats) of the symmetric-definite pencil (Kc, Mc) from a `dps`-digit mpmath
    solution -- for small pencils whose grading (lam_max / lam_min up to 1e20 with the reference's lumped mass) puts
    them beyond what any double-precision comparator resolves."""
    import mpmath as mp
    old = mp.mp.dps
    mp.mp.dps = dps
    try:
        K, M = mp.matrix(np.asarray(Kc).tolist()), mp.matrix(np.asarray(Mc).tolist())
        Li = mp.inverse(mp.cholesky(M))
        Cm = Li * K * Li.T
        E = mp.eigsy((Cm + Cm.T) / 2, eigvals_only=True)
        return np.array(sorted(float(e) for e in E))
    finally:
        mp.mp.dps = old
