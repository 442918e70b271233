"""CPU comparators for the large-single-structure path (config 4)  --  TEST INFRASTRUCTURE.

The reference itself cannot run this size (helpers.py:17 allocates sumdof^2 doubles); SURVEY.md 8d names
scipy.linalg.solveh_banded and scipy.sparse.linalg.eigsh(K, M=M, sigma=0) as the CPU comparators, with a
residual + closed-form parity criterion (cond(K) grows like n^4).  Element matrices come from the pinned
oracle (oracle/beamfe2_oracle.py)."""
import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import beamfe2_oracle as orc


def chain_banded(coords, props, fixed):
    """Condensed K, M of a chain (beam k joins nodes k, k+1) in LOWER band form [6, n_free] and as
    scipy.sparse CSC.  Assembly order = beam order (helpers.py:19-39)."""
    nE = props.shape[0]
    conn = np.stack([np.arange(nE), np.arange(1, nE + 1)], axis=1)
    Kg, Mg = orc.element_km_global(coords, conn, props)
    n = 3 * (nE + 1)
    rows, cols, kv, mv = [], [], [], []
    for a in range(6):
        for b in range(6):
            r = 3 * conn[:, a // 3] + a % 3
            c = 3 * conn[:, b // 3] + b % 3
            rows.append(r); cols.append(c); kv.append(Kg[:, a, b]); mv.append(Mg[:, a, b])
    rows, cols = np.concatenate(rows), np.concatenate(cols)
    K = sp.coo_matrix((np.concatenate(kv), (rows, cols)), shape=(n, n)).tocsr()
    M = sp.coo_matrix((np.concatenate(mv), (rows, cols)), shape=(n, n)).tocsr()
    keep = orc.free_positions(nE + 1, fixed)
    K = K[keep][:, keep].tocsc()
    M = M[keep][:, keep].tocsc()
    return K, M, keep


def to_lower_band(A, hbw=5):
    n = A.shape[0]
    ab = np.zeros((hbw + 1, n))
    for d in range(hbw + 1):
        ab[d, :n - d] = A.diagonal(-d)
    return ab


def static(coords, props, fixed, f):
    """solveh_banded on the condensed system; returns u [sumdof] with zeros at supports."""
    K, _, keep = chain_banded(coords, props, fixed)
    x = sla.solveh_banded(to_lower_band(K), f[keep], lower=True)
    u = np.zeros_like(f)
    u[keep] = x
    return u


def modal(coords, props, fixed, k):
    """eigsh(K, M=M, sigma=0) -- lowest k eigenpairs, shapes expanded to sumdof, M-normalised."""
    K, M, keep = chain_banded(coords, props, fixed)
    lam, vec = spla.eigsh(K, k=k, M=M, sigma=0, which='LM')
    order = np.argsort(lam)
    lam, vec = lam[order], vec[:, order]
    phi = np.zeros((k, 3 * (props.shape[0] + 1)))
    phi[:, keep] = vec.T
    return lam, phi


def cantilever_closed_form(P, L, E, I, rho, A, n_modes):
    """Tip deflection / rotation under a tip force and the Euler-Bernoulli bending frequencies
    omega_k = (beta_k L)^2 sqrt(EI / (rho A L^4))."""
    bl = [1.8751040687119611, 4.694091132974175, 7.854757438237613, 10.995540734875467, 14.13716839104647]
    bl += [(2 * k - 1) * np.pi / 2 for k in range(6, n_modes + 1)]
    w = np.array(bl[:n_modes]) ** 2 * np.sqrt(E * I / (rho * A * L ** 4))
    return P * L ** 3 / (3 * E * I), P * L ** 2 / (2 * E * I), w


def cantilever_beta_l(k):
    """k-th root of cos(x) cosh(x) = -1 (clamped-free Euler-Bernoulli beam), Newton on cos(x) + 1/cosh(x)."""
    x = (2 * k - 1) * np.pi / 2
    if k == 1:
        x = 1.875
    for _ in range(50):
        g = np.cos(x) + 1.0 / np.cosh(x)
        dg = -np.sin(x) - np.sinh(x) / np.cosh(x) ** 2
        dx = g / dg
        x -= dx
        if abs(dx) < 1e-16 * x:
            break
    return float(x)


def cantilever_spectrum(L, E, I, rho, A, n_modes):
    """The n_modes lowest circular frequencies of the continuous clamped-free member in the plane: the union of the
    bending family (beta_k L)^2 sqrt(EI / rho A L^4) and the axial family (2k-1) pi/2 sqrt(E/rho) / L, sorted."""
    bend = np.array([cantilever_beta_l(k) for k in range(1, n_modes + 1)]) ** 2 * np.sqrt(E * I / (rho * A * L ** 4))
    axial = (2 * np.arange(1, n_modes + 1) - 1) * np.pi / 2 * np.sqrt(E / rho) / L
    return np.sort(np.concatenate([bend, axial]))[:n_modes]
