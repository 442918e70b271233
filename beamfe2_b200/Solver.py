"""Solver operators of the drop-in API: same contract as the reference's BeamFE2/Solver.py:16-87 --
`Solver(structure).solve() -> bool` reads the Structure and writes structure.results[...] -- but the
arithmetic is the fused CUDA kernel of libbfe2 (batch of one).  There is no CPU fallback: without a
CUDA device solve() raises."""
import math

import numpy as np

from . import Results


class Solver(object):
    def __init__(self, structure=None):
        self.structure = structure

    def solve(self):
        raise NotImplementedError


MAX_JACOBI_N = 116        # BFE2_MAX_JACOBI_N (include/bfe2.h): largest n_free of the batched all-modes path
LARGE_N_MODES = 40        # modes returned for larger structures (first-k route)


def _run(structure, static, modal):
    import torch
    from . import batch
    if not torch.cuda.is_available():
        raise RuntimeError('beamfe2_b200 solvers need a CUDA device (no CPU fallback)')
    plan, pk = structure.pack()

    def dev(a):
        return torch.as_tensor(np.ascontiguousarray(a[None]), dtype=torch.float64, device='cuda')
    kw = dict(coords=dev(pk['coords']), props=dev(pk['props']))
    if static:
        kw.update(f_nodal=dev(pk['f_nodal']), r_local=dev(pk['r_local']))
    if modal:
        kw.update(nodal_mass=dev(pk['nodal_mass']))
    out = batch.solve_fused(plan, modal=modal, n_modes=plan.n_free if modal else 0, **kw)
    torch.cuda.synchronize()
    host = {k: (v[0].cpu().numpy() if v is not None else None) for k, v in out.items()}
    host['f_nodal'] = pk['f_nodal'] if static else None
    return plan, host


class LinearStaticSolver(Solver):
    def solve(self):
        st = self.structure
        if st._load_vector is None or np.count_nonzero(st._load_vector) == 0:     # Solver.py:79-80
            return False
        plan, out = _run(st, static=True, modal=False)
        if int(out['info']) != 0:
            raise np.linalg.LinAlgError('stiffness matrix is singular (pivot %d)' % int(out['info']))
        disps = np.matrix(out['u'][0]).T                                        # sumdof x 1, zeros at supports
        st.results['linear static'] = Results.LinearStaticResult(structure=st, displacements=[disps],
                                                                 endforces=out['endforces'][0],
                                                                 reactions=out['reactions'][0],
                                                                 f_nodal=out['f_nodal'][0])
        st.results['linear static'].solved = True
        return True


class ModalSolver(Solver):
    def solve(self):
        st = self.structure
        plan, pk = st.pack()
        if plan.n_free == 0:
            return False
        if np.count_nonzero(pk['props'][:, 3] * pk['props'][:, 1]) == 0 and np.count_nonzero(pk['nodal_mass']) == 0:
            return False                                                        # all-zero mass, Solver.py:34-35
        if plan.n_free > MAX_JACOBI_N:
            return self._solve_large(plan, pk)
        plan, out = _run(st, static=False, modal=True)
        info = int(out['info'])
        if info == -1:
            return False
        if info != 0:
            raise np.linalg.LinAlgError('modal solve failed (info %d)' % info)
        lam = out['lam']
        circfreq = [math.sqrt(x) for x in lam if x > 0]                         # Solver.py:42
        shapes = [np.matrix(out['phi'][k]).T for k in range(out['phi'].shape[0])]
        st.results['modal'] = Results.ModalResult(structure=st, circular_freq=circfreq, modalshapes=shapes)
        st.results['modal'].solved = True
        return True


    def _solve_large(self, plan, pk):
        """n_free beyond the batched Jacobi path: the lowest modes by subspace iteration (frame_modes.FrameModes).
        The reference would return all n_free eigenpairs (dense LAPACK, Solver.py:39); this returns the first
        `LARGE_N_MODES`."""
        import torch
        from .frame_modes import FrameModes
        if not torch.cuda.is_available():
            raise RuntimeError('beamfe2_b200 solvers need a CUDA device (no CPU fallback)')

        def dev(a):
            return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device='cuda')
        fm = FrameModes(plan, dev(pk['coords']), dev(pk['props']), dev(pk['nodal_mass']))
        out = fm.modes(n_modes=LARGE_N_MODES)
        if out.get('failed') or out['phi'] is None:
            raise np.linalg.LinAlgError('modal solve failed (subspace iteration did not produce a Ritz basis)')
        if not out['converged'] and float(out['residual'].max().item()) > 1e-8:
            raise np.linalg.LinAlgError('modal solve failed (subspace iteration stalled, scaled residual %.1e)'
                                        % float(out['residual'].max().item()))
        st = self.structure
        lam = out['lam'].cpu().numpy()
        phi = out['phi'].cpu().numpy()
        circfreq = [math.sqrt(x) for x in lam if x > 0]
        shapes = [np.matrix(phi[k]).T for k in range(phi.shape[0])]
        st.results['modal'] = Results.ModalResult(structure=st, circular_freq=circfreq, modalshapes=shapes)
        st.results['modal'].solved = True
        return True


class BucklingSolver(Solver):
    """Linear buckling.  The reference's solver (Solver.py:90-180) is unfinished -- it prints
    debug output and calls exit() -- so the contract here is the one it sketches: run the linear static analysis for
    the loads on the structure, build K_g from the axial forces it gives, solve (K + lambda K_g) phi = 0 on the
    condensed matrices and store results['buckling'] = BucklingResult(criticals = the positive load factors in
    ascending order, bucklingshapes = shapes re-expanded to all DOFs).  Returns False without loads."""

    def __init__(self, structure=None):
        super(BucklingSolver, self).__init__(structure)
        self.nr_modes = 1

    def solve(self):
        import torch
        from . import batch
        st = self.structure
        if st._load_vector is None or np.count_nonzero(st._load_vector) == 0:
            return False
        if not torch.cuda.is_available():
            raise RuntimeError('beamfe2_b200 solvers need a CUDA device (no CPU fallback)')
        plan, pk = st.pack()
        if plan.n_free == 0 or plan.n_free > MAX_JACOBI_N:
            raise NotImplementedError('buckling runs on the dense path: 1 <= n_free <= %d' % MAX_JACOBI_N)

        def dev(a):
            return torch.as_tensor(np.ascontiguousarray(a[None]), dtype=torch.float64, device='cuda')
        nm = max(1, min(int(self.nr_modes), plan.n_free))
        out = batch.buckling(plan, dev(pk['coords']), dev(pk['props']), dev(pk['f_nodal']), dev(pk['r_local']), n_modes=nm)
        torch.cuda.synchronize()
        info = int(out['info'][0].item())
        if info > 0:
            raise np.linalg.LinAlgError('stiffness matrix is singular (pivot %d)' % info)
        if info != 0:
            raise np.linalg.LinAlgError('buckling solve failed (info %d)' % info)
        lf = out['load_factor'][0].cpu().numpy()
        phi = out['phi'][0].cpu().numpy()
        keep = np.isfinite(lf)
        st.results['buckling'] = Results.BucklingResult(
            structure=st, criticals=[float(x) for x in lf[keep]],
            bucklingshapes=[np.matrix(phi[k]).T for k in range(phi.shape[0]) if keep[k]])
        st.results['buckling'].solved = True
        st.results['buckling'].axial_forces = out['axial'][0].cpu().numpy()
        return True


def solve_structures(structures, static=True, modal=True):
    """Solve MANY drop-in Structures that share one topology (same beams / node numbering / supports / mass-matrix
    flags, different coordinates, sections, masses and loads) in ONE fused launch, and fill every
    structure.results[...] exactly as the per-structure solvers do.  Returns a list of (static_ok, modal_ok).
    This is the batch axis the reference does not have (it would loop over Solver.solve())."""
    import torch
    from . import batch
    if not structures:
        return []
    packs = [st.pack() for st in structures]
    plan0, pk0 = packs[0]
    for plan, pk in packs[1:]:
        same = (plan.n_nodes == plan0.n_nodes and plan.n_elem == plan0.n_elem and np.array_equal(plan.conn, plan0.conn)
                and np.array_equal(plan.fixed, plan0.fixed) and np.array_equal(plan.lumped, plan0.lumped))
        if not same:
            raise ValueError('solve_structures: all structures must share one topology plan')

    def dev(key):
        return torch.as_tensor(np.ascontiguousarray(np.stack([pk[key] for _, pk in packs])), dtype=torch.float64,
                               device='cuda')
    has_load = [st._load_vector is not None and np.count_nonzero(st._load_vector) != 0 for st in structures]
    has_mass = [np.count_nonzero(pk['props'][:, 3] * pk['props'][:, 1]) != 0 or np.count_nonzero(pk['nodal_mass']) != 0
                for _, pk in packs]
    do_static = static and any(has_load)
    do_modal = modal and plan0.n_free > 0 and any(has_mass)
    if not (do_static or do_modal):
        return [(False, False)] * len(structures)
    out = batch.solve_fused(plan0, dev('coords'), dev('props'), dev('nodal_mass') if do_modal else None,
                            dev('f_nodal') if do_static else None, dev('r_local') if do_static else None,
                            modal=do_modal, n_modes=plan0.n_free if do_modal else 0)
    torch.cuda.synchronize()
    host = {k: (v.cpu().numpy() if v is not None else None) for k, v in out.items()}
    flags = []
    for b, st in enumerate(structures):
        info = int(host['info'][b])
        s_ok = m_ok = False
        # info -1 / -2 invalidate only the modal outputs (include/bfe2.h); the static ones are checked as such
        if do_static and has_load[b] and info <= 0 and np.isfinite(host['u'][b, 0]).all():
            st.results['linear static'] = Results.LinearStaticResult(
                structure=st, displacements=[np.matrix(host['u'][b, 0]).T], endforces=host['endforces'][b, 0],
                reactions=host['reactions'][b, 0], f_nodal=packs[b][1]['f_nodal'][0])
            st.results['linear static'].solved = s_ok = True
        if do_modal and has_mass[b] and info == 0:
            lam = host['lam'][b]
            st.results['modal'] = Results.ModalResult(
                structure=st, circular_freq=[math.sqrt(x) for x in lam if x > 0],
                modalshapes=[np.matrix(host['phi'][b, k]).T for k in range(host['phi'].shape[1])])
            st.results['modal'].solved = m_ok = True
        flags.append((s_ok, m_ok))
    return flags
