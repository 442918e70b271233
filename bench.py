#!/usr/bin/env python
"""bench.py -- headline benchmark: FP64 static+modal structures solved per second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

Workload (BASELINE.json configs[2] / SURVEY.md 8d "config 3"): B = 1e5 variants PER GPU of the
20-element portal frame P20 (21 nodes, 63 DOFs, 57 free, half bandwidth 5), 3 load cases, static
(displacements, end forces, reactions) + modal (all 57 eigenvalues, first 10 M-normalised modes).
One "step" = one pass of the fused hot path over that batch.  Weak scaling: every rank owns its own
B variants (seeded per rank), no data-path collective; one final NCCL gather of summaries after the
timed region.

Prints ONE JSON line on rank 0 (see the field list in the task contract):
  value      structures/s with inputs resident in HBM (CUDA events around K steps, max over ranks)
  e2e        same metric through HostPipeline: pinned host inputs -> H2D -> kernel -> D2H of all results
  roofline   the fused kernel against the HBM roof with SURVEY 8d's algorithmic 12.7 KB/structure,
             plus its FP64 fraction (2.11 MFLOP/structure over a DFMA peak measured here)
  cpu_baseline  the CPU oracle (numpy/scipy restatement of the reference path) on the host cores
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'FP64 static+modal structures solved/s'
UNIT = 'structures/s'
ALGO_BYTES_PER_STRUCT = 12.7e3      # SURVEY.md 8(d), fused end-to-end
ALGO_FLOP_PER_STRUCT = 2.11e6       # 11.33 n^3 (n = 57) + static
N_MODES = 10
N_LC = 3


# ------------------------------------------------------------------------------------ CPU arm
def _cpu_worker(args):
    """One host process: oracle static+modal on `count` P20 variants; returns seconds of solve time."""
    shard, count = args
    os.environ['OMP_NUM_THREADS'] = '1'
    os.environ['OPENBLAS_NUM_THREADS'] = '1'
    import numpy as np  # noqa: F401
    import torch
    torch.set_num_threads(1)
    from beamfe2_b200 import sweep
    from oracle import beamfe2_oracle as orc
    P = sweep.p20_draw_params(count, shard=1000 + shard)
    pk = {k: v.numpy() for k, v in sweep.p20_pack(torch.from_numpy(P)).items()}
    conn, sup = sweep.p20_topology()
    fixed = orc.positions_to_eliminate(sup)
    t0 = time.perf_counter()
    orc.solve_batch(conn, fixed, pk['coords'], pk['props'], pk['nodal_mass'], pk['f_nodal'], pk['r_local'],
                    n_modes=N_MODES)
    return time.perf_counter() - t0


def cpu_oracle_throughput(per_worker: int, workers: int):
    """structures/s of the CPU oracle with `workers` single-threaded processes (one per host core)."""
    import multiprocessing as mp
    ctx = mp.get_context('spawn')
    with ctx.Pool(workers) as pool:
        pool.map(_cpu_worker, [(w, 8) for w in range(workers)])          # warm-up: imports, BLAS init
        t0 = time.perf_counter()
        pool.map(_cpu_worker, [(w, per_worker) for w in range(workers)])
        wall = time.perf_counter() - t0
    return per_worker * workers / wall, wall


def _ref_worker(args):
    """One host process: the REAL reference (baseline/_ref) on `count` P20 variants through its own object path
    (SURVEY.md 8d arm (A)); returns seconds, or a string when the reference is not installed."""
    shard, count = args
    os.environ['OMP_NUM_THREADS'] = '1'
    os.environ['OPENBLAS_NUM_THREADS'] = '1'
    from beamfe2_b200 import sweep
    from oracle import ref_harness as rh
    if not rh.reference_available():
        return 'reference not installed under baseline/_ref (run __graft_entry__.build() where the checkout exists)'
    P = sweep.p20_draw_params(count, shard=2000 + shard)
    models = [sweep.p20_model_dict(P[k]) for k in range(count)]
    rh.object_path_seconds(models[:1])                                   # imports, first-call costs
    return rh.object_path_seconds(models)


def cpu_reference_throughput(per_worker: int, workers: int):
    """structures/s of the unmodified reference with `workers` single-threaded processes; None if unavailable."""
    import multiprocessing as mp
    ctx = mp.get_context('spawn')
    with ctx.Pool(workers) as pool:
        t0 = time.perf_counter()
        secs = pool.map(_ref_worker, [(w, per_worker) for w in range(workers)])
        wall = time.perf_counter() - t0
    if any(isinstance(x, str) for x in secs):
        return None, [x for x in secs if isinstance(x, str)][0]
    # the workers run side by side: throughput = total structures / the slowest worker's solve time
    return per_worker * workers / max(secs), wall


def cpu_arms(port_per_worker: int, ref_total: int = 256):
    """All four CPU arms of SURVEY.md 8(d): the real reference (A) and the numpy/scipy oracle port (B), each on one
    core and on every host core."""
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 64))
    arms = []
    for w in (1, workers):
        per = max(8, ref_total // w) if w > 1 else min(ref_total, 200)
        v, info = cpu_reference_throughput(per, w)
        a = {'kind': 'reference', 'cores': w, 'unit': UNIT,
             'sample': '%d P20 variants through the unmodified BeamFE2 objects from baseline/_ref: Node / HermitianBeam2D / '
                       'Structure, 3 x (clear + load + linear static solve + reaction_forces_asvector) + modal solve' % (per * w)}
        if v is None:
            a['unavailable'] = info
        else:
            a['value'] = v
        arms.append(a)
    for w in (1, workers):
        per = port_per_worker if w > 1 else min(port_per_worker, 4096)
        v, wall = cpu_oracle_throughput(per, w)
        arms.append({'kind': 'port', 'cores': w, 'unit': UNIT, 'value': v,
                     'sample': '%d P20 variants (%d per worker x %d single-threaded workers, %.1f s wall), numpy/scipy oracle '
                               'port: dense assembly + numpy.linalg.solve + scipy.linalg.eigh(K, M)' % (per * w, per, w, wall)})
    return arms, workers


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 64))
    per_worker = args.cpu_sample
    vals = []
    for _ in range(args.warmup):
        cpu_oracle_throughput(max(8, per_worker // 8), workers)
    t_all = 0.0
    for _ in range(args.steps):
        v, wall = cpu_oracle_throughput(per_worker, workers)
        vals.append(v)
        t_all += wall
    value = sum(vals) / len(vals)
    sample = '%d P20 variants per step (%d per worker x %d single-threaded workers), numpy/scipy oracle port of ' \
             'the reference path: dense assembly, numpy.linalg.solve, scipy.linalg.eigh(K, M)' % (
                 per_worker * workers, per_worker, workers)
    # the unmodified reference through its own objects (arm (A)), one core and all cores, beside the port
    arms = []
    for w in (1, workers):
        per = max(8, args.ref_sample // w) if w > 1 else min(args.ref_sample, 200)
        rv, info = cpu_reference_throughput(per, w)
        a = {'kind': 'reference', 'cores': w, 'unit': UNIT,
             'sample': '%d P20 variants through the unmodified BeamFE2 objects (baseline/_ref): Node / HermitianBeam2D / '
                       'Structure, 3 x (clear + load + linear static solve + reaction_forces_asvector) + modal solve' % (per * w)}
        a.update({'unavailable': info} if rv is None else {'value': rv})
        arms.append(a)
    v1, _ = cpu_oracle_throughput(min(per_worker, 4096), 1)
    arms.append({'kind': 'port', 'cores': 1, 'unit': UNIT, 'value': v1, 'sample': 'the port on one core'})
    arms.append({'kind': 'port', 'cores': workers, 'unit': UNIT, 'value': value, 'sample': sample})
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * t_all / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args, per_gpu=per_worker * workers),
        # headline = the FASTER of the two CPU arms (the vectorised numpy/scipy port, ~50x the reference's own
        # object path), so the driver's GPU/CPU ratio is the conservative one; every arm is listed in `arms`
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': workers, 'kind': 'port', 'sample': sample, 'arms': arms},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, per_gpu):
    return {'workload': 'config 3: P20 portal-frame variants (20 elements, 63 DOFs, 57 free, hbw 5), 3 load cases, '
                        'static (u, end forces, reactions) + modal (57 eigenvalues, first 10 modes)',
            'structures_per_gpu_per_step': int(per_gpu), 'load_cases': N_LC, 'n_modes': N_MODES,
            'seed': 20261019, 'parallelism': 'batch slices, one rank per GPU, no data-path collective',
            'l2': 'inputs+outputs per step (%.0f MB) exceed the 126 MB L2' % (per_gpu * 17.1e3 / 1e6)}


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples nvidia-smi SM clocks / throttle reasons of one GPU while the timed region runs."""
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,' \
        'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,' \
        'clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            parts = [p.strip() for p in r.split(',')]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ------------------------------------------------------------------------------------ GPU arm
def measure_fp64_peak(torch, dev):
    import ctypes
    lib = ctypes.CDLL(os.path.join(ROOT, 'beamfe2_b200', 'libbfe2_peaks.so'))
    lib.bfe2_peaks_dfma.restype = ctypes.c_longlong
    lib.bfe2_peaks_dfma.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    blocks = sms * 8
    out = torch.empty(blocks * 256, dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    lib.bfe2_peaks_dfma(out.data_ptr(), blocks, 200, st)
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = lib.bfe2_peaks_dfma(out.data_ptr(), blocks, 4000, st)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, 2.0 * n / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


def run_gpu_arm(args):
    import numpy as np
    import torch
    from beamfe2_b200 import batch, dist as bdist, sweep
    from beamfe2_b200.plan import Plan

    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device -- the product path has no CPU fallback')
    rank, local, world = bdist.init_from_env()
    if world != args.gpus:
        if rank == 0:
            print('bench.py: WORLD_SIZE=%d but --gpus %d; using WORLD_SIZE' % (world, args.gpus), file=sys.stderr)
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    numa = bdist.bind_to_gpu_numa_node(local) if world > 1 else None
    B = args.batch
    batch.set_modal_engine(args.engine)
    conn, sup = sweep.p20_topology()
    plan = Plan.from_supports(21, conn, sup)
    P = sweep.p20_draw_params(B, shard=None if world == 1 else rank)
    pk = sweep.p20_pack(torch.from_numpy(P).to(dev))
    out = {}
    res = batch.solve_fused(plan, pk['coords'], pk['props'], pk['nodal_mass'], pk['f_nodal'], pk['r_local'],
                            n_modes=N_MODES, out=out)
    out = {k: v for k, v in res.items() if v is not None}
    torch.cuda.synchronize()
    if not bool((res['info'] == 0).all()):
        raise SystemExit('bench.py: solver reported failures (info != 0)')

    def step():
        batch.solve_fused(plan, pk['coords'], pk['props'], pk['nodal_mass'], pk['f_nodal'], pk['r_local'],
                          n_modes=N_MODES, out=out)

    for _ in range(max(3, args.warmup)):
        step()
    sampler = ClockSampler(local)
    bdist.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    bdist.barrier()
    ms = bdist.max_over_ranks(e0.elapsed_time(e1), dev)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms / args.steps
    value = B * world / (ms_per_step * 1e-3)
    kernel_ms = ms_per_step                         # one fused launch per step

    # ---- e2e: host buffers through the public HostPipeline API (packed records, compact load / result layouts)
    io = batch.PackedIO.from_dense(plan, pk['f_nodal'], pk['r_local'], n_modes=N_MODES)
    pipe = batch.HostPipeline(plan, N_LC, N_MODES, chunk=args.chunk, n_streams=3, device=dev, io=io)
    Be = min(B, args.e2e_batch)                 # bounds the pinned host memory of a rank (11.1 KB per structure)
    h_in, h_out = pipe.alloc_host(Be)
    io.pack_inputs(h_in, *[pk[k][:Be].cpu() for k in ('coords', 'props', 'nodal_mass', 'f_nodal', 'r_local')])
    for _ in range(2):
        pipe.run(h_in, h_out)
    bdist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(e2e_steps):
        pipe.run(h_in, h_out)
    f1.record()
    torch.cuda.synchronize()
    e2e_ms = bdist.max_over_ranks(f0.elapsed_time(f1), dev) / e2e_steps
    e2e_wall = (time.perf_counter() - t0) / e2e_steps
    bi, bo = pipe.bytes_per_structure()
    e2e_value = Be * world / (e2e_ms * 1e-3)
    hv = io.out_views(h_out)
    pos = torch.as_tensor(plan.pos_of_free.astype('int64'))
    fixed = torch.as_tensor(plan.fixed.astype('int64'))
    e2e_ok = bool(torch.equal(hv['lam'], out['lam'][:Be].cpu()) and torch.equal(hv['u'], out['u'][:Be].cpu()[:, :, pos])
                  and torch.equal(hv['phi'], out['phi'][:Be].cpu()[:, :, pos])
                  and torch.equal(hv['endforces'], out['endforces'][:Be].cpu())
                  and torch.equal(hv['reactions'], out['reactions'][:Be].cpu()[:, :, fixed])
                  and torch.equal(hv['info'], out['info'][:Be].cpu()))

    # ---- final gather of summaries (off the timed path)
    table = bdist.gather_summaries(bdist.summarize(out['u'], out['lam'], out['info']), B * world)
    n_bad = int((table[:, 0] != 0).sum().item())

    if world > 1:
        import torch.distributed as tdist
        tdist.barrier()
        tdist.destroy_process_group()
    if rank != 0:
        return 0
    peaks = {}
    pj = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    hbm_peak, peak_src = 6650.0, 'fallback'
    if os.path.exists(pj):
        peaks = json.load(open(pj))
        hbm_peak, peak_src = float(peaks.get('hbm_gbs', hbm_peak)), 'measured (MEASURED_PEAKS.json)'
    fp64_peak = measure_fp64_peak(torch, dev)
    achieved = ALGO_BYTES_PER_STRUCT * B / (kernel_ms * 1e-3) / 1e9
    fp64_ach = ALGO_FLOP_PER_STRUCT * B / (kernel_ms * 1e-3) / 1e12
    traffic = None
    tj = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tj):
        try:
            traffic = json.load(open(tj)).get('fused_bytes_per_structure') * B
        except Exception:
            traffic = None
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(3, args.warmup),
        'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic', 'config': workload_config(args, per_gpu=B),
        'clocks': clocks,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': bi * Be, 'd2h_bytes_per_step': bo * Be,
                'structures_per_gpu_per_step': Be,
                'ms_per_step': e2e_ms, 'wall_ms_per_step': e2e_wall * 1e3, 'chunk': args.chunk,
                'copies_per_chunk': 2, 'layout': 'packed records: non-zero load entries in (%d of %d doubles), u / phi '
                'on the free DOFs and reactions at the supports out' % (io.w_in, 713),
                'matches_device_run': e2e_ok, 'numa_node': numa},
        'gpu_launches': args.steps,
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': hbm_peak, 'unit': 'GB/s',
                     'frac': achieved / hbm_peak, 'traffic': traffic, 'peak_source': peak_src,
                     'kernel': 'k_solve<29,2,64,5,true,true> (fused K1->K4)', 'kernel_ms': kernel_ms,
                     'algorithmic_bytes_per_structure': ALGO_BYTES_PER_STRUCT,
                     'note': 'the fused kernel is FP64/shared-memory bound, not HBM bound (SURVEY 8d); see fp64',
                     'fp64': {'achieved_tflops': fp64_ach, 'peak_tflops_measured_dfma': fp64_peak,
                              'frac': fp64_ach / fp64_peak if fp64_peak else None,
                              'algorithmic_flop_per_structure': ALGO_FLOP_PER_STRUCT}},
        'failed_structures': n_bad,
        'mean_sweeps': float(out['sweeps'].double().mean().item()),
    }
    if not args.no_cpu:
        # rank 0 only, after the timed regions and after the process group is gone; at N > 1 a smaller sample
        arms, workers = cpu_arms(args.cpu_sample if world == 1 else max(1024, args.cpu_sample // 8),
                                 args.ref_sample if world == 1 else max(64, args.ref_sample // 4))
        port = [a for a in arms if a['kind'] == 'port' and a['cores'] == workers][-1]
        line['cpu_baseline'] = {'value': port['value'], 'unit': UNIT, 'cores': workers, 'kind': 'port',
                                'sample': port['sample'], 'arms': arms}
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--batch', type=int, default=100000, help='structure variants per GPU per step')
    ap.add_argument('--chunk', type=int, default=3125, help='HostPipeline chunk (structures)')
    ap.add_argument('--e2e-batch', type=int, default=200000, help='end-to-end arm: at most this many structures per GPU per step')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--cpu-sample', type=int, default=16384, help='CPU oracle: variants per worker')
    ap.add_argument('--ref-sample', type=int, default=256, help='real reference (arm A): variants in total')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--engine', default='jacobi', choices=['jacobi', 'two_ended'], help='modal engine of the fused path')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == '__main__':
    sys.exit(main())
