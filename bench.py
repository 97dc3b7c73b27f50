#!/usr/bin/env python
"""bench.py -- headline benchmark of the LUCI per-spaxel fitting hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--nx 2048 --ny 2064 --nz 841]

Workload (BASELINE.json configs[1]): full 2048 x 2064 synthetic SN3 cube, 5-line sincgauss fit with tied velocity /
broadening and the [NII] ratio tie, priors = truth + N(0, 15 km/s), truth * (1 + N(0, 0.15)).  One *step* = one
pass of the hot path over the whole cube (prep + LM fit + derived outputs for every spaxel) and, for N > 1, the
gather of the output maps to rank 0.  The cube is sharded in contiguous x-slabs over the N ranks (strong scaling:
the cube is fixed), no collective on the fit path.

Prints ONE JSON line (rank 0).  `value` = spaxels/s with the cube resident in HBM; `e2e` = the same metric with the
cube in pinned HOST memory, streamed through the C ABI in chunks (H2D of the spectra and D2H of the result rows inside
the timed region); `roofline` = algorithmic FLOP of the fit kernel / its CUDA-event time against the FP32-FMA
throughput measured by a probe kernel in the same run (MEASURED_PEAKS.json carries no FP32 figure); `cpu_baseline` =
the oracle port of the reference's row worker under joblib/loky on all host cores, on a bounded sample.

`--impl reference` times only that CPU arm (rank 0; other ranks exit 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

LINES = ['Halpha', 'NII6548', 'NII6583', 'SII6716', 'SII6731']
METRIC = 'spaxels/sec, SN3 5-line sincgauss fit, 2048x2064 synthetic cube'
F_MODEL = {'gaussian': 20, 'sinc': 30, 'sincgauss': 150}     # nominal flop per (sample, line): SURVEY.md 8(d)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--nx', type=int, default=2048)
    ap.add_argument('--ny', type=int, default=2064)
    ap.add_argument('--nz', type=int, default=841)
    ap.add_argument('--model', default='sincgauss')
    ap.add_argument('--uncertainty', type=int, default=0)
    ap.add_argument('--cpu-seconds', type=float, default=15.0, help='wall-clock budget of the CPU baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-reductions', action='store_true')
    ap.add_argument('--no-api', action='store_true')
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------- CPU arm
def _oracle_row(i, cube_rows, axis32, theta, hdr, n_steps, model, v0, b0, unc):
    """One joblib task = one y-row, like LuciBase.py:514-530 (module-level so loky can pickle it)."""
    import warnings
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from oracle import luci_oracle as O
    nx = cube_rows.shape[0]
    theta_map = np.full((nx, 1), theta)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        return O.fit_calc(0, 0, nx, 0, model, LINES, [1] * 5, [1] * 5, cube_rows, axis32, None, theta_map, hdr,
                          n_steps, 0, uncertainty_bool=unc, nii_cons=True, prior_maps=(v0[None, :], b0[None, :]))


def cpu_sample(nz, model, n_rows, row_len, seed=4242):
    """Synthetic spectra of the benchmark's distribution generated on the host (LuciSim restatement)."""
    from oracle import luci_oracle as O
    from luci_b200.sim import sim_axis
    rng = np.random.default_rng(seed)
    axis, n_steps, delta_x = sim_axis('SN3', n_steps=nz)
    rows = []
    for r in range(n_rows):
        spectra, v0, b0 = [], [], []
        for j in range(row_len):
            vel, broad = rng.uniform(-150, 150), rng.uniform(25, 60)
            snr = float(np.exp(rng.uniform(np.log(10), np.log(100))))
            a83 = 0.3 * rng.uniform(0.5, 1.5)
            amps = [1.0, a83 / 3, a83, 0.15 * rng.uniform(0.5, 1.5), 0.12 * rng.uniform(0.5, 1.5)]
            s = np.zeros(nz)
            corr = 1 / np.cos(np.deg2rad(11.96))
            w = corr * 1e7 / (2 * delta_x * n_steps)
            for k, line in enumerate(LINES):
                lam = 1e7 / O.LINE_DICT[line]
                pos = axis[np.argmin(np.abs(axis - ((vel / 3e5) * lam + lam)))]
                sig = broad * pos / (3e5 * corr)
                s += {'gaussian': lambda: O.gaussian_profile(axis, amps[k], pos, sig),
                      'sinc': lambda: O.sinc_profile(axis, amps[k], pos, w),
                      'sincgauss': lambda: O.sincgauss_profile(axis, amps[k], pos, sig, w)}[model]()
            s = s + 0.1 + (1.0 / snr) * rng.normal(0, 1, nz)
            spectra.append((s * 1e-17).astype(np.float32))
            v0.append(-vel * 299792 / 3e5 + rng.normal(0, 15))
            b0.append(abs(broad * np.cos(np.deg2rad(11.96)) * 299792 / 3e5 * (1 + rng.normal(0, 0.15))))
        rows.append((np.stack(spectra).astype(np.float64), np.array(v0), np.array(b0)))
    hdr = {'FILTER': 'SN3', 'STEP': delta_x}
    return rows, axis.astype(np.float32), hdr, n_steps


def run_cpu_arm(nz, model, unc, budget_s, steps=1, warmup=0):
    """spaxels/s of the oracle port of Luci.fit_calc under joblib (loky, all host cores), one task per row."""
    from joblib import Parallel, delayed
    cores = os.cpu_count() or 1
    row_len = 2
    n_rows = cores                                              # one row per core per step
    rows, axis32, hdr, n_steps = cpu_sample(nz, model, n_rows, row_len)
    par = Parallel(n_jobs=cores, backend='loky')
    # untimed: spin the worker pool up (imports SciPy in every worker)
    par(delayed(_oracle_row)(0, rows[0][0][:1], axis32, 11.96, hdr, n_steps, model, rows[0][1][:1], rows[0][2][:1], unc)
        for _ in range(cores))
    times = []
    total_steps = max(1, warmup) + max(1, steps) if warmup else max(1, steps)
    t_start = time.time()
    for s in range(total_steps):
        t0 = time.time()
        par(delayed(_oracle_row)(i, rows[i][0], axis32, 11.96, hdr, n_steps, model, rows[i][1], rows[i][2], unc)
            for i in range(n_rows))
        times.append(time.time() - t0)
        if time.time() - t_start > budget_s and s + 1 >= (warmup + 1 if warmup else 1):
            break
    timed = times[warmup:] if warmup and len(times) > warmup else times
    per_step = float(np.mean(timed))
    n_sp = n_rows * row_len
    return dict(value=n_sp / per_step, unit='spaxels/s', cores=cores, kind='port',
                sample='%d rows x %d spaxels per step (nz=%d, %s, uncertainty=%d), %d timed step(s); oracle/luci_oracle.fit_calc '
                       'under joblib loky, one task per row' % (n_rows, row_len, nz, model, int(unc), len(timed))), per_step, len(timed)


def reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cb, per_step, k = run_cpu_arm(args.nz, args.model, bool(args.uncertainty), budget_s=240.0, steps=args.steps,
                                  warmup=min(args.warmup, 1))
    line = {'impl': 'reference', 'metric': METRIC, 'value': cb['value'], 'unit': 'spaxels/s', 'n_gpus': args.gpus,
            'steps': k, 'warmup': min(args.warmup, 1), 'ms_per_step': per_step * 1e3, 'higher_is_better': True,
            'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': 'SN3 5-line %s, tied velocity/broadening + [NII] tie, nz=%d; bounded sample of the '
                                   '2048x2064 cube' % (args.model, args.nz)},
            'cpu_baseline': cb,
            'e2e': {'value': cb['value'], 'unit': 'spaxels/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    emit(line)


# ------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
              'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
              'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.FIELDS,
                                          '--format=csv,noheader,nounits', '-lms', '200'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(',')]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2])); power.append(float(parts[3]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(nm)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': float(max(mx)) if mx else None,
                'power_w_max': float(max(power)) if power else None, 'samples': len(sm), 'reasons': sorted(reasons)}


# ------------------------------------------------------------------------------------------------- GPU arm
_REAL_STDOUT = None


def quiet_stdout():
    """Libraries (NCCL's version banner, joblib workers) write to fd 1; the contract is ONE JSON line on stdout.  Send
    fd 1 to stderr for the duration of the run and keep the real stdout for emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), 'w')
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + '\n')
    out.flush()


def main():
    args = parse()
    quiet_stdout()
    if args.impl == 'reference':
        reference_arm(args)
        return
    import torch
    from luci_b200 import dist as ldist
    from luci_b200 import engine
    from luci_b200._capi import FitConfig
    from luci_b200.sim import SyntheticCube

    rank, world, local = ldist.init_process_group()
    cpus = ldist.bind_to_gpu_cpus(local) if world > 1 else None       # NUMA-local pinned buffers for the e2e stream
    assert torch.cuda.is_available(), 'bench.py needs a GPU (there is no CPU fallback)'
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    model, unc = args.model, bool(args.uncertainty)
    x_lo, x_hi = ldist.shard_bounds(args.nx, rank, world)
    nx_local = x_hi - x_lo
    sc = SyntheticCube(nx_local, args.ny, lines=LINES, model=model, n_steps=args.nz, x0=x_lo, nx_total=args.nx)
    cube3 = sc.synthesize(dev, seed=1000 + rank)
    n_local = nx_local * args.ny
    cube = cube3.view(n_local, args.nz)
    v0m, b0m = sc.priors()                                           # (ny, nx) maps
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    v0, b0 = t(v0m.T.reshape(-1).copy()), t(b0m.T.reshape(-1).copy())   # x-major like the cube rows
    theta = torch.full((n_local,), float(sc.theta), dtype=torch.float32, device=dev)
    cfg = FitConfig(model, LINES, [1] * 5, [1] * 5, sc.axis32, 'SN3', sc.delta_x, sc.n_steps, 0,
                    uncertainty_bool=unc, nii_cons=True)
    n_total = args.nx * args.ny
    K = cfg.row_floats

    counts = [(lambda b: (b[1] - b[0]) * args.ny)(ldist.shard_bounds(args.nx, r, world)) for r in range(world)]
    # N > 1: no gather step -- every rank's output kernel stores its rows straight into rank 0's [n_total, K] buffer
    # through peer-mapped memory (dist.PeerRows); NCCL point-to-point into the same buffer only if the mapping fails
    peer = ldist.PeerRows(n_total, K, dev, dst=0) if world > 1 else None
    use_peer = peer is not None and peer.ok
    row_lo = sum(counts[:rank])
    rows_win = peer.window(row_lo, row_lo + n_local)[0] if use_peer else None
    gathered = peer.rows if (peer is not None and rank == 0) else None
    status_loc = torch.empty((n_local,), dtype=torch.int32, device=dev)

    def step():
        if use_peer:
            res = engine.fit_batch(cfg, cube, theta, v0, b0, rows=rows_win, status=status_loc)
            return res, gathered
        res = engine.fit_batch(cfg, cube, theta, v0, b0, status=status_loc)
        rows = ldist.gather_rows(res['rows'], dst=0, counts=counts, out=gathered) if world > 1 else res['rows']
        return res, rows

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        res, rows = step()
    barrier()
    engine.launch_counter['n'] = 0
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    evs[0].record()
    for s in range(args.steps):
        res, rows = step()
        evs[s + 1].record()
    barrier()
    clocks = sampler.stop()
    total_ms = evs[0].elapsed_time(evs[-1])
    step_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    launches = engine.launch_counter['n']
    total_ms = ldist.max_over_ranks(total_ms, dev)
    ms_per_step = total_ms / args.steps
    value = n_total / (ms_per_step * 1e-3)

    # --- kernel time alone (no gather): equals the step at N = 1
    summ = engine.status_summary(res['status'])
    sum_evals = ldist.sum_over_ranks(summ['mean_evals'] * summ['n_fitted'], dev)
    n_fit_total = ldist.sum_over_ranks(summ['n_fitted'], dev)
    mean_E = sum_evals / max(n_fit_total, 1)
    nfit = cfg.c.fit_hi - cfg.c.fit_lo
    L, P_red = 5, 7                                                # 4 free amplitudes + v + beta + continuum
    # lf_lm_kernel only: per evaluation and sample L profiles with partials (F), J^T J / J^T r of the reduced system, and the two
    # FMAs per line of the exact (amplitude, velocity / broadening) Hessian term; per evaluation one Cholesky solve
    W = mean_E * nfit * (L * (F_MODEL[model] + 4) + P_red * (P_red + 3)) + mean_E * P_red ** 3 / 3.0
    # the dominant kernel (lf_lm_kernel) timed alone: the same pipeline stage by stage with CUDA events around every
    # launch on the launching stream (engine.fit_batch(time_stages=True) -> lucifit_fit_stages)
    barrier()
    engine.fit_batch(cfg, cube, theta, v0, b0, time_stages=True)
    stage_ms = engine.fit_batch(cfg, cube, theta, v0, b0, time_stages=True)['stage_ms']
    kern_ms = ldist.max_over_ranks(float(stage_ms['lm']), dev)
    n_slices = (n_local + engine.SLICE_SPAXELS - 1) // engine.SLICE_SPAXELS
    fp32_probe = engine.peak_probe(dev, 0)
    mufu_peak = engine.peak_probe(dev, 1)
    # FP32 roofline: MEASURED_PEAKS.json and the profiling guide carry no FP32 figure, so the denominator is the nominal
    # FMA rate at the SM clock sampled during the timed region (SMs x 128 lanes x 2 flop x f); the FFMA-chain probe of
    # this run is reported beside it (it reaches ~83 % of nominal)
    props = torch.cuda.get_device_properties(dev)
    sm_mhz = clocks.get('sm_mhz') or clocks.get('sm_max_mhz') or 1965.0
    fp32_peak = props.multi_processor_count * 128 * 2 * sm_mhz * 1e6 / 1e12
    achieved = (W * n_local) / (kern_ms * 1e-3) / 1e12
    alg_bytes = n_local * (4 * args.nz + 4 * K + 16)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    traffic, traffic_src = None, None
    try:                                                            # DRAM bytes of the kernel from the committed ncu capture
        tr = json.load(open(os.path.join(ROOT, 'profiles', 'lm_traffic.json')))
        traffic = float(tr['dram_bytes_per_spaxel']) * n_local
        traffic_src = tr['source']
    except Exception:
        pass
    roofline = {'bound': 'fp32', 'achieved': achieved, 'peak': fp32_peak, 'unit': 'TFLOP/s',
                'frac': achieved / fp32_peak if fp32_peak else None, 'traffic': traffic, 'traffic_source': traffic_src,
                'peak_source': 'nominal FP32 FMA rate = %d SMs x 128 lanes x 2 flop x %.0f MHz (SM clock sampled during the timed '
                               'region); MEASURED_PEAKS.json / B200_PROFILING.md state no FP32 figure' % (props.multi_processor_count, sm_mhz),
                'fp32_probe_tflops': fp32_probe, 'mufu_peak_tops': mufu_peak, 'mean_evals_E': mean_E, 'flop_per_spaxel_W': W,
                'kernel': 'lf_lm_kernel<%s,L=5,groups=1>' % model, 'kernel_ms': kern_ms,
                'launches': n_slices, 'stage_ms': {k: float(v) for k, v in stage_ms.items()},
                'kernel_share_of_step': kern_ms / sum(stage_ms.values()),
                'hbm_gbs_sanity': alg_bytes / (kern_ms * 1e-3) / 1e9, 'hbm_peak_gbs': hbm_peak,
                'algorithmic_bytes': alg_bytes}

    line = {'metric': METRIC, 'value': value, 'unit': 'spaxels/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'strong',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'full %dx%d synthetic SN3 cube, nz=%d (N_fit=%d), 5-line %s, tied velocity/broadening '
                                   '+ [NII] ratio tie%s; timed step = one lucifit_fit_batch pass (engine.fit_batch: prep + LM + '
                                   'output kernels) over the cube resident in HBM; the Luci.fit_entire_cube call itself is '
                                   'timed in api_e2e' % (args.nx, args.ny, args.nz, nfit, model, ' + uncertainties' if unc else ''),
                       'sharding': 'contiguous x-slabs, %d rank(s); rows of every rank land in rank 0\'s [n, 7L+5] buffer inside '
                                   'the step (%s)' % (world, 'stored by the output kernel through peer-mapped memory' if use_peer
                                                      else 'single GPU' if world == 1 else 'NCCL point-to-point gather'),
                       'l2': 'inputs (%.1f GB per rank) larger than L2' % (cube.numel() * 4 / 1e9),
                       'priors': 'truth + N(0,15 km/s), truth*(1+N(0,0.15))', 'snr': 'log-uniform 10..100'},
            'gpu_launches': launches, 'clocks': clocks, 'roofline': roofline,
            'fit_status': {k: summ[k] for k in ('n', 'n_converged', 'n_maxiter', 'n_bound', 'mean_evals', 'mean_iters')}}

    # --- the same fit started from priors measured on the device (lucifit_estimate_prior, the stand-in for the
    # reference's CNN) instead of the synthetic CNN-like priors above: prior kernel + fit, whole cube
    try:
        for _ in range(2):
            pv, pb = engine.estimate_prior(cfg, cube, theta)
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        barrier(); e0.record()
        pv, pb = engine.estimate_prior(cfg, cube, theta)
        e1.record()
        res_a = engine.fit_batch(cfg, cube, theta, pv, pb)
        e2.record(); e2.synchronize()
        prior_ms = ldist.max_over_ranks(e0.elapsed_time(e1), dev)
        total_ms_a = ldist.max_over_ranks(e0.elapsed_time(e2), dev)
        summ_a = engine.status_summary(res_a['status'])
        nfit_bytes = n_local * 4 * nfit
        line['auto_prior'] = {'prior_kernel_ms': prior_ms, 'prior_kernel_gbs': nfit_bytes / (prior_ms * 1e-3) / 1e9,
                              'prior_plus_fit_ms': total_ms_a, 'spaxels_per_s': n_total / (total_ms_a * 1e-3),
                              'mean_evals_E': ldist.sum_over_ranks(summ_a['mean_evals'] * summ_a['n_fitted'], dev) / max(n_fit_total, 1),
                              'n_converged': int(ldist.sum_over_ranks(summ_a['n_converged'], dev)),
                              'note': 'not the headline: BASELINE config 2 prescribes CNN-like priors (truth + noise)'}
        del pv, pb, res_a
    except Exception as exc:                                     # reported, never fatal for the headline number
        line['auto_prior'] = {'error': repr(exc)}

    # --- second row of SURVEY 8(d): the same cube with uncertainty_bool=True (adds the FD-stencil Hessian stage)
    if not args.uncertainty:
        try:
            cfg_u = FitConfig(model, LINES, [1] * 5, [1] * 5, sc.axis32, 'SN3', sc.delta_x, sc.n_steps, 0,
                              uncertainty_bool=True, nii_cons=True)
            engine.fit_batch(cfg_u, cube, theta, v0, b0)
            barrier()
            st_u = engine.fit_batch(cfg_u, cube, theta, v0, b0, time_stages=True)['stage_ms']
            tot_u = ldist.max_over_ranks(float(sum(st_u.values())), dev)
            line['with_uncertainties'] = {'ms_per_step': tot_u, 'spaxels_per_s': n_total / (tot_u * 1e-3),
                                          'stage_ms': {k: float(v) for k, v in st_u.items()},
                                          'note': 'uncertainty_bool=True: lf_unc_kernel added (LuciFit.py:713-722, LuciUtility.py:409-434); '
                                                  'sum of the stage kernels timed with CUDA events, no gather'}
        except Exception as exc:
            line['with_uncertainties'] = {'error': repr(exc)}

    # --- reductions (config 5): deep image + SNR map over the same cube, HBM-bound
    if not args.no_reductions:
        red = {}
        flo, fhi, nlo, nhi = 300, 360, 10, 60
        try:
            from luci_b200._capi import argmin_abs
            from luci_b200.constants import snr_windows
            (f0, f1), (n0, n1) = snr_windows('SN3')
            flo, fhi = argmin_abs(sc.axis32, f0), argmin_abs(sc.axis32, f1)
            nlo, nhi = argmin_abs(sc.axis32, n0), argmin_abs(sc.axis32, n1)
        except Exception:
            pass
        for name, fn in (('deep_image', lambda: engine.deep_image(cube)),
                         ('snr_map_method1', lambda: engine.snr_map(cube, 1, flo, fhi, nlo, nhi)),
                         ('snr1_plus_deep_fused', lambda: engine.snr_map(cube, 1, flo, fhi, nlo, nhi, with_deep=True)),
                         ('snr_map_method2', lambda: engine.snr_map(cube, 2, flo, fhi, nlo, nhi))):
            for _ in range(3 if name != 'snr_map_method2' else 1):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10 if name != 'snr_map_method2' else 2
            torch.cuda.synchronize(); e0.record()
            for _ in range(reps):
                fn()
            e1.record(); e1.synchronize()
            ms = e0.elapsed_time(e1) / reps
            gbs = cube.numel() * 4 / (ms * 1e-3) / 1e9
            red[name] = {'ms': ms, 'achieved': gbs, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': gbs / hbm_peak,
                         'bound': 'hbm'}
        # 2x2 spatial binning of the slab (bin_cube_function, LuciUtility.py:319-355): reads the cube, writes a quarter
        nxl = (nx_local // 2) * 2
        if nxl >= 2:
            c3 = cube.view(nx_local, args.ny, args.nz)
            fnb = lambda: engine.bin_cube(c3, 2, 0, nxl, 0, (args.ny // 2) * 2)
            for _ in range(2):
                fnb()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for _ in range(5):
                fnb()
            e1.record(); e1.synchronize()
            ms = e0.elapsed_time(e1) / 5
            gbs = 1.25 * nxl * ((args.ny // 2) * 2) * args.nz * 4 / (ms * 1e-3) / 1e9
            red['bin_cube_2x2'] = {'ms': ms, 'achieved': gbs, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': gbs / hbm_peak,
                                   'bound': 'hbm'}
        line['roofline_reductions'] = red

    # --- end to end: cube in pinned host memory, streamed in chunks through the C ABI
    if not args.no_e2e:
        line['e2e'] = run_e2e(torch, engine, ldist, cfg, cube, theta, v0, b0, dev, args, world, n_total, K)
    # --- the drop-in call: Luci.fit_entire_cube on the full cube (every rank holds it, like every joblib worker of the
    # reference sees cube_final), sharded over the ranks, FITS maps written by rank 0 -- wall clock around the call
    if not args.no_api:
        try:
            line['api_e2e'] = run_api(torch, ldist, args, dev, rank, world, model, unc)
        except Exception as exc:
            line['api_e2e'] = {'error': repr(exc)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:      # reported baseline: rank 0 at N = 1 only
        cb, _, _ = run_cpu_arm(args.nz, model, unc, budget_s=args.cpu_seconds)
        line['cpu_baseline'] = cb
    if world > 1:
        torch.distributed.barrier()
    if rank == 0:
        emit(line)
    if world > 1:
        torch.distributed.destroy_process_group()


def run_api(torch, ldist, args, dev, rank, world, model, unc):
    """Wall clock of ``Luci.fit_entire_cube`` (the call a LUCI user makes) on the full synthetic cube at N GPUs: rank r
    fits its y-rows, rows land in rank 0's maps through peer memory, rank 0 writes the FITS maps the reference writes
    (38 files for 5 lines).  The cube is resident on every GPU before the call (as ``cube_final`` is resident in host
    RAM before the reference's call); priors are passed as host maps like a user would."""
    import shutil
    import tempfile
    from luci_b200.luci import Luci
    from luci_b200.sim import SyntheticCube
    sc = SyntheticCube(args.nx, args.ny, lines=LINES, model=model, n_steps=args.nz)
    cube3 = sc.synthesize(dev, seed=1000)
    out = tempfile.mkdtemp(prefix='luci_bench_api_r%d_' % rank)
    try:
        luci = Luci.from_arrays(cube3, sc.hdr_dict, out, 'BENCH', spectrum_axis=sc.axis32, device=dev)
        luci.create_deep_image()                                       # (fit_cube writes it once if missing: not part of the fit)
        v0m, b0m = sc.priors()
        times = []
        for it in range(3):
            if world > 1:
                torch.distributed.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            luci.fit_entire_cube(LINES, model, [1] * 5, [1] * 5, prior_values=(v0m, b0m), uncertainty_bool=unc)
            torch.cuda.synchronize()
            if world > 1:
                torch.distributed.barrier()
            times.append(time.perf_counter() - t0)
        dt = ldist.max_over_ranks(min(times[1:]), dev)
        n = args.nx * args.ny
        summ = luci.last_fit_summary if rank == 0 else None
        n_files = sum(len(f) for _, _, f in os.walk(out)) if rank == 0 else 0
        return {'value': n / dt, 'unit': 'spaxels/s', 'seconds_per_call': dt, 'calls_timed': 2, 'n_gpus': world,
                'fits_files_written': n_files, 'n_converged': summ['n_converged'] if summ else None,
                'note': 'wall clock around Luci.fit_entire_cube(lines, %r, vel_rel, sigma_rel, prior_values=host maps): sharding, '
                        'fit, rows to rank 0, host maps, FITS output included; best of 2 calls after one warm-up' % model}
    finally:
        shutil.rmtree(out, ignore_errors=True)


def run_e2e(torch, engine, ldist, cfg, cube, theta, v0, b0, dev, args, world, n_total, K):
    """Host-resident cube through the public streaming entry (engine.HostCubeStream -> lucifit_fit_batch): every step
    copies the spectra H2D from pinned memory (chunked, overlapped with the fit of the previous chunk) and the result
    rows D2H."""
    n_local, nz = cube.shape
    avail_kb = 0
    try:
        for ln in open('/proc/meminfo'):
            if ln.startswith('MemAvailable'):
                avail_kb = int(ln.split()[1])
    except Exception:
        pass
    need = n_local * nz * 4 * world
    frac = 1.0
    if avail_kb and need * 2.5 > avail_kb * 1024:
        frac = max(0.05, (avail_kb * 1024 / 2.5) / need)
    n_e2e = int(n_local * frac)
    host_cube = torch.empty((n_e2e, nz), dtype=torch.float32, pin_memory=True)
    host_cube.copy_(cube[:n_e2e])
    host_rows = torch.empty((n_e2e, K), dtype=torch.float32, pin_memory=True)
    host_par = torch.empty((3, n_e2e), dtype=torch.float32, pin_memory=True)     # theta, prior velocity, prior broadening
    host_par[0].copy_(theta[:n_e2e]); host_par[1].copy_(v0[:n_e2e]); host_par[2].copy_(b0[:n_e2e])
    torch.cuda.synchronize()
    # the PCIe ceiling of this box, for context: one big pinned H2D copy
    probe = torch.empty((min(n_e2e, 1 << 18), nz), dtype=torch.float32, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    probe.copy_(host_cube[:probe.shape[0]], non_blocking=True)
    torch.cuda.synchronize(); e0.record()
    probe.copy_(host_cube[:probe.shape[0]], non_blocking=True)
    e1.record(); e1.synchronize()
    h2d_gbs = probe.numel() * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9
    # ... and with ALL ranks copying at the same time (what the streamed fit asks of the host at N > 1)
    h2d_conc = h2d_gbs
    if world > 1:
        torch.distributed.barrier()
        torch.cuda.synchronize(); e0.record()
        for _ in range(4):
            probe.copy_(host_cube[:probe.shape[0]], non_blocking=True)
        e1.record(); e1.synchronize()
        mine = 4 * probe.numel() * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9
        h2d_conc = ldist.sum_over_ranks(mine, dev) / world
    del probe
    # chunk: 131 072 spaxels (441 MB) when the rank streams the whole cube; smaller shards keep >= 16 chunks in the pipeline
    # so that its fill (first copy) and drain (last fit + copy back) stay a small part of the step
    chunk = int(min(1 << 17, max(1 << 14, (n_e2e // 16 + 1023) // 1024 * 1024)))
    stream = engine.HostCubeStream(cfg, dev, chunk=chunk)
    stream.run(host_cube, host_par, host_rows)                       # warm-up
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    reps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(reps):
        stream.run(host_cube, host_par, host_rows)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    dt = ldist.max_over_ranks(dt, dev)
    n_done = ldist.sum_over_ranks(n_e2e, dev)
    h2d = int(n_done * (nz + 3) * 4)
    return {'value': n_done / dt, 'unit': 'spaxels/s', 'h2d_bytes_per_step': h2d,
            'd2h_bytes_per_step': int(n_done * K * 4), 'ms_per_step': dt * 1e3,
            'h2d_gbs_achieved_per_gpu': h2d / world / dt / 1e9, 'h2d_gbs_single_copy': h2d_gbs,
            'h2d_gbs_concurrent_copy_per_gpu': h2d_conc,
            'note': 'cube in pinned host memory (process bound to the GPU-local CPU cores when N > 1), engine.HostCubeStream: %d-spaxel chunks, %d device buffers, copy / fit / '
                    'copy-back on three streams; fraction of the cube streamed per step: %.2f'
                    % (stream.chunk, stream.nb, frac)}


if __name__ == '__main__':
    main()
