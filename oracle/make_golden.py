"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED reference
(``/root/reference``: ``LuciBase.Luci.fit_calc`` -> ``LUCI.LuciFit.Fit``) under ``oracle/ref_shim.py`` on seeded
synthetic spectra built with the reference's own ``LuciSim.Spectrum``.  Run in the build container only:

    python -m oracle.make_golden            # all cases
    python -m oracle.make_golden fit_sincgauss_sn3

Each fixture stores the exact float32 inputs (cube rows, axis, priors, theta) and the reference's outputs (the 12
quantities of fit_calc plus fit_sol / uncertainties / scale / SLSQP status) so that tests on the GPU box can
compare without the reference tree.
"""
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')

SN3 = ['Halpha', 'NII6548', 'NII6583', 'SII6716', 'SII6731']

CASES = {
    # name: (filter, resolution, lines, model, vel_rel, sigma_rel, n, uncertainty, broad range, extra)
    'fit_sincgauss_sn3': dict(filter='SN3', R=5000, lines=SN3, model='sincgauss', vel_rel=[1] * 5, sigma_rel=[1] * 5,
                              n=24, unc=True, broad=(25, 60), seed=101),
    'fit_gaussian_sn3': dict(filter='SN3', R=5000, lines=SN3, model='gaussian', vel_rel=[1] * 5, sigma_rel=[1] * 5,
                             n=24, unc=True, broad=(60, 100), seed=102),
    'fit_sinc_sn3': dict(filter='SN3', R=5000, lines=SN3, model='sinc', vel_rel=[1] * 5, sigma_rel=[1] * 5,
                         n=6, unc=True, broad=(20, 60), seed=103),
    'fit_gaussian_sn2': dict(filter='SN2', R=5000, lines=['Hbeta', 'OIII4959', 'OIII5007'], model='gaussian',
                             vel_rel=[1] * 3, sigma_rel=[1] * 3, n=16, unc=False, broad=(60, 100), seed=104),
    'fit_gaussian_sn1': dict(filter='SN1', R=5000, lines=['OII3726', 'OII3729'], model='gaussian', vel_rel=[1, 1],
                             sigma_rel=[1, 1], n=16, unc=False, broad=(60, 100), seed=105),
    'fit_sincgauss_groups': dict(filter='SN3', R=5000, lines=SN3, model='sincgauss', vel_rel=[1, 2, 2, 3, 3],
                                 sigma_rel=[1, 2, 2, 3, 3], n=12, unc=True, broad=(30, 60), seed=106),
    # the same three independent tie groups at SNR 40-100, where every group is well determined (no zero-width spikes)
    'fit_sincgauss_groups_hi': dict(filter='SN3', R=5000, lines=SN3, model='sincgauss', vel_rel=[1, 2, 2, 3, 3],
                                    sigma_rel=[1, 2, 2, 3, 3], n=24, unc=True, broad=(30, 60), seed=114, snr=(40, 100)),
    'fit_sincgauss_double': dict(filter='SN3', R=5000, lines=SN3 + ['Halpha'], model='sincgauss',
                                 vel_rel=[1, 1, 1, 1, 1, 2], sigma_rel=[1, 1, 1, 1, 1, 2], n=10, unc=True,
                                 broad=(30, 50), seed=107, double=True),
    # BASELINE config 4 with the two components started apart: one (velocity, broadening) prior per tie group.  The
    # reference's CNN can only give one pair per spectrum, so for this fixture ``Fit.line_vals_estimate``
    # (LuciFit.py:382-430) is wrapped to set ``vel_ml`` / ``broad_ml`` to the prior of the line's group before the
    # unmodified body runs; everything downstream (SLSQP, ties, hessianComp uncertainties) is the reference as is.
    'fit_sincgauss_double_pg': dict(filter='SN3', R=5000, lines=SN3 + ['Halpha'], model='sincgauss',
                                    vel_rel=[1, 1, 1, 1, 1, 2], sigma_rel=[1, 1, 1, 1, 1, 2], n=24, unc=True,
                                    broad=(30, 50), seed=113, double=True, per_group=True),
    'fit_sincgauss_single': dict(filter='SN3', R=5000, lines=['Halpha'], model='sincgauss', vel_rel=[1],
                                 sigma_rel=[1], n=12, unc=True, broad=(25, 60), seed=108),
    'fit_sinc_noml': dict(filter='SN3', R=5000, lines=SN3, model='sinc', vel_rel=[1] * 5, sigma_rel=[1] * 5,
                          n=4, unc=False, broad=(20, 60), seed=109, noml=True),
    # transmission curve (LuciFit.py:124-125, 197-207): divides the spectrum where T > 0.5 AFTER the normalisation, so
    # it moves the scale and the chi2 but not the optimum
    'fit_sincgauss_trans': dict(filter='SN3', R=5000, lines=SN3, model='sincgauss', vel_rel=[1] * 5, sigma_rel=[1] * 5,
                                n=8, unc=True, broad=(25, 60), seed=112, trans=True),
    # the reference's freeze branch (initial_values = frozen velocity / broadening, amplitudes + continuum fitted
    # without constraints, LuciFit.py:688-709); only sincgauss works there (gaussian / sinc raise IndexError)
    'fit_sincgauss_freeze': dict(filter='SN3', R=5000, lines=SN3, model='sincgauss', vel_rel=[1] * 5, sigma_rel=[1] * 5,
                                 n=16, unc=False, broad=(25, 60), seed=111, freeze=True),
    # the benchmark's configuration (BASELINE.json configs[1]: nz = 841, 5-line sincgauss, ties + [NII] ratio) on a
    # sample large enough to test the north_star's "99.9 % of spaxels" (run with --jobs 8: ~0.8 s/spaxel/core)
    'parity_sincgauss_cfg2': dict(filter='SN3', R=5917, lines=SN3, model='sincgauss', vel_rel=[1] * 5,
                                  sigma_rel=[1] * 5, n=1000, unc=False, broad=(25, 60), seed=110),
    # uncertainties (hessianComp, LuciUtility.py:409-434) at the benchmark's spectral length nz = 841
    'unc_sincgauss_cfg2': dict(filter='SN3', R=5917, lines=SN3, model='sincgauss', vel_rel=[1] * 5,
                               sigma_rel=[1] * 5, n=32, unc=True, broad=(25, 60), seed=115),
    # NaN channels: the row worker drops them and fits the rest (LuciBase.py:358-360); samples are knocked out inside
    # the fit, noise and continuum windows, next to a line core and as a block
    'fit_sincgauss_nan': dict(filter='SN3', R=5000, lines=SN3, model='sincgauss', vel_rel=[1] * 5, sigma_rel=[1] * 5,
                              n=16, unc=True, broad=(25, 60), seed=117, nan=True, snr=(20, 100)),
    # HELD-OUT sample of the same configuration (different seed): thresholds of the device code are never tuned on it
    'parity_sincgauss_heldout': dict(filter='SN3', R=5917, lines=SN3, model='sincgauss', vel_rel=[1] * 5,
                                     sigma_rel=[1] * 5, n=400, unc=False, broad=(25, 60), seed=2718),
}

BASE_AMP = {'Halpha': 1.0, 'NII6583': 0.3, 'NII6548': 0.1, 'SII6716': 0.15, 'SII6731': 0.12, 'Hbeta': 1.0,
            'OIII5007': 0.9, 'OIII4959': 0.3, 'OII3726': 1.0, 'OII3729': 0.8}


def build_inputs(case):
    """Seeded spectra from the reference's LuciSim (global numpy RNG, like the reference)."""
    from oracle import ref_shim
    ref_shim.install()
    import LUCI.LuciSim as LS
    rng = np.random.default_rng(case['seed'])
    n = case['n']
    lines = case['lines']
    rows, v0, b0 = [], [], []
    theta = 11.96
    for i in range(n):
        vel = float(rng.uniform(-40, 40) if case.get('noml') else rng.uniform(-150, 150))
        broad = float(rng.uniform(*case['broad']))
        snr_lo, snr_hi = case.get('snr', (10, 100))
        snr = float(np.exp(rng.uniform(np.log(snr_lo), np.log(snr_hi))))
        amps = [BASE_AMP[l] if BASE_AMP[l] == 1.0 else BASE_AMP[l] * rng.uniform(0.5, 1.5) for l in lines]
        if 'NII6548' in lines and 'NII6583' in lines:
            amps[lines.index('NII6548')] = amps[lines.index('NII6583')] / 3
        vels = [vel] * len(lines)
        broads = [broad] * len(lines)
        vfit = [-vel * 299792 / 3e5]
        if case.get('double'):
            amps[-1] = 0.5
            vels[-1] = vel - 200.0 * 3e5 / 299792     # second component at +200 km/s in the fitted convention
            vfit.append(-vels[-1] * 299792 / 3e5)
        np.random.seed(case['seed'] * 1000 + i)
        sp = LS.Spectrum(lines, case['model'], amps, 0.0, 0.0, case['filter'], case['R'], snr)
        sp.velocity, sp.broadening = vels, broads
        axis, spec = sp.create_spectrum()
        rows.append(((spec + 0.1) * 1e-17).astype(np.float32))
        bfit = broad * np.cos(np.deg2rad(theta)) * 299792 / 3e5
        # per velocity/sigma group priors are not expressible through the reference's CNN (one (v, b) pair per
        # spectrum), so the prior is that of the first component
        if case.get('per_group'):                    # one prior per tie group (order of first appearance)
            v0.append([vfit[0] + rng.normal(0, 10), vfit[1] + rng.normal(0, 10)])
            b0.append([abs(bfit * (1 + rng.normal(0, 0.1))), abs(bfit * (1 + rng.normal(0, 0.1)))])
        elif case.get('freeze'):                     # frozen values: a previous fit's result, close to the truth
            v0.append(vfit[0] + rng.normal(0, 3))
            b0.append(abs(bfit * (1 + rng.normal(0, 0.05))))
        else:
            v0.append(vfit[0] + rng.normal(0, 15))
            b0.append(abs(bfit * (1 + rng.normal(0, 0.15))))
    cube = np.stack(rows)
    extra = {}
    if case.get('nan'):
        ax = np.asarray(axis, dtype=np.float64)
        inwin = lambda lo, hi: np.flatnonzero((ax > lo) & (ax < hi))
        cores = [1e7 / 656.28, 1e7 / 658.341, 1e7 / 654.803, 1e7 / 671.647, 1e7 / 673.085]
        far_from_cores = np.array([np.min(np.abs(ax[j] - np.array(cores))) > 12.0 for j in range(len(ax))])
        fitw = np.array([j for j in inwin(14760, 15390) if far_from_cores[j]])
        for i in range(n):
            idx = list(rng.choice(fitw, 3, replace=False)) + list(rng.choice(inwin(15602, 15800), 2, replace=False)) \
                + list(rng.choice(inwin(14855, 14945), 1)) + list(rng.choice(inwin(14000, 14500), 2, replace=False))
            if i % 4 == 1:                            # a block of four consecutive samples in the fit window
                j0 = int(rng.choice(fitw[:-8]))
                idx += [j0, j0 + 1, j0 + 2, j0 + 3]
            if i % 4 == 2:                            # next to the Halpha core (two samples off the expected centre)
                pos = 1e7 / (656.28 * (1 + float(v0[i]) / 299792.0))
                idx += [int(np.argmin(np.abs(ax - pos))) + 3]
            cube[i, np.array(idx, dtype=int)] = np.nan
    if case.get('trans'):                             # smooth band-pass: 0.3 at the edges (below the 0.5 cut), 0.95 inside
        z = np.linspace(-1, 1, cube.shape[1])
        extra['trans'] = 0.3 + 0.65 * np.exp(-(z / 0.8) ** 8)
    return dict(**extra, cube=cube, axis32=np.asarray(axis, dtype=np.float32), v0=np.array(v0, np.float32),
                b0=np.array(b0, np.float32), theta=np.full(n, theta, np.float32), n_steps=sp.n_steps,
                delta_x=sp.delta_x)


def _run_slice(args):
    case, inp, lo, hi = args
    sub = dict(inp)
    for k in ('cube', 'v0', 'b0', 'theta'):
        sub[k] = inp[k][lo:hi]
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        return run_reference(case, sub)


def run_reference_parallel(case, inp, jobs):
    """The same per-spaxel reference runs, fanned out over processes (each spaxel is independent)."""
    import multiprocessing as mp
    n = inp['cube'].shape[0]
    edges = np.linspace(0, n, jobs + 1).astype(int)
    with mp.get_context('fork').Pool(jobs) as pool:
        parts = pool.map(_run_slice, [(case, inp, int(a), int(b)) for a, b in zip(edges[:-1], edges[1:]) if b > a])
    out = {}
    for k in parts[0]:
        if k == 'ref_seconds_per_spaxel':
            out[k] = np.array(np.mean([float(p[k]) for p in parts]))
        else:
            out[k] = np.concatenate([p[k] for p in parts], axis=0)
    return out


def run_reference(case, inp):
    from oracle import ref_shim
    Luci, LF = ref_shim.modules()
    n, nz = inp['cube'].shape
    lines = case['lines']
    captured = []
    orig_fit = LF.Fit.fit

    def spy(self, *a, **k):
        d = orig_fit(self, *a, **k)
        captured.append((np.array(self.fit_sol, dtype=float), np.array(self.uncertainties, dtype=float),
                         float(self.spectrum_scale), float(self.noise)))
        return d

    LF.Fit.fit = spy
    noml = bool(case.get('noml'))
    per_group = bool(case.get('per_group'))
    orig_lve = LF.Fit.line_vals_estimate
    if per_group:
        dense = lambda rel: [list(dict.fromkeys(rel)).index(r) for r in rel]
        gv, gs = dense(case['vel_rel']), dense(case['sigma_rel'])

        def lve(self, line_name):
            k = getattr(self, '_lve_k', 0)
            self.vel_ml = float(inp['v0'][_cur['i']][gv[k]])
            self.broad_ml = float(inp['b0'][_cur['i']][gs[k]])
            self._lve_k = (k + 1) % len(lines)
            return orig_lve(self, line_name)

        LF.Fit.line_vals_estimate = lve
    first = (lambda a: float(a[0])) if per_group else float
    ref_shim.install_prior(None if noml else (lambda fit: (first(inp['v0'][_cur['i']]), first(inp['b0'][_cur['i']]))))
    out = {k: [] for k in ('amplitudes', 'fluxes', 'flux_errors', 'velocities', 'vels_errors', 'sigmas',
                           'sigmas_errors', 'chi2', 'corr', 'axis_step', 'continuum', 'continuum_error')}
    wsyn = np.linspace(14700, 15600, 460).astype(np.float32)     # only feeds the (stubbed) CNN input
    t0 = time.time()
    try:
        for i in range(n):
            _cur['i'] = i
            cube_slice = inp['cube'][i:i + 1].astype(np.float64)        # [x=1, z]
            theta_map = np.full((1, 1), float(inp['theta'][i]))
            res = Luci.fit_calc(0, 0, 1, 0, case['model'], lines, case['vel_rel'], case['sigma_rel'],
                                cube_slice=cube_slice, spectrum_axis=inp['axis32'], wavenumbers_syn=wsyn,
                                transmission_interpolated=inp.get('trans'), interferometer_theta=theta_map,
                                hdr_dict={'FILTER': case['filter'], 'STEP': inp['delta_x']}, step_nb=inp['n_steps'],
                                zpd_index=0, mdn=False, ML_bool=not noml, uncertainty_bool=case['unc'],
                                nii_cons=True, resolution=case['R'], Luci_path='/root/reference/',
                                **({'initial_values': [np.array([[float(inp['v0'][i])]]), np.array([[float(inp['b0'][i])]])]}
                                   if case.get('freeze') else {}))
            for key, lst in zip(out.keys(), res[1:]):
                out[key].append(lst[0])
    finally:
        LF.Fit.fit = orig_fit
        LF.Fit.line_vals_estimate = orig_lve
    dt = time.time() - t0
    res = {('ref_' + k): np.array(v, dtype=np.float64) for k, v in out.items()}
    res['ref_fit_sol'] = np.stack([c[0] for c in captured])
    res['ref_fit_unc'] = np.stack([c[1] for c in captured])
    res['ref_scale'] = np.array([c[2] for c in captured])
    res['ref_noise'] = np.array([c[3] for c in captured])
    res['ref_seconds_per_spaxel'] = np.array(dt / n)
    return res


_cur = {'i': 0}


def make(name, jobs=1):
    case = CASES[name]
    inp = build_inputs(case)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        ref = run_reference_parallel(case, inp, jobs) if jobs > 1 else run_reference(case, inp)
    meta = dict(filter=case['filter'], model=case['model'], lines=np.array(case['lines']),
                vel_rel=np.array(case['vel_rel']), sigma_rel=np.array(case['sigma_rel']),
                uncertainty=np.array(case['unc']), noml=np.array(bool(case.get('noml'))),
                freeze=np.array(bool(case.get('freeze'))), per_group=np.array(bool(case.get('per_group'))),
                n_steps=np.array(inp['n_steps']), delta_x=np.array(inp['delta_x']))
    os.makedirs(GOLDEN, exist_ok=True)
    path = os.path.join(GOLDEN, name + '.npz')
    if 'trans' in inp:
        meta['trans'] = inp['trans']
    np.savez_compressed(path, cube=inp['cube'], axis32=inp['axis32'], v0=inp['v0'], b0=inp['b0'], theta=inp['theta'],
                        **meta, **ref)
    print('%s: n=%d  %.2f s/spaxel (reference) -> %s' % (name, case['n'], float(ref['ref_seconds_per_spaxel']), path))


if __name__ == '__main__':
    argv = sys.argv[1:]
    jobs = 1
    if '--jobs' in argv:
        jobs = int(argv[argv.index('--jobs') + 1])
        del argv[argv.index('--jobs'):argv.index('--jobs') + 2]
    sys.argv = sys.argv[:1] + argv
    names = [a for a in argv if a not in ("reductions", "region_strip")] or ([] if ("reductions" in argv or "region_strip" in argv) else
                                                       [c for c in CASES if not c.startswith('parity_')])
    for nm in names:
        make(nm, jobs)


# ------------------------------------------------------------------------------------------------ reductions
class _FakeLuci:
    """Minimal stand-in for a ``Luci`` instance so that the reference's own ``create_deep_image`` /
    ``create_snr_map`` / ``bin_cube`` method bodies (LuciBase.py:147-215, 1048-1191, 840-842) run on an in-memory
    cube (the real constructor needs an ORBS HDF5 file)."""

    def __init__(self, cube, axis32, out_dir):
        self.cube_final = cube
        self.cube_binned = None
        self.header_binned = None
        self.cube_path = 'none'
        self.output_dir = out_dir
        self.object_name = 'golden'
        self.dimz = 0
        self.hdr_dict = {'FILTER': 'SN3'}
        self.spectrum_axis = axis32
        self.header = {'CRPIX1': 10.0, 'CRPIX2': 12.0, 'PC1_1': 1.0, 'PC1_2': 0.0, 'PC2_1': 0.0, 'PC2_2': 1.0}
        self.deep_image = None


def make_reductions():
    import tempfile
    from oracle import ref_shim, luci_oracle as O
    Luci, LF = ref_shim.modules()
    rng = np.random.default_rng(2024)
    nx, ny = 23, 14
    axis, nsteps, dx, w = O.sim_axis('SN3', 5000)
    axis32 = axis.astype(np.float32)
    cube = np.zeros((nx, ny, len(axis)), np.float32)
    for x in range(nx):
        for y in range(ny):
            am = [1.0, 0.1, 0.3, 0.15, 0.12]
            snr = float(np.exp(rng.uniform(np.log(3), np.log(100))))
            _, s = O.sim_spectrum(SN3, 'sincgauss', am, rng.uniform(-150, 150), rng.uniform(25, 60), 'SN3', 5000, snr, rng=rng)
            cube[x, y] = ((s + 0.1) * 1e-17).astype(np.float32)
    cube[3, 4, 100:104] = np.nan          # NaNs exercise nansum / nanstd
    cube[7, 2, :] = np.nan
    cube64 = cube.astype(np.float64)
    tmp = tempfile.mkdtemp()
    fake = _FakeLuci(cube64, axis32, tmp)
    fake.bin_cube = lambda *a: Luci.bin_cube(fake, *a)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        Luci.create_deep_image(fake)
        deep = np.array(fake.deep_image)
        snr = {}
        for method in (1, 2):
            masks = Luci.create_snr_map(fake, 0, nx, 0, ny, method=method, n_threads=1)
            snr[method] = np.array(masks[0].data)
        from LUCI.LuciUtility import bin_cube_function, bin_mask
        _, binned = bin_cube_function(cube64, dict(fake.header), 2, 1, 21, 2, 12)
        mask = rng.uniform(size=(nx, ny)) > 0.6
        bmask = bin_mask(mask, 2, 0, nx, 0, ny)
    path = os.path.join(GOLDEN, 'reductions_sn3.npz')
    np.savez_compressed(path, cube=cube, axis32=axis32, ref_deep=deep, ref_snr1=snr[1], ref_snr2=snr[2],
                        ref_binned=binned.astype(np.float32), bin_args=np.array([2, 1, 21, 2, 12]), mask=mask,
                        ref_bin_mask=bmask)
    print('reductions_sn3 ->', path, deep.shape, snr[1].shape)


if __name__ == '__main__' and (len(sys.argv) == 1 or 'reductions' in sys.argv[1:]):
    make_reductions()


# ------------------------------------------------------------------------------------------ region strip (config 1)
def _strip_row(args):
    """One y-row of the strip through the reference's own row worker (LuciBase.py:258-406) with the mask."""
    sl, cube, axis32, mask, theta_map, v0map, b0map, n_steps, delta_x = args
    from oracle import ref_shim
    Luci, LF = ref_shim.modules()
    captured = []
    orig_fit = LF.Fit.fit

    def spy(self, *a, **k):
        d = orig_fit(self, *a, **k)
        captured.append((np.array(self.fit_sol, dtype=float), float(self.spectrum_scale)))
        return d

    LF.Fit.fit = spy
    nx = cube.shape[0]
    key = lambda spec: np.asarray(spec[:6], dtype=np.float64).tobytes()
    lut = {key(cube[x, sl].astype(np.float64)): (float(v0map[sl, x]), float(b0map[sl, x])) for x in range(nx)}
    ref_shim.install_prior(lambda fit: lut[key(fit.spectrum)])
    wsyn = np.linspace(14700, 15600, 460).astype(np.float32)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            res = Luci.fit_calc(sl, 0, nx, 0, 'sinc', SN3, [1] * 5, [1] * 5, cube_slice=cube[:, sl, :].astype(np.float64),
                                spectrum_axis=axis32, wavenumbers_syn=wsyn, transmission_interpolated=None,
                                interferometer_theta=theta_map, hdr_dict={'FILTER': 'SN3', 'STEP': delta_x},
                                step_nb=n_steps, zpd_index=0, mdn=False, mask=mask, ML_bool=True, uncertainty_bool=False,
                                nii_cons=True, resolution=5000, Luci_path='/root/reference/')
    finally:
        LF.Fit.fit = orig_fit
    return res, captured


def make_region_strip(jobs=8):
    """BASELINE config 1 as a strip: 100 x 8 spaxels, SN3 five-line sinc model, ``fit_region`` with a mask.  Every y-row goes
    through the reference's row worker exactly as ``fit_region`` hands it out (LuciBase.py:691-705) and the maps are
    assembled like :706-718.  Stored both as maps (ny, nx[, L]) and, for the masked-in spaxels in (y, x) order, in the flat
    layout of the other fixtures."""
    import multiprocessing as mp
    from oracle import ref_shim
    ref_shim.install()
    import LUCI.LuciSim as LS
    rng = np.random.default_rng(116)
    nx, ny, theta = 100, 8, 11.96
    cube, v0map, b0map = None, np.zeros((ny, nx), np.float32), np.zeros((ny, nx), np.float32)
    for x in range(nx):
        for y in range(ny):
            vel, broad = float(rng.uniform(-150, 150)), float(rng.uniform(20, 60))
            snr = float(np.exp(rng.uniform(np.log(10), np.log(100))))
            amps = [1.0, 0.1, 0.3 * rng.uniform(0.5, 1.5), 0.15 * rng.uniform(0.5, 1.5), 0.12 * rng.uniform(0.5, 1.5)]
            amps[1] = amps[2] / 3
            np.random.seed(116000 + x * ny + y)
            sp = LS.Spectrum(SN3, 'sinc', amps, 0.0, 0.0, 'SN3', 5000, snr)
            sp.velocity, sp.broadening = [vel] * 5, [broad] * 5
            axis, spec = sp.create_spectrum()
            if cube is None:
                cube = np.zeros((nx, ny, len(axis)), np.float32)
            cube[x, y] = ((spec + 0.1) * 1e-17).astype(np.float32)
            v0map[y, x] = -vel * 299792 / 3e5 + rng.normal(0, 15)
            b0map[y, x] = abs(broad * np.cos(np.deg2rad(theta)) * 299792 / 3e5 * (1 + rng.normal(0, 0.15)))
    mask = rng.uniform(size=(nx, ny)) < 0.7
    axis32 = np.asarray(axis, dtype=np.float32)
    theta_map = np.full((nx, ny), theta)
    args = [(sl, cube, axis32, mask, theta_map, v0map, b0map, sp.n_steps, sp.delta_x) for sl in range(ny)]
    with mp.get_context('fork').Pool(min(jobs, ny)) as pool:
        out = pool.map(_strip_row, args)
    L = 5
    names = ('amplitudes', 'fluxes', 'flux_errors', 'velocities', 'vels_errors', 'sigmas', 'sigmas_errors', 'chi2', 'corr',
             'axis_step', 'continuum', 'continuum_error')
    maps = {k: (np.zeros((ny, nx, L), np.float32) if i < 7 else np.zeros((ny, nx), np.float32)) for i, k in enumerate(names)}
    flat = {'cube': [], 'v0': [], 'b0': [], 'ref_fit_sol': [], 'ref_scale': [], 'yx': []}
    flat.update({'ref_' + k: [] for k in names})
    for res, captured in out:
        i = res[0]
        for k, lst in zip(names, res[1:]):
            maps[k][i] = lst                                    # LuciBase.py:706-718
        cap = iter(captured)
        for x in range(nx):
            if mask[x, i]:
                sol, scale = next(cap)
                flat['cube'].append(cube[x, i]); flat['v0'].append(v0map[i, x]); flat['b0'].append(b0map[i, x])
                flat['ref_fit_sol'].append(sol); flat['ref_scale'].append(scale); flat['yx'].append((i, x))
                for k, lst in zip(names, res[1:]):
                    flat['ref_' + k].append(np.asarray(lst[x], dtype=np.float64))
    n = len(flat['yx'])
    path = os.path.join(GOLDEN, 'region_sinc_strip.npz')
    np.savez_compressed(path, cube3d=cube, mask=mask, v0map=v0map, b0map=b0map, axis32=axis32,
                        **{'map_' + k: v for k, v in maps.items()},
                        cube=np.stack(flat['cube']), v0=np.array(flat['v0'], np.float32), b0=np.array(flat['b0'], np.float32),
                        theta=np.full(n, theta, np.float32), yx=np.array(flat['yx']), ref_fit_sol=np.stack(flat['ref_fit_sol']),
                        ref_scale=np.array(flat['ref_scale']), **{'ref_' + k: np.stack(flat['ref_' + k]) for k in names},
                        filter='SN3', model='sinc', lines=np.array(SN3), vel_rel=np.array([1] * 5), sigma_rel=np.array([1] * 5),
                        uncertainty=np.array(False), noml=np.array(False), freeze=np.array(False), per_group=np.array(False),
                        n_steps=np.array(sp.n_steps), delta_x=np.array(sp.delta_x))
    print('region_sinc_strip: %d fitted of %d spaxels ->' % (n, nx * ny), path)


if __name__ == '__main__' and 'region_strip' in sys.argv[1:]:
    make_region_strip()
