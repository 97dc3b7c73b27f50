"""TEST INFRASTRUCTURE ONLY -- golden vectors of the reference's sky-line fit (``Fit.fit(sky_line=True)``,
LUCI/LuciFit.py:837-928, the kernel of ``Luci.skyline_calibration``, LuciBase.py:1205-1284) from the UNMODIFIED
reference under ``oracle/ref_shim.py``.  Run in the build container only:

    python -m oracle.make_golden_skyline

Spectra: the OH 6498.729 A line of ``Data/sky_lines.dat`` as a sinc of the instrument's width at a velocity offset of
30-130 km/s on an SN3 R = 5000 axis (the reference's own ``LuciSim`` axis), a flat continuum and Gaussian noise.
Stored: float32 spectra, axis, theta, and the reference's (velocity, velocity_error, fit_sol, uncertainties)."""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
OH_NM = 649.8729            # Data/sky_lines.dat: 6498.729 Angstrom


def main():
    from oracle import ref_shim
    ref_shim.install()
    import LUCI.LuciSim as LS
    from LUCI.LuciFit import Fit
    from LUCI.LuciFunctions import Sinc
    rng = np.random.default_rng(314)
    theta = 11.96
    np.random.seed(1)
    sp = LS.Spectrum(['Halpha'], 'sinc', [1.0], 0.0, 30.0, 'SN3', 5000, 50.0)
    axis, _ = sp.create_spectrum()
    axis32 = np.asarray(axis, dtype=np.float32)
    corr = 1.0 / np.cos(np.deg2rad(theta))
    sinc_w = corr * 1e7 / (2 * sp.delta_x * sp.n_steps)
    n = 12
    rows, truth, out = [], [], dict(vel=[], vel_err=[], fit_sol=[], unc=[])
    for i in range(n):
        v = float(rng.uniform(30, 130))
        snr = float(np.exp(rng.uniform(np.log(15), np.log(150))))
        p = 1e7 / (OH_NM * (1 + v / 299792))
        spec = np.asarray(Sinc().function(axis32.astype(np.float64), [1.0, p, 1.0], sinc_w), dtype=np.float64) + 0.08 + rng.normal(0, 1.0 / snr, len(axis32))
        rows.append((spec * 3e-17).astype(np.float32))
        truth.append(v)
    cube = np.stack(rows)
    for i in range(n):
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            fit = Fit(cube[i].astype(np.float64), axis32, None, 'sinc', ['OH_649'], [1], [1], trans_filter=None, theta=theta,
                      delta_x=sp.delta_x, n_steps=sp.n_steps, zpd_index=0, uncertainty_bool=True, filter='SN3', ML_bool=False,
                      bayes_bool=False, bayes_method='emcee', sky_lines={'OH_0': OH_NM}, sky_lines_scale=[1], resolution=5000,
                      Luci_path=None)
            vel, vel_err, fit_vector = fit.fit(sky_line=True)
        out['vel'].append(float(vel)); out['vel_err'].append(float(vel_err))
        out['fit_sol'].append(np.array(fit.fit_sol, dtype=np.float64)); out['unc'].append(np.array(fit.uncertainties, dtype=np.float64))
        print(i, 'truth %.2f  reference %.3f +- %.3f' % (truth[i], vel, vel_err))
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'skyline_sn3.npz'), cube=cube, axis32=axis32,
                        theta=np.full(n, theta, np.float32), truth=np.array(truth), n_steps=sp.n_steps, delta_x=sp.delta_x,
                        oh_nm=OH_NM, ref_velocity=np.array(out['vel']), ref_velocity_error=np.array(out['vel_err']),
                        ref_fit_sol=np.stack(out['fit_sol']), ref_fit_unc=np.stack(out['unc']))


if __name__ == '__main__':
    main()
