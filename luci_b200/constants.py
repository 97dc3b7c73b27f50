"""Constant tables of the LUCI fit path, shared by the host glue (one table instead of the reference's
``if/elif`` ladders).

Sources in the reference: line rest wavelengths ``LUCI/LuciFit.py:90-99``; fit / noise / continuum windows
``LUCI/LuciFit.py:233-270, 291-325, 452-472``; SNR windows ``LuciBase.py:1083-1106``; LuciSim interferometer
parameters ``LUCI/LuciSim.py:53-73``.
"""

SPEED_OF_LIGHT = 299792  # km/s -- integer in the reference (LuciFit.py:26)

LINE_DICT = {
    'Halpha': 656.280, 'NII6583': 658.341, 'NII6548': 654.803,
    'SII6716': 671.647, 'SII6731': 673.085, 'OII3726': 372.603,
    'OII3729': 372.882, 'OIII4959': 495.891, 'OIII5007': 500.684,
    'Hbeta': 486.133, 'OH': 649.873, 'HalphaC4': 807.881, 'NII6583C4': 810.417,
    'NII6548C4': 804.7,
    'OIII5007C2': 616.342, 'OIII4959C2': 610.441821, 'HbetaC2': 598.429723,
    'OII3729C1': 459.017742, 'OII3726C1': 458.674293, 'OI6364': 636.3776,
    'FeXIV5303': 530.286, 'NI5200': 520.026, 'FeVII5158': 515.89, 'HeII5411': 541.152,
    'FeXI6624': 662.43, 'NiXV6703': 670.332,
}

AVAILABLE_FUNCTIONS = ['gaussian', 'sinc', 'sincgauss', 'gauss']
MODEL_CODE = {'gaussian': 0, 'gauss': 0, 'sinc': 1, 'sincgauss': 2}


def fit_window(filter_, lines=(), obj_redshift_corr=1.0):
    """(spec_min, spec_max) in cm^-1 -- LuciFit.py:233-270."""
    if filter_ == 'SN3':
        return 14750, 15400
    if filter_ == 'SN2':
        return 19500, 20750
    if filter_ == 'SN1':
        return 26000, 28000
    if filter_ == 'C3':
        return (26000, 29000) if 'OII3726' in lines else (18100, 19500)
    if filter_ == 'C4':
        return 12150 * obj_redshift_corr, 12550 * obj_redshift_corr
    if filter_ == 'C2':
        return 15990 * obj_redshift_corr, 17880 * obj_redshift_corr
    if filter_ == 'C1':
        return 20408 * obj_redshift_corr, 25974 * obj_redshift_corr
    raise ValueError('The filter of your datacube is not supported by LUCI. We only support C1, C2, C3, C4, '
                     'SN1, SN2, and SN3 at the moment.')


def noise_window(filter_, lines=(), obj_redshift_corr=1.0):
    """Out-of-filter window whose standard deviation is the noise -- LuciFit.py:291-325."""
    if filter_ == 'SN3':
        return 15600, 15800
    if filter_ == 'SN2':
        return 18600, 19000
    if filter_ == 'SN1':
        return 26000, 26200
    if filter_ == 'C3' and 'OII3726' in lines:
        return 26000, 26200
    if filter_ == 'C4' and 'Halpha' in lines:
        return 11800 * obj_redshift_corr, 12150 * obj_redshift_corr
    if filter_ == 'C2':
        return 15500 * obj_redshift_corr, 15990 * obj_redshift_corr
    if filter_ == 'C1':
        return 18000 * obj_redshift_corr, 20665 * obj_redshift_corr
    raise ValueError('noise window undefined for filter %r with these lines (the reference leaves it unset too)'
                     % (filter_,))


def continuum_window(filter_):
    """Window of the sigma-clipped continuum guess -- LuciFit.py:450-472."""
    table = {'SN3': (14850, 14950), 'SN2': (19500, 19550), 'SN1': (26000, 26250), 'C4': (12180, 12550),
             'C2': (16000, 17800), 'C1': (21500, 25900), 'C3': (18000, 19000)}
    return table.get(filter_, (0, 1e6))


def snr_windows(filter_, lines=(None,)):
    """((flux_min, flux_max), (noise_min, noise_max)) -- LuciBase.py:1083-1106."""
    if filter_ == 'SN3':
        return (15150, 15300), (14500, 14600)
    if filter_ == 'SN2':
        if 'OIII' in lines:
            return (1e7 / 505, 1e7 / 495), (19000, 19500)
        return (1e7 / 486, 1e7 / 482), (19000, 19500)
    if filter_ == 'SN1':
        return (26550, 27550), (25700, 26300)
    if filter_ == 'C3':
        return (18500, 20500), (20500, 21500)
    raise ValueError('SNR Calculation for this filter has not been implemented')


# (STEP [nm], folding order) used by LuciSim -- LuciSim.py:53-73
SIM_FILTERS = {'SN3': (2943, 8), 'SN2': (1680, 6), 'SN1': (1647, 8), 'C1': (570, 2), 'C2': (1680, 5),
               'C3': (1778, 6), 'C4': (5272, 12)}
