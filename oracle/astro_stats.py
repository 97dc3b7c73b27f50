"""TEST INFRASTRUCTURE ONLY -- restatement of the two ``astropy.stats`` routines the reference calls.

``sigma_clip(..., masked=False, axis=None)`` and ``mad_std`` as used by ``Fit.cont_estimate``
(``LUCI/LuciFit.py:478-480``) and ``Luci.create_snr_map.SNR_calc`` (``LuciBase.py:1146-1147``), following
astropy 6.0 (pinned in ``luci.yml:32``): flatten, drop non-finite values, then up to ``maxiters`` times
compute centre (median by default) and spread (``stdfunc``; default population std) and keep
``centre - sigma*spread <= x <= centre + sigma*spread``; stop early when an iteration removes nothing.
astropy is not installed here (no network), so this is restated from its documented behaviour.
"""
import numpy as np

_MAD_TO_STD = 1.482602218505602  # 1 / Phi^-1(3/4)


def mad_std(data, **_ignored):
    data = np.asarray(data, dtype=float)
    med = np.nanmedian(data)
    return _MAD_TO_STD * np.nanmedian(np.abs(data - med))


def sigma_clip(data, sigma=3, sigma_lower=None, sigma_upper=None, maxiters=5, cenfunc="median",
               stdfunc="std", axis=None, masked=True, return_bounds=False, copy=True, grow=False):
    if axis is not None or masked is not False:
        raise NotImplementedError("only the masked=False, axis=None form is restated")
    if cenfunc in ("median", None):
        cenfunc = np.nanmedian
    elif cenfunc == "mean":
        cenfunc = np.nanmean
    if stdfunc in ("std", None):
        stdfunc = np.nanstd
    elif stdfunc == "mad_std":
        stdfunc = mad_std
    lo_s = sigma if sigma_lower is None else sigma_lower
    hi_s = sigma if sigma_upper is None else sigma_upper
    d = np.asarray(data, dtype=float).ravel()
    d = d[np.isfinite(d)]
    it = 0
    while maxiters is None or it < maxiters:
        n = d.size
        if n == 0:
            break
        c = cenfunc(d)
        s = stdfunc(d)
        d = d[(d >= c - lo_s * s) & (d <= c + hi_s * s)]
        it += 1
        if d.size == n:
            break
    return d
