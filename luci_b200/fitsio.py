"""Minimal FITS primary-HDU writer/reader (astropy is not available in this environment).

Writes what ``astropy.io.fits.writeto(path, data, header, overwrite=True)`` writes for the 2-D maps of
``LUCI/LuciUtility.py:69-85`` and ``LuciBase.py:214, 1184``: a single primary HDU, BITPIX -32 / -64, NAXIS1 = last
numpy axis, followed by the caller's header cards (WCS keywords of the cut-out).
"""
import os
from collections import OrderedDict

import numpy as np

_BLOCK = 2880
_STRUCTURAL = ('SIMPLE', 'BITPIX', 'NAXIS', 'NAXIS1', 'NAXIS2', 'NAXIS3', 'EXTEND', 'END')


def _card(key, value, comment=None):
    key = str(key).upper()[:8]
    if isinstance(value, (bool, np.bool_)):
        val = ('T' if value else 'F').rjust(20)
    elif isinstance(value, (int, np.integer)):
        val = str(int(value)).rjust(20)
    elif isinstance(value, (float, np.floating)):
        v = float(value)
        s = repr(v).upper() if (abs(v) >= 1e16 or (v != 0 and abs(v) < 1e-4)) else repr(v)
        if 'E' not in s and '.' not in s and 'N' not in s:
            s += '.0'
        val = s.rjust(20)
    else:
        s = str(value).replace("'", "''")
        val = ("'%-8s'" % s).ljust(20)
    card = '%-8s= %s' % (key, val)
    if comment:
        card += ' / ' + str(comment)
    return card[:80].ljust(80)


def write_fits(path, data, header=None, overwrite=True):
    data = np.asarray(data)
    if data.dtype == np.float32:
        bitpix, out = -32, data.astype('>f4')
    elif data.dtype in (np.float64, np.float16):
        bitpix, out = -64, data.astype('>f8')
    elif data.dtype in (np.int32, np.int64, np.bool_, np.uint8, np.int16):
        bitpix, out = 32, data.astype('>i4')
    else:
        raise TypeError('unsupported dtype %s' % data.dtype)
    if os.path.exists(path) and not overwrite:
        raise OSError('%s exists' % path)
    cards = [_card('SIMPLE', True, 'conforms to FITS standard'), _card('BITPIX', bitpix, 'array data type'),
             _card('NAXIS', data.ndim, 'number of array dimensions')]
    for i, n in enumerate(reversed(data.shape)):
        cards.append(_card('NAXIS%d' % (i + 1), n))
    if header:
        for k, v in header.items():
            if str(k).upper() in _STRUCTURAL or v is None:
                continue
            cards.append(_card(k, v))
    cards.append('END'.ljust(80))
    hdr = ''.join(cards).encode('ascii')
    hdr += b' ' * ((-len(hdr)) % _BLOCK)
    raw = out.tobytes()
    raw += b'\0' * ((-len(raw)) % _BLOCK)
    with open(path, 'wb') as f:
        f.write(hdr)
        f.write(raw)


def _parse_value(s):
    s = s.strip()
    if s.startswith("'"):
        end = s.find("'", 1)
        while end != -1 and s[end:end + 2] == "''":
            end = s.find("'", end + 2)
        return s[1:end].rstrip().replace("''", "'")
    s = s.split('/')[0].strip()
    if s == 'T':
        return True
    if s == 'F':
        return False
    try:
        return int(s)
    except ValueError:
        try:
            return float(s.replace('D', 'E'))
        except ValueError:
            return s


def read_fits(path):
    """Return (data, header OrderedDict) of the primary HDU."""
    with open(path, 'rb') as f:
        header = OrderedDict()
        done = False
        while not done:
            block = f.read(_BLOCK)
            if len(block) < _BLOCK:
                raise ValueError('truncated FITS header')
            for i in range(0, _BLOCK, 80):
                card = block[i:i + 80].decode('ascii', 'replace')
                key = card[:8].strip()
                if key == 'END':
                    done = True
                    break
                if card[8:10] == '= ':
                    header[key] = _parse_value(card[10:])
        naxis = int(header.get('NAXIS', 0))
        shape = tuple(int(header['NAXIS%d' % (i + 1)]) for i in range(naxis))[::-1]
        bitpix = int(header['BITPIX'])
        dt = {-32: '>f4', -64: '>f8', 32: '>i4', 16: '>i2', 8: 'u1', 64: '>i8'}[bitpix]
        count = int(np.prod(shape)) if shape else 0
        data = np.frombuffer(f.read(count * np.dtype(dt).itemsize), dtype=dt).reshape(shape)
    return data.astype(np.dtype(dt).newbyteorder('=')), header


def save_fits(output_dir, object_name, lines, ampls_fits, flux_fits, flux_errors_fits, velocities_fits,
              broadenings_fits, velocities_errors_fits, broadenings_errors_fits, chi2_fits, continuum_fits,
              continuum_error_fits, header, binning=1, suffix='', fit_function=None):
    """Same directory layout and file names as ``LUCI.LuciUtility.save_fits`` (LuciUtility.py:22-85)."""
    for sub in ('Amplitudes', 'Fluxes', 'Velocity', 'Broadening'):
        os.makedirs(os.path.join(output_dir, sub), exist_ok=True)
    output_name = object_name + suffix
    if binning is not None:
        output_name += "_" + str(binning)
    if fit_function is not None:
        output_name += "_" + fit_function
    lines_fit = []
    jobs = []                     # (path, 2-D map): 7 L + 3 independent files
    for ct, line_ in enumerate(lines):
        # NOTE: the reference never appends to its `lines_fit` list (LuciUtility.py:63-68), so repeated line names
        # overwrite one another there; here the second component gets the documented `_2` suffix.
        if lines_fit.count(line_) >= 1:
            name = line_ + '_' + str(lines_fit.count(line_) + 1)
        else:
            name = line_
        lines_fit.append(line_)
        jobs.append((output_dir + '/Amplitudes/' + output_name + '_' + name + '_Amplitude.fits', ampls_fits[:, :, ct]))
        jobs.append((output_dir + '/Fluxes/' + output_name + '_' + name + '_Flux.fits', flux_fits[:, :, ct]))
        jobs.append((output_dir + '/Fluxes/' + output_name + '_' + name + '_Flux_err.fits', flux_errors_fits[:, :, ct]))
        jobs.append((output_dir + '/Velocity/' + output_name + '_' + name + '_velocity.fits', velocities_fits[:, :, ct]))
        jobs.append((output_dir + '/Broadening/' + output_name + '_' + name + '_broadening.fits', broadenings_fits[:, :, ct]))
        jobs.append((output_dir + '/Velocity/' + output_name + '_' + name + '_velocity_err.fits', velocities_errors_fits[:, :, ct]))
        jobs.append((output_dir + '/Broadening/' + output_name + '_' + name + '_broadening_err.fits', broadenings_errors_fits[:, :, ct]))
    jobs.append((output_dir + '/' + output_name + '_Chi2.fits', chi2_fits))
    jobs.append((output_dir + '/' + output_name + '_continuum.fits', continuum_fits))
    jobs.append((output_dir + '/' + output_name + '_continuum_error.fits', continuum_error_fits))
    if sum(np.asarray(m).size for _, m in jobs) < (1 << 22):
        for path, m in jobs:
            write_fits(path, m, header)
    else:                         # large maps: the byte swap and the file writes release the GIL
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
            list(pool.map(lambda job: write_fits(job[0], job[1], header), jobs))
