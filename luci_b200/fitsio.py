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


class WcslibStr(str):
    """A string card value formatted the way WCSLIB writes it (no padding to 8 characters inside the quotes): astropy's
    ``WCS.to_header()`` keeps WCSLIB's card images."""


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
    elif isinstance(value, WcslibStr):
        val = ("'%s'" % str(value).replace("'", "''")).ljust(20)
    else:
        s = str(value).replace("'", "''")
        val = ("'%-8s'" % s).ljust(20)
    card = '%-8s= %s' % (key, val)
    if comment:
        card += ' / ' + str(comment)
    return card[:80].ljust(80)


def _header_block(shape, bitpix, header):
    """The padded header of a primary HDU holding an array of ``shape`` (numpy order)."""
    cards = [_card('SIMPLE', True, 'conforms to FITS standard'), _card('BITPIX', bitpix, 'array data type'),
             _card('NAXIS', len(shape), 'number of array dimensions')]
    for i, n in enumerate(reversed(shape)):
        cards.append(_card('NAXIS%d' % (i + 1), n))
    if header:
        for k, v in header.items():
            if str(k).upper() in _STRUCTURAL or v is None:
                continue
            if isinstance(v, tuple):                      # (value, comment), e.g. from wcs_to_header
                cards.append(_card(k, v[0], v[1]))
            else:
                cards.append(_card(k, v))
    cards.append('END'.ljust(80))
    hdr = ''.join(cards).encode('ascii')
    return hdr + b' ' * ((-len(hdr)) % _BLOCK)


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
    raw = out.tobytes()
    raw += b'\0' * ((-len(raw)) % _BLOCK)
    with open(path, 'wb') as f:
        f.write(_header_block(data.shape, bitpix, header))
        f.write(raw)


def write_fits_be(path, be_data, shape, header=None, bitpix=-32):
    """Like :func:`write_fits` for a data unit that is ALREADY big-endian (a contiguous buffer of ``prod(shape)``
    4-byte values, e.g. a row of ``lucifit_pack_maps``' byte-reversed output): header + the buffer as it is."""
    view = memoryview(be_data).cast('B')
    if len(view) != int(np.prod(shape)) * (abs(bitpix) // 8):
        raise ValueError('buffer does not hold %s values of %d bits' % (shape, abs(bitpix)))
    with open(path, 'wb') as f:
        f.write(_header_block(tuple(shape), bitpix, header))
        f.write(view)
        f.write(b'\0' * ((-len(view)) % _BLOCK))


# ---- what `astropy.wcs.WCS(header, naxis=2).to_header()` writes for a 2-D celestial WCS (the header of every map the
# reference saves: LuciBase.py:511-513 `Cutout2D(...).wcs.to_header()`, LuciUtility.py:69-85).  WCSLIB's canonical card
# order and comments, pinned card for card on the reference's own example output
# Data/ExampleData/Luci_outputs/Velocity/NGC628_1_sincgauss_Halpha_velocity.fits (tests/test_fits_wcs.py).
_PROJ_NAMES = {'TAN': 'gnomonic', 'SIN': 'orthographic/synthesis', 'ARC': 'zenithal/azimuthal equidistant',
               'STG': 'stereographic', 'ZEA': "zenithal/azimuthal equal area", 'CAR': 'plate caree', 'AIT': "Hammer-Aitoff",
               'SFL': 'Sanson-Flamsteed', 'MER': "Mercator", 'CEA': 'cylindrical equal area'}
_ZENITHAL = ('TAN', 'SIN', 'ARC', 'STG', 'ZEA', 'AZP', 'SZP', 'ZPN', 'AIR')


def _ctype_comment(ctype):
    c = str(ctype)
    proj = c[5:8] if len(c) >= 8 and c[4] == '-' else ''
    kind = {'RA--': 'Right ascension', 'DEC-': 'Declination', 'GLON': 'galactic longitude', 'GLAT': 'galactic latitude',
            'ELON': 'ecliptic longitude', 'ELAT': 'ecliptic latitude'}.get(c[:4])
    if kind and proj in _PROJ_NAMES:
        return '%s, %s projection' % (kind, _PROJ_NAMES[proj])
    return None


def _mjd_of_date(date):
    """MJD of an ISO-8601 date / date-time string (what WCSLIB derives MJD-OBS from when only DATE-OBS is given)."""
    import datetime
    d = str(date).strip()
    try:
        t = datetime.datetime.fromisoformat(d)
    except ValueError:
        return None
    return (t - datetime.datetime(1858, 11, 17)).total_seconds() / 86400.0


def wcs_to_header(header, origin=(0, 0)):
    """Canonical 2-D celestial WCS cards ``{key: (value, comment)}`` of ``header`` (a dict of FITS cards), with the
    reference pixel moved by a cut-out whose first pixel is ``origin = (x0, y0)`` of the parent image
    (``Cutout2D``: ``wcs.wcs.crpix -= origin``).  An empty / WCS-less header gives WCSLIB's defaults (CRPIX 0, CDELT 1,
    linear axes), which is what the reference writes for cubes without celestial keywords.

    Not reproduced (documented in INTEGRATION.md): distortion keywords (SIP / TPV), alternate WCS versions (``CRPIX1A``),
    ``PVi_m`` / ``PSi_m`` parameters, and the time-system cards astropy >= 4.3 adds beyond MJDREF (``TIMESYS`` ...)."""
    h = {str(k).upper(): (v[0] if isinstance(v, tuple) else v) for k, v in (header or {}).items()}
    out = OrderedDict()
    out['WCSAXES'] = (2, 'Number of coordinate axes')
    for i in (1, 2):
        out['CRPIX%d' % i] = (float(h.get('CRPIX%d' % i, 0.0)) - float(origin[i - 1]), 'Pixel coordinate of reference point')
    have_cd = any(('CD%d_%d' % (i, j)) in h for i in (1, 2) for j in (1, 2))
    pc = [[1.0, 0.0], [0.0, 1.0]]
    for i in (1, 2):
        for j in (1, 2):
            if have_cd:
                pc[i - 1][j - 1] = float(h.get('CD%d_%d' % (i, j), 0.0))
            elif ('PC%d_%d' % (i, j)) in h:
                pc[i - 1][j - 1] = float(h['PC%d_%d' % (i, j)])
    if pc != [[1.0, 0.0], [0.0, 1.0]]:                    # WCSLIB omits an identity matrix
        for i in (1, 2):
            for j in (1, 2):
                out['PC%d_%d' % (i, j)] = (pc[i - 1][j - 1], 'Coordinate transformation matrix element')
    celestial = all(_ctype_comment(h.get('CTYPE%d' % i, '')) for i in (1, 2))
    unit = 'deg' if celestial else str(h.get('CUNIT1', '')).strip()
    for i in (1, 2):
        cd = 1.0 if have_cd else float(h.get('CDELT%d' % i, 1.0))
        out['CDELT%d' % i] = (cd, ('[%s] ' % unit if unit else '') + 'Coordinate increment at reference point')
    if unit:
        for i in (1, 2):
            out['CUNIT%d' % i] = (WcslibStr(unit), 'Units of coordinate increment and value')
    for i in (1, 2):
        ct = h.get('CTYPE%d' % i)
        if ct:
            out['CTYPE%d' % i] = (WcslibStr(ct), _ctype_comment(ct))
    for i in (1, 2):
        out['CRVAL%d' % i] = (float(h.get('CRVAL%d' % i, 0.0)), ('[%s] ' % unit if unit else '') + 'Coordinate value at reference point')
    if celestial:
        proj = str(h['CTYPE1'])[5:8]
        crval2 = float(h.get('CRVAL2', 0.0))
        zen = proj in _ZENITHAL
        lonpole = float(h['LONPOLE']) if 'LONPOLE' in h else (180.0 if (zen and crval2 < 90.0) or (not zen and crval2 <= 0.0) else 0.0)
        latpole = float(h['LATPOLE']) if 'LATPOLE' in h else (crval2 if zen else 90.0)
        out['LONPOLE'] = (lonpole, '[deg] Native longitude of celestial pole')
        out['LATPOLE'] = (latpole, '[deg] Native latitude of celestial pole')
    out['MJDREF'] = (float(h.get('MJDREF', 0.0)), '[d] MJD of fiducial time')
    date = h.get('DATE-OBS')
    mjd = h.get('MJD-OBS')
    if date is not None:
        out['DATE-OBS'] = (WcslibStr(date), 'ISO-8601 time of observation')
        if mjd is None:
            mjd = _mjd_of_date(date)
    if mjd is not None:
        out['MJD-OBS'] = (float(mjd), '[d] MJD of observation')
    if celestial and str(h.get('CTYPE1', ''))[:4] == 'RA--':
        equinox = h.get('EQUINOX', h.get('EPOCH'))
        radesys = h.get('RADESYS', h.get('RADECSYS'))
        if radesys is None:
            radesys = 'ICRS' if equinox is None else ('FK4' if float(equinox) < 1984.0 else 'FK5')
        out['RADESYS'] = (WcslibStr(str(radesys).strip()), 'Equatorial coordinate system')
        if equinox is None and str(radesys).strip() in ('FK5', 'FK4', 'FK4-NO-E'):
            equinox = 2000.0 if str(radesys).strip() == 'FK5' else 1950.0
        if equinox is not None:
            out['EQUINOX'] = (float(equinox), '[yr] Equinox of equatorial coordinates')
    return out


def cutout_origin(x_min, x_max, y_min, y_max, nx_full, ny_full, strict_reference=True):
    """First pixel ``(x0, y0)`` of the cut-out whose WCS the reference writes (``LuciBase.py:511-513``):
    ``Cutout2D(deep, position=((x_max+x_min)/2, (y_max+y_min)/2), size=(x_max-x_min, y_max-y_min), wcs=wcs)``.
    ``Cutout2D`` reads ``size`` as ``(ny, nx)``, so the reference's cut-out is ``y_max - y_min`` pixels wide in x and
    ``x_max - x_min`` tall in y around the right centre: for a square region that is the fitted rectangle, otherwise its
    origin is displaced (``astropy.nddata.utils.overlap_slices``: ``ceil(centre - size / 2)``, trimmed to the image).
    ``strict_reference=False`` returns the origin of the fitted rectangle itself."""
    if not strict_reference:
        return float(x_min), float(y_min)
    import math
    xc, yc = (x_max + x_min) / 2.0, (y_max + y_min) / 2.0
    ncols, nrows = int(y_max - y_min), int(x_max - x_min)           # size = (rows, cols) as Cutout2D reads it
    x0 = int(math.ceil(xc - ncols / 2.0))
    y0 = int(math.ceil(yc - nrows / 2.0))
    x0 = min(max(x0, 0), max(int(nx_full) - 1, 0))
    y0 = min(max(y0, 0), max(int(ny_full) - 1, 0))
    return float(x0), float(y0)


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


def read_cards(path):
    """The raw 80-character header cards of the primary HDU (up to and excluding END)."""
    cards = []
    with open(path, 'rb') as f:
        while True:
            block = f.read(_BLOCK)
            if len(block) < _BLOCK:
                raise ValueError('truncated FITS header')
            for i in range(0, _BLOCK, 80):
                card = block[i:i + 80].decode('ascii', 'replace')
                if card[:8].strip() == 'END':
                    return cards
                cards.append(card)


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


def map_jobs(output_dir, object_name, lines, binning=1, suffix='', fit_function=None):
    """``(path, field, line index or None)`` of every file ``LUCI.LuciUtility.save_fits`` writes (LuciUtility.py:22-85):
    same directory layout and file names; ``field`` is the name of the quantity in ``engine.ROW_FIELDS`` / ``ROW_SCALARS``."""
    for sub in ('Amplitudes', 'Fluxes', 'Velocity', 'Broadening'):
        os.makedirs(os.path.join(output_dir, sub), exist_ok=True)
    output_name = object_name + suffix
    if binning is not None:
        output_name += "_" + str(binning)
    if fit_function is not None:
        output_name += "_" + fit_function
    lines_fit = []
    jobs = []                     # 7 L + 3 independent files
    for ct, line_ in enumerate(lines):
        # NOTE: the reference never appends to its `lines_fit` list (LuciUtility.py:63-68), so repeated line names
        # overwrite one another there; here the second component gets the documented `_2` suffix.
        if lines_fit.count(line_) >= 1:
            name = line_ + '_' + str(lines_fit.count(line_) + 1)
        else:
            name = line_
        lines_fit.append(line_)
        jobs.append((output_dir + '/Amplitudes/' + output_name + '_' + name + '_Amplitude.fits', 'amplitudes', ct))
        jobs.append((output_dir + '/Fluxes/' + output_name + '_' + name + '_Flux.fits', 'fluxes', ct))
        jobs.append((output_dir + '/Fluxes/' + output_name + '_' + name + '_Flux_err.fits', 'flux_errors', ct))
        jobs.append((output_dir + '/Velocity/' + output_name + '_' + name + '_velocity.fits', 'velocities', ct))
        jobs.append((output_dir + '/Broadening/' + output_name + '_' + name + '_broadening.fits', 'sigmas', ct))
        jobs.append((output_dir + '/Velocity/' + output_name + '_' + name + '_velocity_err.fits', 'vels_errors', ct))
        jobs.append((output_dir + '/Broadening/' + output_name + '_' + name + '_broadening_err.fits', 'sigmas_errors', ct))
    jobs.append((output_dir + '/' + output_name + '_Chi2.fits', 'chi2', None))
    jobs.append((output_dir + '/' + output_name + '_continuum.fits', 'continuum', None))
    jobs.append((output_dir + '/' + output_name + '_continuum_error.fits', 'continuum_error', None))
    return jobs


def _run_jobs(fn, jobs, n_values):
    if n_values < (1 << 22):
        for job in jobs:
            fn(job)
    else:                         # large maps: the byte swap and the file writes release the GIL
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
            list(pool.map(fn, jobs))


def save_fits(output_dir, object_name, lines, ampls_fits, flux_fits, flux_errors_fits, velocities_fits,
              broadenings_fits, velocities_errors_fits, broadenings_errors_fits, chi2_fits, continuum_fits,
              continuum_error_fits, header, binning=1, suffix='', fit_function=None):
    """Same arguments, directory layout and file names as ``LUCI.LuciUtility.save_fits`` (LuciUtility.py:22-85)."""
    maps = {'amplitudes': ampls_fits, 'fluxes': flux_fits, 'flux_errors': flux_errors_fits, 'velocities': velocities_fits,
            'sigmas': broadenings_fits, 'vels_errors': velocities_errors_fits, 'sigmas_errors': broadenings_errors_fits,
            'chi2': chi2_fits, 'continuum': continuum_fits, 'continuum_error': continuum_error_fits}
    jobs = map_jobs(output_dir, object_name, lines, binning, suffix, fit_function)
    pick = lambda field, k: maps[field] if k is None else maps[field][:, :, k]
    _run_jobs(lambda job: write_fits(job[0], pick(job[1], job[2]), header), jobs,
              sum(np.asarray(pick(f, k)).size for _, f, k in jobs))


def save_fits_be(output_dir, object_name, lines, be_maps, field_index, shape, header, binning=1, suffix='', fit_function=None):
    """The same files from big-endian data units that already exist: ``be_maps [K, ny * nx]`` (4-byte values, every row
    one 2-D map ``shape = (ny, nx)`` as ``lucifit_pack_maps`` lays them out) and ``field_index(field, line) -> row``.
    Nothing is converted or copied on the host; the writes run in a thread pool."""
    jobs = map_jobs(output_dir, object_name, lines, binning, suffix, fit_function)
    _run_jobs(lambda job: write_fits_be(job[0], be_maps[field_index(job[1], job[2])], shape, header), jobs,
              len(jobs) * int(np.prod(shape)))
