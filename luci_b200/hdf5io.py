"""Ingest of an ORBS/SITELLE HDF5 cube -- the host side of ``Luci.__init__`` / ``Luci.read_in_cube``
(``LuciBase.py:51-145``) and of the ``LUCI/LuciUtility.py`` readers it calls.

The reference assembles a float64 ``cube_final[x, y, z]`` on the host, quadrant by quadrant, clamping each one with
three boolean-mask passes (``LuciBase.py:123-129``).  Here every quadrant goes straight to its place in the float32
cube in HBM and the clamp is ONE in-place device pass (``lucifit_sanitize_cube``).

``h5py`` is only needed to open the file (:func:`open_hdf5`); every function below takes the opened file object, so
anything with the same mapping interface (``file.attrs``, ``file[name]``) works.  astropy is optional: without it the
WCS header is assembled from the celestial keywords of the HDF5 header directly.
"""
import numpy as np
import torch

from . import engine

CLAMP_LO, CLAMP_HI, CLAMP_FILL = -1e-16, 1e-9, 1e-22          # LuciBase.py:127-129


def open_hdf5(cube_path):
    """``h5py.File(cube_path + '.hdf5', 'r')`` (``LuciBase.py:115``)."""
    try:
        import h5py
    except ImportError as exc:
        raise ImportError('h5py is needed to read %s.hdf5; build the object from arrays with Luci.from_arrays(...) '
                          'otherwise' % cube_path) from exc
    return h5py.File(cube_path + '.hdf5', 'r')


def get_quadrant_dims(quad_number, quad_nb, dimx, dimy):
    """``LUCI/LuciUtility.py:88-117``: bounds of quadrant ``quad_number`` of a ``sqrt(quad_nb)`` x ``sqrt(quad_nb)``
    tiling; x runs fastest, the last row / column takes the remainder."""
    div_nb = int(np.sqrt(quad_nb))
    if quad_number < 0 or quad_number > quad_nb - 1:
        raise Exception('quad_number out of bounds [0,' + str(quad_nb - 1) + ']')
    index_x = quad_number % div_nb
    index_y = (quad_number - index_x) // div_nb
    step_x, step_y = dimx // div_nb, dimy // div_nb
    x_min, y_min = index_x * step_x, index_y * step_y
    x_max = (index_x + 1) * step_x if index_x != div_nb - 1 else dimx
    y_max = (index_y + 1) * step_y if index_y != div_nb - 1 else dimy
    return int(x_min), int(x_max), int(y_min), int(y_max)


def read_cube(file, device):
    """``Luci.read_in_cube`` (``LuciBase.py:106-134``): float32 ``[dimx, dimy, dimz]`` tensor on ``device``.

    Old-style files hold ``quad_nb`` quadrants ``quad00<i>/data``; each is copied to its slab of the cube and the cube
    is clamped in place afterwards (element-wise, so clamping after assembly equals clamping every quadrant).  Files
    without the quadrant attributes hold one ``data`` set whose real part is taken unclamped (``:132-133``)."""
    attrs = file.attrs
    if all(k in attrs for k in ('quad_nb', 'dimx', 'dimy', 'dimz')):
        quad_nb, dimx, dimy, dimz = (int(attrs[k]) for k in ('quad_nb', 'dimx', 'dimy', 'dimz'))
        cube = torch.zeros((dimx, dimy, dimz), dtype=torch.float32, device=device)
        for iquad in range(quad_nb):
            xmin, xmax, ymin, ymax = get_quadrant_dims(iquad, quad_nb, dimx, dimy)
            quad = np.asarray(file['quad00%i' % iquad]['data'][:])
            cube[xmin:xmax, ymin:ymax, :] = torch.from_numpy(np.ascontiguousarray(quad, dtype=np.float32)).to(device)
        engine.sanitize_cube(cube, CLAMP_LO, CLAMP_HI, CLAMP_FILL)
        return cube
    data = np.real(np.asarray(file['data']))
    return torch.from_numpy(np.ascontiguousarray(data, dtype=np.float32)).to(device)


def _clean(val):
    return str(val).replace("'b", '').replace("'", '').replace('b', '')


def header_dict(file):
    """The ``hdr_dict`` of ``LUCI/LuciUtility.update_header`` (``:168-260``) for both HDF5 layouts: old files carry a
    ``header`` table of ``(key, value, comment, type)`` rows, new files carry the cards as file attributes."""
    hdr = {}
    attrs = file.attrs
    if 'quad_nb' in list(attrs):                                   # old HDF5 (:183-201)
        for row in list(file['header'][()]):
            col, val, typ = _clean(row[0]), _clean(row[1]), str(row[3])
            if 'bool' in typ:
                hdr[col] = bool(val)
            elif 'float' in typ:
                hdr[col] = float(val)
            elif 'int' in typ:
                hdr[col] = int(val)
            else:
                try:
                    hdr[col] = float(val)
                except ValueError:
                    hdr[col] = str(val)
    else:                                                          # new HDF5 (:202-231)
        for col in list(attrs):
            val = attrs[col]
            if isinstance(val, (np.floating, float)):
                hdr[col] = float(val)
            elif isinstance(val, (np.bool_, bool)):
                hdr[col] = bool(val)
            elif isinstance(val, (np.integer, int)):
                hdr[col] = int(val)
            elif isinstance(val, np.ndarray):
                hdr[col] = np.array(val)
            else:
                hdr[col] = '' if 'path' in col else (val.decode() if isinstance(val, bytes) else str(val))
    hdr['CTYPE3'] = 'WAVE-SIP'
    hdr['CUNIT3'] = 'm'
    for key in ('A_ORDER', 'AP_ORDER'):
        if key in hdr:
            hdr[key] = int(hdr[key])
    if 'NAXIS' in hdr:
        hdr['NAXIS'] = int(hdr['NAXIS'])
    if 'NAXIS1' not in hdr:
        hdr['NAXIS1'], hdr['NAXIS2'] = 2048, 2064
    # new-style lower-case keys mapped onto the names the fit path reads (:250-259; the reference tests 'STEP_NB' but
    # reads 'step_nb' -- a KeyError there; either spelling is accepted here)
    for src, dst, conv in (('step_nb', 'STEPNB', int), ('STEP_NB', 'STEPNB', int), ('zpd_index', 'ZPDINDEX', int),
                           ('filter_name', 'FILTER', str), ('axis_min', 'CRVAL3', float), ('axis_step', 'CDELT3', float),
                           ('step', 'STEP', float)):
        if src in hdr and (dst not in hdr or src != 'step'):
            hdr[dst] = conv(hdr[src])
    return hdr


_WCS_KEYS = ('WCSAXES', 'CRPIX1', 'CRPIX2', 'PC1_1', 'PC1_2', 'PC2_1', 'PC2_2', 'CD1_1', 'CD1_2', 'CD2_1', 'CD2_2',
             'CDELT1', 'CDELT2', 'CUNIT1', 'CUNIT2', 'CTYPE1', 'CTYPE2', 'CRVAL1', 'CRVAL2', 'LONPOLE', 'LATPOLE',
             'RADESYS', 'EQUINOX')


def wcs_header(hdr):
    """The 2-D celestial header the reference obtains from ``WCS(hdr_dict, naxis=2).to_header()``
    (``LuciUtility.py:244-248``).  With astropy installed exactly that; otherwise the celestial cards of ``hdr`` are
    carried over, a ``CDi_j`` matrix rewritten as ``PCi_j`` with unit ``CDELT`` the way ``to_header`` does."""
    try:
        from astropy.wcs import WCS
        clean = {k: v for k, v in hdr.items() if isinstance(v, (int, float, str, bool))}
        return dict(WCS(clean, naxis=2).to_header())
    except ImportError:
        pass
    out = {'WCSAXES': 2}
    have_cd = any(k in hdr for k in ('CD1_1', 'CD1_2', 'CD2_1', 'CD2_2'))
    for key in _WCS_KEYS[1:]:
        is_cd_matrix = key.startswith('CD') and not key.startswith('CDELT')
        if key in hdr and not is_cd_matrix:
            out[key] = hdr[key]
    if have_cd:
        for i in (1, 2):
            for j in (1, 2):
                out['PC%d_%d' % (i, j)] = float(hdr.get('CD%d_%d' % (i, j), 0.0))
            out['CDELT%d' % i] = 1.0
    return out


def interferometer_angles(file, hdr):
    """``LUCI/LuciUtility.py:120-140``: theta[x, y] = arccos(lambda_ref / calib_map) in degrees (float64)."""
    calib_map = np.asarray(file['calib_map'][()]).astype('float')
    calib_ref = np.float32(hdr['CALIBNM'] if 'CALIBNM' in hdr else hdr['nm_laser'])
    return np.rad2deg(np.arccos(calib_ref / calib_map))


def read_in_transmission(Luci_path, hdr, spectrum_axis_unshifted):
    """``LUCI/LuciUtility.py:306-316``: ``<Luci_path>/Data/<FILTER>_filter.dat`` (cm^-1, percent), linearly
    interpolated -- and linearly extrapolated, ``interp1d(kind='slinear', fill_value='extrapolate')`` -- onto the
    unshifted spectral axis, as a fraction."""
    table = np.loadtxt('%s/Data/%s_filter.dat' % (Luci_path, hdr['FILTER']))
    x, y = table[:, 0].astype(np.float64), table[:, 1].astype(np.float64) / 100.0
    order = np.argsort(x, kind='stable')
    x, y = x[order], y[order]
    q = np.asarray(spectrum_axis_unshifted, dtype=np.float64)
    seg = np.clip(np.searchsorted(x, q, side='right') - 1, 0, len(x) - 2)
    slope = (y[seg + 1] - y[seg]) / (x[seg + 1] - x[seg])
    return y[seg] + slope * (q - x[seg])
