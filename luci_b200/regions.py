"""ds9 region files -> boolean spaxel masks, the input side of ``fit_region`` / ``fit_spectrum_region`` /
``extract_spectrum_region`` (``LuciBase.py:622-624, 917-919, 985-987`` -> ``LuciConvenience.reg_to_mask :49-57``).

The reference does ``pyregion.open(path).as_imagecoord(header).get_mask(shape=(2064, 2048)).T``.  pyregion (third party,
``requirements.txt``; not vendored in the reference and absent from this image) is restated here from its documented
behaviour -- *parity unpinned against pyregion itself*, pinned on analytic cases in ``tests/test_regions.py``:

  * the DS9 4.x text format: ``#`` comments, ``global`` line, a coordinate-system line (``image`` / ``physical`` /
    ``fk5`` / ``icrs`` / ``j2000``) or a ``system;shape(...)`` prefix, ``-shape`` exclusions, trailing ``# properties``;
    shapes with an area: ``circle``, ``ellipse``, ``box``, ``polygon``, ``annulus`` (others -- point, line, text ... --
    carry no area and are skipped, like pyregion's ``get_mask``);
  * ``as_imagecoord``: sky positions go through the header's celestial WCS (zenithal TAN, the projection of SITELLE
    cubes; core WCS only, no SIP -- pyregion calls ``wcs_world2pix`` as well), angular sizes are divided by the local
    pixel scale, angles are carried over by the local direction of north / east; ``physical`` equals ``image`` unless the
    header holds ``LTV`` / ``LTM`` cards;
  * ``get_mask``: 1-based DS9 pixel coordinates shifted to 0-based, a pixel belongs to a shape if its CENTRE is inside;
    an exclusion removes its area from everything listed before it.

The mask comes back as ``mask[x, y]`` (the ``.T`` of the reference), ready for ``Luci._mask_rows``.
"""
import math
import re

import numpy as np

_SKY_SYSTEMS = {'fk5', 'icrs', 'j2000'}
_PIX_SYSTEMS = {'image', 'physical'}
_OTHER_SYSTEMS = {'fk4', 'b1950', 'galactic', 'ecliptic', 'wcs', 'wcsa', 'linear', 'amplifier', 'detector'}
_AREA_SHAPES = {'circle', 'ellipse', 'box', 'polygon', 'annulus'}
_SKIPPED_SHAPES = {'point', 'line', 'vector', 'text', 'ruler', 'compass', 'projection', 'segment',
                   'circle point', 'box point', 'diamond point', 'cross point', 'x point', 'arrow point', 'boxcircle point'}
_SHAPE_RE = re.compile(r'^\s*([+-]?)\s*([a-zA-Z ]+?)\s*\((.*?)\)?\s*$')


class Shape:
    """One region shape: ``name``, raw argument strings, the coordinate system it was written in and its sign."""

    def __init__(self, name, args, system, exclude):
        self.name, self.args, self.system, self.exclude = name, args, system, exclude

    def __repr__(self):
        return '%s%s(%s) [%s]' % ('-' if self.exclude else '', self.name, ','.join(self.args), self.system)


# ------------------------------------------------------------------------------------------------------ parsing
def parse_ds9(text):
    """List of :class:`Shape` in file order.  Unknown coordinate systems raise; shapes without an area are dropped."""
    system = 'image'                                   # DS9's default when a file names none is physical == image here
    shapes = []
    for raw in text.replace('\r', '').split('\n'):
        line = raw.split('#', 1)[0].strip()
        if not line:
            continue
        for part in line.split(';'):
            part = part.strip()
            if not part or part.lower().startswith('global'):
                continue
            low = part.lower()
            if low in _SKY_SYSTEMS or low in _PIX_SYSTEMS:
                system = low
                continue
            if low in _OTHER_SYSTEMS:
                raise ValueError('ds9 coordinate system %r is not supported (image, physical, fk5, icrs, j2000 are)' % part)
            m = _SHAPE_RE.match(part)
            if not m:
                raise ValueError('cannot read ds9 region line %r' % raw)
            name = ' '.join(m.group(2).lower().split())
            if name in _SKIPPED_SHAPES:
                continue
            if name not in _AREA_SHAPES:
                raise ValueError('ds9 shape %r is not supported' % name)
            args = [a for a in re.split(r'[,\s]+', m.group(3).strip()) if a]
            shapes.append(Shape(name, args, system, m.group(1) == '-'))
    return shapes


def _sexagesimal(tok):
    """``dd:mm:ss.s`` / ``hh:mm:ss.s`` / ``12h30m5s`` / ``12d30m5s`` -> (value in the leading unit, had_hour_marker)."""
    t = tok.strip()
    hours = 'h' in t.lower()
    t = re.sub(r'[hdm:]', ' ', t.lower()).replace('s', ' ')
    parts = t.split()
    sign = -1.0 if parts[0].startswith('-') else 1.0
    vals = [abs(float(p)) for p in parts] + [0.0, 0.0]
    return sign * (vals[0] + vals[1] / 60.0 + vals[2] / 3600.0), hours


def parse_position(tok_lon, tok_lat):
    """Sky position of a ``fk5`` / ``icrs`` shape in degrees.  A colon-separated longitude is in hours (DS9's convention for
    equatorial systems), ``h`` / ``d`` markers say so explicitly, a bare number is degrees."""
    lon_t = tok_lon.strip().lower()
    if ':' in lon_t or 'h' in lon_t:
        v, _ = _sexagesimal(lon_t)
        lon = v * 15.0
    elif lon_t.endswith('d') or 'm' in lon_t:
        lon, _ = _sexagesimal(lon_t)
    elif lon_t.endswith('r'):
        lon = math.degrees(float(lon_t[:-1]))
    else:
        lon = float(lon_t)
    lat_t = tok_lat.strip().lower()
    if ':' in lat_t or 'd' in lat_t or 'm' in lat_t:
        lat, _ = _sexagesimal(lat_t)
    elif lat_t.endswith('r'):
        lat = math.degrees(float(lat_t[:-1]))
    else:
        lat = float(lat_t)
    return lon, lat


def parse_size(tok, sky):
    """A length: ``(value, unit)`` with unit ``'deg'`` or ``'pix'``.  ``"`` arcsec, ``'`` arcmin, ``d`` degrees, ``r``
    radians, ``p`` / ``i`` pixels; a bare number is degrees in a sky system and pixels otherwise."""
    t = tok.strip()
    if t.endswith('"'):
        return float(t[:-1]) / 3600.0, 'deg'
    if t.endswith("'"):
        return float(t[:-1]) / 60.0, 'deg'
    if t.endswith('d'):
        return float(t[:-1]), 'deg'
    if t.endswith('r'):
        return math.degrees(float(t[:-1])), 'deg'
    if t.endswith('p') or t.endswith('i'):
        return float(t[:-1]), 'pix'
    return float(t), ('deg' if sky else 'pix')


# ---------------------------------------------------------------------------------------------------------- WCS
class TanWCS:
    """Core celestial WCS of a 2-D header (dict of cards, values or ``(value, comment)``): ``CRPIX``, ``CRVAL``,
    ``CDi_j`` or ``PCi_j`` x ``CDELT``, ``LONPOLE``; zenithal gnomonic projection (``RA---TAN`` / ``DEC--TAN``).
    Pixel coordinates are FITS 1-based.  Formulas: Calabretta & Greisen 2002, eqs. 2, 5, 54, 55."""

    def __init__(self, header):
        h = {str(k).upper(): (v[0] if isinstance(v, tuple) else v) for k, v in (header or {}).items()}
        ct1, ct2 = str(h.get('CTYPE1', '')), str(h.get('CTYPE2', ''))
        if not (ct1.startswith('RA--') and ct2.startswith('DEC-')) or ct1[5:8] != 'TAN' or ct2[5:8] != 'TAN':
            raise ValueError('sky-coordinate regions need a RA---TAN / DEC--TAN header (CTYPE1 = %r, CTYPE2 = %r)' % (ct1, ct2))
        self.crpix = np.array([float(h.get('CRPIX1', 0.0)), float(h.get('CRPIX2', 0.0))])
        self.crval = np.array([float(h.get('CRVAL1', 0.0)), float(h.get('CRVAL2', 0.0))])
        if any(('CD%d_%d' % (i, j)) in h for i in (1, 2) for j in (1, 2)):
            m = [[float(h.get('CD%d_%d' % (i, j), 0.0)) for j in (1, 2)] for i in (1, 2)]
        else:
            m = [[float(h.get('PC%d_%d' % (i, j), 1.0 if i == j else 0.0)) * float(h.get('CDELT%d' % i, 1.0))
                  for j in (1, 2)] for i in (1, 2)]
        self.m = np.array(m)
        self.minv = np.linalg.inv(self.m)
        self.lonpole = math.radians(float(h.get('LONPOLE', 180.0 if self.crval[1] < 90.0 else 0.0)))

    def world2pix(self, lon, lat):
        a, d = np.radians(np.asarray(lon, float)) - math.radians(self.crval[0]), np.radians(np.asarray(lat, float))
        d0 = math.radians(self.crval[1])
        phi = self.lonpole + np.arctan2(-np.cos(d) * np.sin(a), np.sin(d) * math.cos(d0) - np.cos(d) * math.sin(d0) * np.cos(a))
        sin_theta = np.sin(d) * math.sin(d0) + np.cos(d) * math.cos(d0) * np.cos(a)
        r = np.degrees(np.sqrt(np.maximum(1.0 - sin_theta * sin_theta, 0.0)) / sin_theta)       # (180/pi) cot(theta)
        x, y = r * np.sin(phi), -r * np.cos(phi)
        px = self.crpix[0] + self.minv[0, 0] * x + self.minv[0, 1] * y
        py = self.crpix[1] + self.minv[1, 0] * x + self.minv[1, 1] * y
        return px, py

    def pix2world(self, px, py):
        u, v = np.asarray(px, float) - self.crpix[0], np.asarray(py, float) - self.crpix[1]
        x = self.m[0, 0] * u + self.m[0, 1] * v
        y = self.m[1, 0] * u + self.m[1, 1] * v
        phi = np.arctan2(x, -y)
        theta = np.arctan2(180.0 / math.pi, np.hypot(x, y))
        d0 = math.radians(self.crval[1])
        dphi = phi - self.lonpole
        lat = np.arcsin(np.sin(theta) * math.sin(d0) + np.cos(theta) * math.cos(d0) * np.cos(dphi))
        lon = math.radians(self.crval[0]) + np.arctan2(-np.cos(theta) * np.sin(dphi),
                                                       np.sin(theta) * math.cos(d0) - np.cos(theta) * math.sin(d0) * np.cos(dphi))
        return np.degrees(lon) % 360.0, np.degrees(lat)

    def local_frame(self, lon, lat, step=1.0 / 3600.0):
        """Pixel-space unit-sky vectors at a sky position: ``(d pix / d east, d pix / d north)`` per degree, and the
        pixel scale ``sqrt(|det|)`` in degrees per pixel (what pyregion's ``estimate_cdelt`` measures)."""
        x0, y0 = self.world2pix(lon, lat)
        xe, ye = self.world2pix(lon + step / max(math.cos(math.radians(lat)), 1e-12), lat)
        xn, yn = self.world2pix(lon, lat + step)
        east = np.array([xe - x0, ye - y0]) / step
        north = np.array([xn - x0, yn - y0]) / step
        det = abs(east[0] * north[1] - east[1] * north[0])
        return east, north, 1.0 / math.sqrt(det)


def _physical_to_image(h):
    """``image = LTM * physical + LTV`` when the header carries the IRAF cards (pyregion's physical -> image step)."""
    hh = {str(k).upper(): (v[0] if isinstance(v, tuple) else v) for k, v in (h or {}).items()}
    ltv = (float(hh.get('LTV1', 0.0)), float(hh.get('LTV2', 0.0)))
    ltm = (float(hh.get('LTM1_1', 1.0)), float(hh.get('LTM2_2', 1.0)))
    return ltm, ltv


# ------------------------------------------------------------------------------------------- image coordinates
def as_imagecoord(shapes, header):
    """Shapes with numeric parameters in 1-based image pixels and angles in degrees from the image +x axis:
    list of ``(name, params, exclude)``; ``params`` = ``[xc, yc, sizes..., angle]`` or the polygon's vertex list."""
    out = []
    wcs = None
    for sh in shapes:
        sky = sh.system in _SKY_SYSTEMS
        if sky and wcs is None:
            wcs = TanWCS(header)
        ltm, ltv = ((1.0, 1.0), (0.0, 0.0)) if sh.system != 'physical' else _physical_to_image(header)

        def position(i):
            if sky:
                lon, lat = parse_position(sh.args[i], sh.args[i + 1])
                px, py = wcs.world2pix(lon, lat)
                return float(px), float(py), (lon, lat)
            return float(sh.args[i]) * ltm[0] + ltv[0], float(sh.args[i + 1]) * ltm[1] + ltv[1], None

        if sh.name == 'polygon':
            if len(sh.args) < 6 or len(sh.args) % 2:
                raise ValueError('polygon needs an even number (>= 6) of coordinates: %r' % sh)
            pts = []
            for i in range(0, len(sh.args), 2):
                px, py, _ = position(i)
                pts += [px, py]
            out.append((sh.name, pts, sh.exclude))
            continue
        xc, yc, world = position(0)
        rest = sh.args[2:]
        n_sizes = {'circle': 1, 'ellipse': 2, 'box': 2, 'annulus': len(rest)}[sh.name]
        if sh.name == 'ellipse' and len(rest) > 3:
            raise ValueError('elliptical annuli are not supported: %r' % sh)
        if len(rest) < n_sizes:
            raise ValueError('too few parameters: %r' % sh)
        scale, east, north = 1.0, None, None
        if world is not None:
            east, north, scale = wcs.local_frame(*world)
        sizes = []
        for tok in rest[:n_sizes]:
            v, unit = parse_size(tok, sky)
            if unit == 'deg':
                if world is None:
                    raise ValueError('angular size in a pixel-coordinate region needs a sky system: %r' % sh)
                v = v / scale
            elif sh.system == 'physical':
                v = v * ltm[0]
            sizes.append(v)
        angle = 0.0
        if len(rest) > n_sizes and sh.name in ('ellipse', 'box'):
            angle = float(rest[n_sizes].rstrip('d'))
            if world is not None:
                # a sky angle turns from west towards north; carry that direction into the pixel frame
                a = math.radians(angle)
                d = math.cos(a) * (-east) + math.sin(a) * north
                angle = math.degrees(math.atan2(d[1], d[0]))
        out.append((sh.name, [xc, yc] + sizes + [angle], sh.exclude))
    return out


# ------------------------------------------------------------------------------------------------ rasterisation
def _inside(name, p, X, Y):
    """Boolean array: pixel centres (0-based ``X``, ``Y``) inside one image-coordinate shape (1-based parameters)."""
    if name == 'polygon':
        xs = np.asarray(p[0::2], float) - 1.0
        ys = np.asarray(p[1::2], float) - 1.0
        inside = np.zeros(X.shape, bool)
        n = len(xs)
        for i in range(n):                                  # crossing number, edges half-open in y
            x1, y1, x2, y2 = xs[i], ys[i], xs[(i + 1) % n], ys[(i + 1) % n]
            if y1 == y2:
                continue
            cond = (y1 <= Y) != (y2 <= Y)
            xint = x1 + (Y - y1) * (x2 - x1) / (y2 - y1)
            inside ^= cond & (X < xint)
        return inside
    xc, yc = p[0] - 1.0, p[1] - 1.0
    dx, dy = X - xc, Y - yc
    if name == 'circle':
        return dx * dx + dy * dy <= p[2] * p[2]
    if name == 'annulus':
        r2 = dx * dx + dy * dy
        radii = p[2:-1]
        return (r2 <= radii[-1] * radii[-1]) & (r2 > radii[0] * radii[0])
    a = math.radians(p[-1])
    u = dx * math.cos(a) + dy * math.sin(a)
    v = -dx * math.sin(a) + dy * math.cos(a)
    if name == 'ellipse':
        return (u / p[2]) ** 2 + (v / p[3]) ** 2 <= 1.0
    if name == 'box':
        return (np.abs(u) <= 0.5 * p[2]) & (np.abs(v) <= 0.5 * p[3])
    raise ValueError(name)


def _bbox(name, p, nx, ny):
    if name == 'polygon':
        xs, ys = np.asarray(p[0::2]) - 1.0, np.asarray(p[1::2]) - 1.0
        x0, x1, y0, y1 = xs.min(), xs.max(), ys.min(), ys.max()
    else:
        r = {'circle': lambda: p[2], 'annulus': lambda: max(p[2:-1]), 'ellipse': lambda: max(p[2], p[3]),
             'box': lambda: 0.5 * math.hypot(p[2], p[3])}[name]()
        x0, x1, y0, y1 = p[0] - 1 - r, p[0] - 1 + r, p[1] - 1 - r, p[1] - 1 + r
    lo = lambda v: int(max(0, math.floor(v)))
    return lo(x0), int(min(nx, math.ceil(x1) + 1)), lo(y0), int(min(ny, math.ceil(y1) + 1))


def get_mask(image_shapes, nx, ny):
    """``mask[x, y]`` (bool, ``[nx, ny]``) of a list from :func:`as_imagecoord`.  Shapes are applied in file order: an
    included shape adds its pixels, an excluded one removes its pixels from what has been collected so far."""
    mask = np.zeros((nx, ny), bool)
    for name, p, exclude in image_shapes:
        x0, x1, y0, y1 = _bbox(name, p, nx, ny)
        if x1 <= x0 or y1 <= y0:
            continue
        X, Y = np.meshgrid(np.arange(x0, x1, dtype=np.float64), np.arange(y0, y1, dtype=np.float64), indexing='ij')
        ins = _inside(name, p, X, Y)
        if exclude:
            mask[x0:x1, y0:y1] &= ~ins
        else:
            mask[x0:x1, y0:y1] |= ins
    return mask


def reg_to_mask(region, header, shape):
    """``LuciConvenience.reg_to_mask`` for a cube of ``shape = (nx, ny[, nz])``: the reference hard-codes SITELLE's
    2048 x 2064 detector, here the mask has the size of the cube it will index."""
    with open(region) as fh:
        shapes = parse_ds9(fh.read())
    if not shapes:
        raise ValueError('no region with an area in %s' % region)
    return get_mask(as_imagecoord(shapes, header), int(shape[0]), int(shape[1]))
