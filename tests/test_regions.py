"""ds9 region files -> masks (luci_b200.regions; LuciConvenience.reg_to_mask :49-57).  pyregion is not in this image, so the
restatement is pinned on analytic cases: pixel-centre rasterisation with DS9's 1-based coordinates, the gnomonic
projection written independently of the module's native-spherical form, angular sizes through the pixel scale."""
import math

import numpy as np
import pytest

from luci_b200 import regions as R

HEAD = ('# Region file format: DS9 version 4.1\n'
        'global color=green dashlist=8 3 width=1 font="helvetica 10 normal roman" select=1 highlite=1 dash=0 fixed=0 '
        'edit=1 move=1 delete=1 include=1 source=1\n')


def _mask(text, nx, ny, header=None):
    return R.get_mask(R.as_imagecoord(R.parse_ds9(text), header), nx, ny)


def test_circle_physical_pixel_centres():
    # DS9 pixel (3, 5) is array element [2, 4]; radius 0.5 holds that one centre only
    m = _mask(HEAD + 'physical\ncircle(3,5,0.5)\n', 10, 12)
    assert m.shape == (10, 12) and m.sum() == 1 and m[2, 4]
    # the shape of the reference's example files (Examples/regions/bkg_M33.reg: a circle in physical coordinates)
    m = _mask(HEAD + 'physical\ncircle(248.64416,1618.2529,59.719676)\n', 2048, 2064)
    xs, ys = np.nonzero(m)
    assert abs(m.sum() / (math.pi * 59.719676 ** 2) - 1) < 2e-3
    assert abs(xs.mean() - 247.64416) < 0.05 and abs(ys.mean() - 1617.2529) < 0.05
    brute = (np.arange(2048)[:, None] - 247.64416) ** 2 + (np.arange(2064)[None, :] - 1617.2529) ** 2 <= 59.719676 ** 2
    assert (m == brute).all()


def test_clipped_at_the_edge_and_outside():
    m = _mask('image\ncircle(1,1,3)\n', 8, 8)
    assert m[0, 0] and m[3, 0] and not m[4, 0] and m.sum() == 11
    assert _mask('image\ncircle(100,100,3)\n', 8, 8).sum() == 0


def test_box_ellipse_polygon_annulus_and_exclusion():
    m = _mask('image\nbox(10,20,4,6,0)\n', 30, 40)
    assert m.sum() == 5 * 7 and m[7:12, 16:23].all()
    m90 = _mask('image\nbox(10,20,4,6,90)\n', 30, 40)
    assert m90.sum() == 7 * 5 and m90[6:13, 17:22].all()
    e = _mask('image\nellipse(50,50,20,10,30)\n', 100, 100)
    assert abs(e.sum() / (math.pi * 200) - 1) < 0.02
    a = math.radians(30)
    assert e[49 + int(18 * math.cos(a)), 49 + int(18 * math.sin(a))] and not e[49 - int(18 * math.sin(a)), 49 + int(18 * math.cos(a))]
    p = _mask('image\npolygon(2,2,12,2,12,7,2,7)\n', 20, 20)
    assert p.sum() == 10 * 5 and p[1:11, 1:6].all()              # half-open edges: a 10 x 5 rectangle
    an = _mask('image\nannulus(50,50,10,20)\n', 100, 100)
    assert abs(an.sum() / (math.pi * 300) - 1) < 0.02 and not an[49, 49] and an[49 + 15, 49]
    ex = _mask('image\ncircle(50,50,20)\n-circle(50,50,10)\n', 100, 100)
    assert (ex == an).all()
    # an exclusion only removes what was listed before it
    later = _mask('image\n-circle(50,50,10)\ncircle(50,50,20)\n', 100, 100)
    assert later.sum() > ex.sum() and later[49, 49]


def test_parser_variants():
    s = R.parse_ds9(HEAD + 'fk5\ncircle(1:32:45.7313,+30:38:59.006,30.754") # color=red\npoint(1:32:45,+30:38:59) # point=x\n'
                           'image;box(5,5,2,2,0)\n')
    assert [(x.name, x.system, x.exclude) for x in s] == [('circle', 'fk5', False), ('box', 'image', False)]
    lon, lat = R.parse_position('1:32:45.7313', '+30:38:59.006')
    assert abs(lon - 15 * (1 + 32 / 60 + 45.7313 / 3600)) < 1e-12 and abs(lat - (30 + 38 / 60 + 59.006 / 3600)) < 1e-12
    assert R.parse_position('23.5', '-0:30:00') == (23.5, -0.5)
    assert R.parse_position('1h30m0s', '-10d30m0s') == (22.5, -10.5)
    assert R.parse_size('30"', True) == (30 / 3600.0, 'deg') and R.parse_size("2'", True)[0] == 2 / 60.0
    assert R.parse_size('7', False) == (7.0, 'pix') and R.parse_size('0.01', True) == (0.01, 'deg')
    with pytest.raises(ValueError):
        R.parse_ds9('galactic\ncircle(10,10,1)\n')
    with pytest.raises(ValueError):
        R.as_imagecoord(R.parse_ds9('fk5\ncircle(10,10,1")\n'), {'CTYPE1': 'LINEAR', 'CTYPE2': 'LINEAR'})


def _gnomonic(ra, dec, ra0, dec0):
    """Standard coordinates (degrees) of the textbook tangent-plane projection, written without native angles."""
    a, d, a0, d0 = map(math.radians, (ra, dec, ra0, dec0))
    den = math.sin(d) * math.sin(d0) + math.cos(d) * math.cos(d0) * math.cos(a - a0)
    xi = math.cos(d) * math.sin(a - a0) / den
    eta = (math.sin(d) * math.cos(d0) - math.cos(d) * math.sin(d0) * math.cos(a - a0)) / den
    return math.degrees(xi), math.degrees(eta)


def test_tan_wcs_against_textbook_projection_and_roundtrip():
    s = 0.321 / 3600.0                                           # SITELLE's pixel scale
    rot = math.radians(12.0)
    hdr = {'CTYPE1': 'RA---TAN', 'CTYPE2': 'DEC--TAN', 'CRVAL1': 23.19, 'CRVAL2': 30.65, 'CRPIX1': 1024.0, 'CRPIX2': 1032.0,
           'CD1_1': -s * math.cos(rot), 'CD1_2': s * math.sin(rot), 'CD2_1': s * math.sin(rot), 'CD2_2': s * math.cos(rot)}
    w = R.TanWCS(hdr)
    cd = np.array([[hdr['CD1_1'], hdr['CD1_2']], [hdr['CD2_1'], hdr['CD2_2']]])
    rng = np.random.default_rng(5)
    for _ in range(20):
        ra, dec = 23.19 + rng.uniform(-0.1, 0.1), 30.65 + rng.uniform(-0.1, 0.1)
        xi, eta = _gnomonic(ra, dec, 23.19, 30.65)
        want = np.array([1024.0, 1032.0]) + np.linalg.solve(cd, np.array([xi, eta]))
        px, py = w.world2pix(ra, dec)
        assert abs(px - want[0]) < 1e-7 and abs(py - want[1]) < 1e-7
        lon, lat = w.pix2world(px, py)
        assert abs(lon - ra) < 1e-9 and abs(lat - dec) < 1e-9
    # PCi_j x CDELT form of the same matrix (what WCS.to_header writes, hdf5io.wcs_header)
    hdr2 = {k: v for k, v in hdr.items() if not k.startswith('CD')}
    hdr2.update({'CDELT1': 1.0, 'CDELT2': 1.0, 'PC1_1': (hdr['CD1_1'], 'c'), 'PC1_2': hdr['CD1_2'], 'PC2_1': hdr['CD2_1'], 'PC2_2': hdr['CD2_2']})
    px2, py2 = R.TanWCS(hdr2).world2pix(23.25, 30.7)
    px1, py1 = w.world2pix(23.25, 30.7)
    assert px1 == px2 and py1 == py2
    east, north, scale = w.local_frame(23.25, 30.7)
    assert abs(scale / s - 1) < 1e-5


def test_fk5_circle_lands_where_the_wcs_says(tmp_path):
    s = 0.321 / 3600.0
    hdr = {'CTYPE1': 'RA---TAN', 'CTYPE2': 'DEC--TAN', 'CRVAL1': 308.6061, 'CRVAL2': 60.1889, 'CRPIX1': 150.0, 'CRPIX2': 120.0,
           'PC1_1': -s, 'PC1_2': 0.0, 'PC2_1': 0.0, 'PC2_2': s, 'CDELT1': 1.0, 'CDELT2': 1.0}
    reg = tmp_path / 'reg1.reg'
    reg.write_text(HEAD + 'fk5\ncircle(20:34:25.4737,+60:11:20.006,18.923")\n')
    m = R.reg_to_mask(str(reg), hdr, (300, 260, 5))
    assert m.shape == (300, 260)
    ra, dec = 15 * (20 + 34 / 60 + 25.4737 / 3600), 60 + 11 / 60 + 20.006 / 3600
    xi, eta = _gnomonic(ra, dec, 308.6061, 60.1889)
    xc, yc, r = 150.0 + xi / -s - 1.0, 120.0 + eta / s - 1.0, 18.923 / 0.321
    brute = (np.arange(300)[:, None] - xc) ** 2 + (np.arange(260)[None, :] - yc) ** 2 <= r * r
    assert m.sum() > 10000 and (m ^ brute).sum() <= 2          # the local pixel scale differs from CDELT by ~1e-6
    # a box whose sky angle is 0 is aligned with east-west: on a cube rotated by 90 degrees its long side runs along y
    hdr_rot = dict(hdr, PC1_1=0.0, PC1_2=s, PC2_1=s, PC2_2=0.0)
    reg.write_text('fk5\nbox(20:34:25.4737,+60:11:20.006,20",4",0)\n')
    upright = R.reg_to_mask(str(reg), hdr, (300, 260))
    turned = R.reg_to_mask(str(reg), hdr_rot, (300, 260))
    ux, uy = np.nonzero(upright)
    tx, ty = np.nonzero(turned)
    assert np.ptp(ux) > 3 * np.ptp(uy) and np.ptp(ty) > 3 * np.ptp(tx)
