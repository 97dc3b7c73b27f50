"""FITS / WCS layout of the saved maps against the reference's own example outputs
(Data/ExampleData/Luci_outputs/**, written by LuciUtility.save_fits with `Cutout2D(...).wcs.to_header()`,
LuciBase.py:511-513; LuciUtility.py:69-85): file names, card order, values, comments, BITPIX, axes."""
import glob
import hashlib
import json
import os

import numpy as np
import pytest

from luci_b200.fitsio import cutout_origin, read_cards, read_fits, save_fits, wcs_to_header, write_fits

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'fits_example_cards.json')
REF = '/root/reference/Data/ExampleData/Luci_outputs'


def test_header_cards_equal_the_reference_example_card_for_card(tmp_path):
    """Every example header, parsed to a plain {key: value} dict (what Luci.header holds) and written back through
    wcs_to_header + write_fits, gives the reference's 80-character cards exactly (order, values, WCSLIB comments)."""
    gold = json.load(open(GOLD))
    assert len(gold) >= 90
    for rel, g in gold.items():
        plain = {}
        for card in g['cards']:
            key = card[:8].strip()
            if card[8:10] == '= ' and key not in ('SIMPLE', 'BITPIX', 'NAXIS', 'NAXIS1', 'NAXIS2'):
                v = card[10:].split('/')[0].strip()
                plain[key] = v.strip("'").strip() if v.startswith("'") else (float(v) if ('.' in v or 'E' in v) else int(v))
        data = np.zeros(g['shape'], dtype=g['dtype'])
        out = str(tmp_path / 'x.fits')
        write_fits(out, data, wcs_to_header(plain))
        assert [c.rstrip() for c in read_cards(out)] == g['cards'], rel


@pytest.mark.reference
def test_example_files_round_trip_byte_for_byte(tmp_path):
    files = sorted(glob.glob(os.path.join(REF, '**', '*.fits'), recursive=True))
    assert len(files) >= 90
    gold = json.load(open(GOLD))
    for path in files:
        data, hdr = read_fits(path)
        out = str(tmp_path / 'x.fits')
        write_fits(out, data, wcs_to_header(dict(hdr)))
        raw = open(out, 'rb').read()
        assert raw == open(path, 'rb').read(), path
        assert hashlib.sha1(raw).hexdigest() == gold[os.path.relpath(path, REF)]['file_sha1']


def test_cutout_origin_is_cutout2d_arithmetic():
    """Data/ExampleData/Luci_outputs/Velocity/NGC628_1_sincgauss_Halpha_velocity.fits is 150 x 200 with CRPIX = (-1025, -925)
    on a header without celestial keywords: fit_cube(..., 1050, 1200, 900, 1100).  Cutout2D reads size = (dx, dy) as
    (ny, nx), so the origin is (ceil(1125 - 100), ceil(1000 - 75)), not (1050, 900)."""
    assert cutout_origin(1050, 1200, 900, 1100, 2048, 2064) == (1025.0, 925.0)
    assert cutout_origin(1050, 1200, 900, 1100, 2048, 2064, strict_reference=False) == (1050.0, 900.0)
    assert cutout_origin(2, 9, 1, 8, 12, 10) == (2.0, 1.0)                       # square regions: the rectangle itself
    assert cutout_origin(0, 2048, 0, 2064, 2048, 2064) == (0.0, 8.0)             # full detector: x trimmed to the image, y displaced
    g = json.load(open(GOLD))['Velocity/NGC628_1_sincgauss_Halpha_velocity.fits']
    h = wcs_to_header({}, cutout_origin(1050, 1200, 900, 1100, 2048, 2064))
    assert h['CRPIX1'][0] == -1025.0 and h['CRPIX2'][0] == -925.0
    assert g['shape'] == [200, 150]
    cards = {c[:8].strip(): c for c in g['cards']}
    assert float(cards['CRPIX1'][10:30]) == -1025.0 and float(cards['CRPIX2'][10:30]) == -925.0


def test_save_fits_names_match_the_example_tree(tmp_path):
    """The files save_fits writes for the example's call (5 SN3 lines, sincgauss, binning 1) carry the names of the
    reference's example tree (LuciUtility.py:58-85)."""
    lines = ['Halpha', 'NII6548', 'NII6583', 'SII6716', 'SII6731']
    m3 = np.zeros((4, 3, 5), np.float32)
    m2 = np.zeros((4, 3), np.float32)
    save_fits(str(tmp_path), 'NGC628', lines, m3, m3, m3, m3, m3, m3, m3, m2, m2, m2, wcs_to_header({}), 1, fit_function='sincgauss')
    mine = {os.path.relpath(p, str(tmp_path)) for p in glob.glob(str(tmp_path / '**' / '*.fits'), recursive=True)}
    gold = {k for k in json.load(open(GOLD)) if k.split('/')[-1].startswith('NGC628_1_sincgauss_')}
    # the example tree predates the *_velocity_err / *_broadening_err / *_Amplitude maps of the current save_fits
    assert gold <= mine, sorted(gold - mine)
    d, h = read_fits(str(tmp_path / 'Velocity' / 'NGC628_1_sincgauss_Halpha_velocity.fits'))
    assert d.shape == (4, 3) and d.dtype == np.float32 and h['NAXIS1'] == 3 and h['WCSAXES'] == 2
