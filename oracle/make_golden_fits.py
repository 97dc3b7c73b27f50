"""Golden header cards of the reference's own example outputs (Data/ExampleData/Luci_outputs/**/*.fits, written by
LuciUtility.save_fits with `Cutout2D(...).wcs.to_header()`, LuciBase.py:511-513) -> tests/golden/fits_example_cards.json.
Test infrastructure; run in the build container where /root/reference exists:   python -m oracle.make_golden_fits"""
import glob
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from luci_b200.fitsio import read_cards, read_fits                       # noqa: E402

REF = '/root/reference/Data/ExampleData/Luci_outputs'


def main():
    out = {}
    for path in sorted(glob.glob(os.path.join(REF, '**', '*.fits'), recursive=True)):
        rel = os.path.relpath(path, REF)
        data, _ = read_fits(path)
        out[rel] = {'cards': [c.rstrip() for c in read_cards(path)], 'shape': list(data.shape), 'dtype': str(data.dtype),
                    'file_sha1': hashlib.sha1(open(path, 'rb').read()).hexdigest(),
                    'data_sha1': hashlib.sha1(data.astype('>f4' if data.dtype == 'float32' else '>f8').tobytes()).hexdigest()}
    with open(os.path.join(ROOT, 'tests', 'golden', 'fits_example_cards.json'), 'w') as fh:
        json.dump(out, fh, indent=0, sort_keys=True)
    print(len(out), 'files')


if __name__ == '__main__':
    main()
