import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')
    config.addinivalue_line('markers', 'reference: needs the read-only reference tree at /root/reference')


def pytest_collection_modifyitems(config, items):
    from oracle import ref_shim
    have_ref = ref_shim.reference_available()
    skip_ref = pytest.mark.skip(reason='reference tree not present (GPU box)')
    for item in items:
        if 'reference' in item.keywords and not have_ref:
            item.add_marker(skip_ref)


@pytest.fixture(scope='session')
def hostemu():
    """The device fit pipeline compiled for the CPU (LANES = 1); test infrastructure only."""
    import ctypes as C
    from luci_b200._capi import lucifit_config
    here = os.path.join(ROOT, 'tests', 'hostemu')
    so = os.path.join(here, '_build', 'libhostemu.so')
    srcs = [os.path.join(here, 'hostemu.cpp')] + [os.path.join(ROOT, 'luci_b200', 'csrc', f) for f in
                                                    ('lucifit_core.cuh', 'lucifit_math.cuh', 'lucifit_host.hpp',
                                                     'lucifit_instances.h', 'lucifit_prior.cuh')]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-o', so, srcs[0]])
    lib = C.CDLL(so)
    vp = C.c_void_p
    lib.hostemu_fit_batch.argtypes = [C.POINTER(lucifit_config), C.c_int64] + [vp] * 10
    lib.hostemu_fit_batch.restype = C.c_int
    lib.hostemu_profile.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, vp, vp, vp, vp]
    lib.hostemu_faddeeva.argtypes = [C.c_int, vp, vp, vp, vp]
    lib.hostemu_eval.argtypes = [C.POINTER(lucifit_config), vp, C.c_float, vp, vp, vp]
    lib.hostemu_eval.restype = C.c_int
    lib.hostemu_estimate_prior.argtypes = [C.POINTER(lucifit_config), C.c_int64, vp, vp, vp, C.c_double, C.c_double, C.c_double, vp, vp]
    lib.hostemu_estimate_prior.restype = C.c_int
    return lib
