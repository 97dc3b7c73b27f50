"""TEST INFRASTRUCTURE ONLY -- import shims that let the unmodified reference run in this container.

The reference (``/root/reference``: ``LuciBase.py``, ``LUCI/LuciFit.py`` ...) imports packages that are
not installed here (astropy, h5py, emcee, dynesty, keras, tensorflow, matplotlib, ppxf, ...), decorates
its methods with ``numba.jit(fastmath=True)`` which numba >= 0.59 rejects (the pinned 0.58.1 silently
fell back to object mode, i.e. plain Python), and obtains its velocity/broadening prior from a Keras CNN
(``LuciFit.py:177-195, 335-362``).  This module

* installs permissive stub modules for the missing packages (nothing on the non-Bayesian fit path calls
  into them, except the two items below),
* turns ``numba.jit`` into the identity decorator (== object mode),
* gives the stubbed ``astropy.stats`` real ``sigma_clip`` / ``mad_std`` (from ``oracle.astro_stats``),
* lets a test inject the CNN prior via ``StubCNN`` (``install_prior``), which exercises the reference's own
  ML-prior branch (``interpolate_spectrum`` + ``estimate_priors_ML``, ``LuciFit.py:776-780``).

It only works where ``/root/reference`` exists (the build container).  It is used by
``oracle/make_golden.py`` to generate the committed fixtures under ``tests/golden/`` and by the
``reference``-marked tests.  The GPU box never imports it.
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("LUCI_REFERENCE_ROOT", "/root/reference")

_MISSING = {
    "h5py", "astropy", "emcee", "dynesty", "keras", "tensorflow", "tensorflow_probability",
    "matplotlib", "ppxf", "pyregion", "photutils", "corner", "seaborn", "skimage", "astroquery",
}


class _DummyMeta(type):
    def __getattr__(cls, name):            # e.g. ``fits.writeto`` where ``fits`` itself is the _Dummy class
        if name.startswith("__"):
            raise AttributeError(name)
        return _Dummy


class _Dummy(metaclass=_DummyMeta):
    """Swallows any construction / call / attribute access."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Dummy()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Dummy()


class _StubModule(types.ModuleType):
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Dummy


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path=None, target=None):
        if name.split(".")[0] in _MISSING:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        mod = _StubModule(spec.name)
        mod.__file__ = "/nonexistent/%s/__init__.py" % spec.name.replace(".", "/")  # LuciSim.py:10 reads it
        return mod

    def exec_module(self, module):
        pass


class StubCNN:
    """Stands in for the Keras model: returns a fixed ``[[velocity, broadening]]`` (km/s)."""

    def __init__(self, vel, broad):
        self.vel, self.broad = float(vel), float(broad)

    def __call__(self, spectrum, training=False):
        return [[self.vel, self.broad]]


_installed = False
_prior = {"fn": None}


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "LuciBase.py"))


def install():
    """Idempotently install the shims and put the reference on ``sys.path``."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    missing_now = set()
    for name in _MISSING:
        try:
            importlib.import_module(name)
        except Exception:
            missing_now.add(name)
    _MISSING.intersection_update(missing_now)
    sys.meta_path.insert(0, _Finder())
    import numba

    def _identity_jit(*a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]
        return lambda f: f

    numba.jit = _identity_jit
    if "astropy" in _MISSING:
        import astropy.stats as aps
        from oracle import astro_stats

        aps.sigma_clip = astro_stats.sigma_clip
        aps.mad_std = astro_stats.mad_std
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import LUCI.LuciFit as LF

    def _get_ml_model(self):
        fn = _prior["fn"]
        if self.ML_bool is True and fn is not None:
            v, b = fn(self)
            self.ML_model = StubCNN(v, b)

    LF.Fit.get_ML_model = _get_ml_model
    if "h5py" in _MISSING:
        import h5py

        class _NoDeepFrameFile:
            """Stands in for ``h5py.File`` inside ``Luci.create_deep_image`` (LuciBase.py:162-168): a file without a
            pre-computed ``deep_frame`` so that the reference sums the cube itself."""

            def __init__(self, *a, **k):
                pass

            def __contains__(self, key):
                return False

            def close(self):
                pass

        h5py.File = _NoDeepFrameFile
    _installed = True


def install_prior(fn):
    """``fn(fit_instance) -> (vel_km_s, broad_km_s)`` is what the stub CNN will answer for that spectrum.

    Pass ``None`` to remove the prior (the reference then behaves as with no ML model)."""
    _prior["fn"] = fn


def modules():
    """Return the reference's ``(LuciBase.Luci, LUCI.LuciFit)`` after installing the shims."""
    install()
    from LuciBase import Luci
    import LUCI.LuciFit as LF

    return Luci, LF
