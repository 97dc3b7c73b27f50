"""luci_b200 -- B200-native implementation of LUCI's per-spaxel emission-line fitting hot path.

Compute lives in ``liblucifit.so`` (hand-written CUDA for sm_100a behind the C ABI of ``include/lucifit.h``);
this package is the host glue that keeps LUCI's Python API.  There is no CPU fallback.
"""
__version__ = '0.1.0'
