"""polardecoding_b200 -- B200-native (sm_100a) batched polar-code decoding engine.

The product is the C-ABI library ``libpolargpu.so`` (include/polargpu.h, sources in csrc/) and the
drop-in C host programs in host/.  This Python package is only the ctypes binding that tests and
bench.py use to call that C ABI; it contains no decoding logic and no CPU fallback."""
from .capi import Engine, PgParams, PgCounters, load_library, LIB_PATH, PROGRAMS  # noqa: F401
