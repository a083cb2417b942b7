"""Development aid: A/B the list kernel across library variants (tools/variant.sh).
  python tools/ab.py build/variants/libpolargpu_v1.so ...     ('base' = the in-tree library)
Each variant runs in its own process: fp64 parity against the oracle on a few frames, then device-side timing.
(Batches are at most one chunk of the library: last_kernel_ms() times the LAST launch only -- the round-1 "SC 198 M frames/s" came from a two-chunk batch.)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
from polardecoding_b200 import Engine
from oracle_lib import Oracle
ok = True
for prog, B, snr in (("CASCL_1024_L8", 48, 1.5), ("CASCL_128", 256, 1.5), ("SC_1024", 64, 2.0)):
    e = Engine(prog, real="f64", seed=11, data_mode=1)
    llr, u = e.channel(snr, 0, B)
    got, fl = e.decode_llr(llr)
    want, aux = Oracle(prog).decode(llr)
    bad = int((got != want).any(1).sum())
    ok &= bad == 0
    print("  parity %%-14s f64: %%d of %%d frames differ" %% (prog, bad, B))
    e.close()
for prog, real, snr, mult in (("CASCL_1024_L8", "f32", 2.0, 6), ("CASCL_1024_L8", "f64", 2.0, 1), ("SC_1024", "f32", 2.0, 1)):
    e = Engine(prog, real=real)
    B = int(e.wave_frames()) * mult
    e.simulate_batch(snr, 0, B)
    best = 1e9
    for rep in range(3):
        acc, _ = e.simulate_batch(snr, 1 << 22, B)
        best = min(best, e.last_kernel_ms()[0])
    print("  %%-14s %%s B=%%d decode %%.3f ms -> %%.3f Mframes/s  FER %%.4g ties %%d" %% (prog, real, B, best, B / best / 1e3, acc.err_blocks / acc.frames, acc.tie_frames))
    e.close()
print("  PARITY", "OK" if ok else "FAILED")
'''
for lib in sys.argv[1:]:
    env = dict(os.environ)
    if lib != "base":
        env["POLARGPU_LIB"] = os.path.abspath(lib)
    print("==", lib, flush=True)
    subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT}], env=env)
