"""Development aid: A/B the BP kernel across library variants: python tools/ab_bp.py lib.so ...  ('base' = in-tree)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
from polardecoding_b200 import Engine
from oracle_lib import Oracle
e = Engine("BP_1024", real="f64", seed=11, data_mode=1)
llr, u = e.channel(2.0, 0, 8)
got, fl = e.decode_llr(llr)
want, aux = Oracle("BP_1024").decode(llr)
print("  parity BP_1024 f64: %%d of 8 frames differ" %% int((got != want).any(1).sum()))
e.close()
for real, mult in (("f32", 16), ("h2", 16)):
    e = Engine("BP_1024", real=real)
    B = int(e.wave_frames()) * mult
    e.simulate_batch(2.5, 0, B)
    best = 1e9
    for rep in range(2):
        acc, _ = e.simulate_batch(2.5, 1 << 22, B)
        best = min(best, e.last_kernel_ms()[0])
    print("  BP_1024 %%s B=%%d decode %%.3f ms -> %%.4f Mframes/s  FER %%.4g" %% (real, B, best, B / best / 1e3, acc.err_blocks / acc.frames))
    e.close()
'''
for lib in sys.argv[1:]:
    env = dict(os.environ)
    if lib != "base":
        env["POLARGPU_LIB"] = os.path.abspath(lib)
    print("==", lib, flush=True)
    subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT}], env=env)
