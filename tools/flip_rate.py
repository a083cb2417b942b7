"""fp32 throughput mode against the bit-exact fp64 mode on the SAME float-representable LLRs (both on the GPU; the fp64
kernels are bit-exact with the reference, tests/test_gpu_parity.py): fraction of frames in which any decided bit differs,
and how many of those frames are flagged as exact path-metric ties.  python tools/flip_rate.py [--frames 200000]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from polardecoding_b200 import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=200000)
ap.add_argument("--out", default=None)
a = ap.parse_args()
rows = ["# fp32 vs fp64 decisions on identical LLRs (%d frames per point, Philox channel, random payload)" % a.frames, "",
        "| program | Eb/N0 dB | frames differing | rate | of which tie-flagged | FER fp64 | FER fp32 |", "|---|---|---|---|---|---|---|"]
for prog, snrs in (("SC_1024", (1.0, 2.0, 3.0)), ("SCL_1024", (1.0, 2.0, 3.0)), ("CASCL_1024_L8", (1.0, 1.5, 2.0, 2.5)), ("CASCL_128", (1.0, 2.0, 3.0)),
                   ("BP_1024", (2.0, 3.0))):
    e32 = Engine(prog, real="f32", seed=99, data_mode=1)
    e64 = Engine(prog, real="f64", seed=99, data_mode=1)
    for snr in snrs:
        B = a.frames if not prog.startswith("BP") else a.frames // 10
        diff = ties = 0
        err32 = err64 = 0
        for off in range(0, B, 20000):
            b = min(20000, B - off)
            llr, u = e32.channel(snr, off, b)                      # float32 LLRs
            d32, f32 = e32.decode_llr(llr, packed=True)
            d64, _ = e64.decode_llr(llr, packed=True)              # converted to double on the device: same values
            bad = (d32 != d64).any(1)
            diff += int(bad.sum())
            ties += int((bad & ((f32 & 1) != 0)).sum())
            up = np.packbits(u, axis=1, bitorder="little").view(np.uint32)
            err32 += int((d32 != up).any(1).sum())
            err64 += int((d64 != up).any(1).sum())
        rows.append("| %s | %.1f | %d / %d | %.2e | %d | %.3e | %.3e |" % (prog, snr, diff, B, diff / B, ties, err64 / B, err32 / B))
        print(rows[-1], flush=True)
    e32.close(); e64.close()
if a.out:
    open(a.out, "w").write("\n".join(rows) + "\n")
