"""FER parity sweep: runs every Eb/N0 point the reference's captured result files publish
through the C ABI (pg_simulate: Philox channel + decode + on-device count) and compares block-error rates with
a 95 % interval.  Both numbers are Monte-Carlo estimates stopped at a fixed number of block errors, so the
relative standard deviation of their difference is sqrt(1/e_ref + 1/e_gpu); |z| < 1.96 <=> inside the interval.

  python tools/fer_sweep.py [--errors 400] [--real f32|f64] [--out profiles/r1_fer_parity.md]
"""
import argparse
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from polardecoding_b200 import Engine  # noqa: E402

# program, source inside the reference's zips, errors the reference stopped at, {Eb/N0: BLER}
REF = [
    ("SC_128", "myResult_128/SC128out.txt", 100, {1.0: 0.396825, 1.5: 0.274725, 2.0: 0.141443, 2.5: 0.066445, 3.0: 0.020982, 3.5: 0.006499, 4.0: 0.001880}),
    ("SC_1024", "myResult_1024/SC1024out.dat", 100, {1.0: 0.730, 1.5: 0.375, 2.0: 0.0901, 2.5: 0.01451, 3.0: 1.768e-3, 3.5: 1.964e-4}),
    ("SCL_128", "myResult_128/SCL128out_errblock50.dat (L=8)", 50, {1.0: 0.2451, 1.5: 0.1217, 2.0: 0.05917, 2.5: 0.02560, 3.0: 8.697e-3, 3.5: 2.963e-3}),
    ("CASCL_128", "myResult_128/CASCL_128_L8.txt (SEED 8392)", 200, {1.0: 0.2372, 1.5: 0.1168, 2.0: 0.04182, 2.5: 9.499e-3, 3.0: 1.888e-3}),
    ("SCL_1024", "myResult_1024/SCL1024out.dat (L=8)", 50, {1.0: 0.2203, 1.5: 0.04873, 2.0: 8.522e-3, 2.5: 2.318e-3, 3.0: 2.796e-4}),
    ("CASCL_1024_L8", "myResult_1024/CASCL_L8.dat lines 1-4", 200, {1.0: 0.3976, 1.5: 0.07130, 2.0: 4.088e-3, 2.5: 9.649e-5}),
    ("BP_128", "myResult_128/BP128_BER.txt (SEED 945)", 200, {1.0: 0.4386, 1.5: 0.2545, 2.0: 0.1133, 2.5: 0.04430, 3.0: 0.01913, 3.5: 6.813e-3, 4.0: 2.039e-3, 4.5: 5.304e-4}),
    ("BP_1024", "myResult_1024/BP1024out_NewSEED.dat (SEED 771)", 200, {1.0: 0.4494, 1.5: 0.1546, 2.0: 0.03292, 2.5: 5.675e-3, 3.0: 1.228e-3, 3.5: 2.173e-4}),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--errors", type=int, default=400)
    ap.add_argument("--real", default="f32")
    ap.add_argument("--out", default=None)
    ap.add_argument("--max-frames", type=int, default=60_000_000)
    ap.add_argument("--programs", nargs="*", default=None)
    a = ap.parse_args()
    lines = ["# FER parity sweep (%s arithmetic, Philox channel, random payload, stop at %d block errors or %d frames)" % (a.real, a.errors, a.max_frames), "",
             "| program | Eb/N0 dB | reference BLER (errors) | GPU BLER | GPU errors / frames | z | inside 95 % |", "|---|---|---|---|---|---|---|"]
    worst = 0.0
    outside = 0
    npts = 0
    t0 = time.time()
    for prog, src, eref, pts in REF:
        if a.programs and prog not in a.programs:
            continue
        eng = Engine(prog, real=a.real, seed=20261018, data_mode=1, bp_early_stop=1 if prog.startswith("BP") else 0)
        first = 0
        for snr, bler in pts.items():
            r = eng.simulate(snr, first_frame=first, target_err_blocks=a.errors, max_frames=a.max_frames, exact_stop=True)
            first += r.frames
            g = r.err_blocks / r.frames
            if r.err_blocks == 0:
                z = float("nan")
            else:
                z = (g - bler) / (bler * math.sqrt(1.0 / eref + 1.0 / r.err_blocks))
            ok = (not math.isnan(z)) and abs(z) < 1.96
            npts += 1
            outside += 0 if ok else 1
            if not math.isnan(z):
                worst = max(worst, abs(z))
            lines.append("| %s | %.1f | %.4g (%d) | %.4g | %d / %d | %+.2f | %s |" % (prog, snr, bler, eref, g, r.err_blocks, r.frames, z, "yes" if ok else "NO"))
            print(lines[-1], flush=True)
        eng.close()
    lines += ["", "%d points, %d outside the 95 %% interval (expected ~5 %% by chance), largest |z| = %.2f, wall %.0f s." % (npts, outside, worst, time.time() - t0),
              "Reference values: BASELINE.md section 2 (the author's captured stdout); BP uses the bit-exact fixed-point stop."]
    txt = "\n".join(lines) + "\n"
    if a.out:
        open(a.out, "w").write(txt)
    print(lines[-2])


if __name__ == "__main__":
    main()
