"""Quick device-side timing of the decode kernels (development aid; bench.py is the contract)."""
import sys
import time

sys.path.insert(0, ".")
from polardecoding_b200 import Engine

cases = [("CASCL_1024_L8", "f32", 2.0, 1 << 16), ("CASCL_1024_L8", "f64", 2.0, 1 << 14), ("SCL_1024", "f32", 2.0, 1 << 16),
         ("SC_1024", "f32", 2.0, 1 << 16), ("SC_128", "f32", 2.0, 1 << 19), ("CASCL_128", "f32", 2.0, 1 << 18),
         ("BP_1024", "f32", 2.5, 1 << 13), ("BP_1024", "f64", 2.5, 1 << 11), ("BP_1024", "h2", 2.5, 1 << 14), ("BP_128", "h2", 2.5, 1 << 17), ("BP_128", "f32", 2.5, 1 << 16)]
if len(sys.argv) > 1:
    cases = [c for c in cases if c[0] in sys.argv[1:] or c[1] in sys.argv[1:]]
for prog, real, snr, B in cases:
    for early in ([0, 1] if prog.startswith("BP") else [0]):
        eng = Engine(prog, real=real, bp_early_stop=early)
        eng.simulate_batch(snr, 0, min(B, 4096))
        t = time.time()
        acc, _ = eng.simulate_batch(snr, 1 << 20, B)
        wall = time.time() - t
        dms, cms = eng.last_kernel_ms()
        K = eng.K
        print("%-14s %s early=%d B=%d decode %.2f ms channel %.2f ms wall %.1f ms -> %.3f Mframes/s %.4f Gb/s | FER %.4g sweeps/frame %.1f ties %d"
              % (prog, real, early, B, dms, cms, wall * 1e3, B / dms / 1e3, B * K / dms / 1e6, acc.err_blocks / acc.frames,
                 acc.bp_sweeps / acc.frames, acc.tie_frames), flush=True)
        eng.close()
