"""Development probe: device-resident decode of B frames as ONE full-grid launch against four quarter-grid engines running side by
side (each with its own stream and scratch), with long and with short launches.  python tools/lane_probe.py f32|f64"""
import os
import sys
import time

import torch

sys.path.insert(0, ".")
from polardecoding_b200 import Engine

real = sys.argv[1] if len(sys.argv) > 1 else "f32"
N = 1024
full = Engine("CASCL_1024_L8", real=real)
wave = full.wave_frames()
B = wave * 32
dt = torch.float64 if real == "f64" else torch.float32
llr = torch.empty(B * N, dtype=dt, device="cuda")
truth = torch.empty(B * 32, dtype=torch.int32, device="cuda")
full.channel_device(2.0, 0, B, llr.data_ptr(), truth.data_ptr())
full.sync()
esz = 8 if real == "f64" else 4


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t)
    return best


t1 = timeit(lambda: (full.decode_count_device(llr.data_ptr(), B, truth.data_ptr()), full.sync()))
print("%s one full-grid launch of %d frames: %.3f M frames/s" % (real, B, B / t1 / 1e6))
per_sm = wave // 148 // 4   # CTAs per SM of the full grid
os.environ["POLARGPU_LIST_CTAS_PER_SM"] = str(per_sm // 4)
q = [Engine("CASCL_1024_L8", real=real) for _ in range(4)]
del os.environ["POLARGPU_LIST_CTAS_PER_SM"]
for chunk_waves in (8, 1, 0.25):
    chunk = int(wave * chunk_waves)

    def run():
        off = 0
        i = 0
        while off < B:
            b = min(chunk, B - off)
            e = q[i % 4]
            e.decode_count_device(llr.data_ptr() + off * N * esz, b, truth.data_ptr() + off * 32 * 4)
            off += b; i += 1
        for e in q:
            e.sync()
    t = timeit(run)
    print("%s four quarter-grid engines, launches of %.2f waves: %.3f M frames/s" % (real, chunk_waves, B / t / 1e6))
