"""Device-resident fp32-vs-fp64 flip count of a list decoder on identical float LLRs (both decoders on the GPU; the fp64
instantiation is bit-exact with the reference, tests/test_gpu_parity.py).  Frames stay in HBM: Philox channel -> both decode
kernels -> packed decisions compared on the device.   python tools/flip_rate_dev.py [--frames 4000000] [--prog CASCL_1024_L8]"""
import argparse
import os
import sys

import torch
from scipy.stats import chi2

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from polardecoding_b200 import Engine  # noqa: E402


def flips(prog, ebn0, frames, seed=99, first=0):
    """-> (frames decoded, frames whose decisions differ, of those: flagged as exact-tie frames by the fp32 kernel)"""
    e32 = Engine(prog, real="f32", seed=seed, data_mode=1)
    e64 = Engine(prog, real="f64", seed=seed, data_mode=1)
    N, W = e32.N, e32.N // 32
    chunk = e32.wave_frames() * 6
    llr = torch.empty(chunk * N, dtype=torch.float32, device="cuda")
    o32 = torch.empty((chunk, W), dtype=torch.int32, device="cuda")
    o64 = torch.empty((chunk, W), dtype=torch.int32, device="cuda")
    f32 = torch.empty(chunk, dtype=torch.int32, device="cuda")
    diff = ties = done = 0
    while done < frames:
        b = min(chunk, frames - done)
        e32.channel_device(ebn0, first + done, b, llr.data_ptr(), None)
        e32.sync()
        e32.decode_llr_device(llr.data_ptr(), False, b, o32.data_ptr(), f32.data_ptr())
        e64.decode_llr_device(llr.data_ptr(), False, b, o64.data_ptr(), None)     # converted to double on the device: same values
        e32.sync(); e64.sync()
        bad = (o32[:b] != o64[:b]).any(1)
        diff += int(bad.sum())
        ties += int((bad & ((f32[:b] >> 16) & 1).bool()).sum())
        done += b
    e32.close(); e64.close()
    return done, diff, ties


def upper95(k, n):
    """one-sided 95 % upper confidence bound of a Poisson rate from k events in n trials"""
    return 0.5 * chi2.ppf(0.95, 2 * (k + 1)) / n


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=4000000)
    ap.add_argument("--prog", default="CASCL_1024_L8")
    ap.add_argument("--ebn0", default="1.0,1.5,2.0")
    a = ap.parse_args()
    print("| program | Eb/N0 dB | frames | differing | rate | 95 %% upper bound | tie-flagged |\n|---|---|---|---|---|---|---|")
    for snr in [float(x) for x in a.ebn0.split(",")]:
        n, d, t = flips(a.prog, snr, a.frames)
        print("| %s | %.1f | %d | %d | %.2e | %.2e | %d |" % (a.prog, snr, n, d, d / n, upper95(d, n), t), flush=True)
