"""Development aid: host-pointer decode (pg_decode_llr_packed) time vs batch size -> per-chunk steady state and fill cost."""
import sys, time
sys.path.insert(0, ".")
import torch
from polardecoding_b200 import Engine

prog = sys.argv[1] if len(sys.argv) > 1 else "CASCL_1024_L8"
eng = Engine(prog, real="f32")
N, wave = eng.N, int(eng.wave_frames())
maxw = 8
llr, _ = eng.channel(2.0, 0, wave)
h = torch.empty(maxw * wave * N, dtype=torch.float32).pin_memory()
hv = h.view(maxw, wave * N)
hv[:] = torch.from_numpy(llr).reshape(1, -1)
out = torch.empty(maxw * wave * (N // 32), dtype=torch.int32).pin_memory()
fl = torch.empty(maxw * wave, dtype=torch.int32).pin_memory()
for w in (1, 2, 3, 4, 6, 8):
    B = w * wave
    for _ in range(2):
        eng.decode_llr_host_ptr(h.data_ptr(), False, B, out.data_ptr(), fl.data_ptr())
    t = time.perf_counter()
    for _ in range(5):
        eng.decode_llr_host_ptr(h.data_ptr(), False, B, out.data_ptr(), fl.data_ptr())
    dt = (time.perf_counter() - t) / 5
    print("%d waves (%d frames): %.3f ms  -> %.2f Mframes/s, %.1f GB/s H2D" % (w, B, dt * 1e3, B / dt / 1e6, B * N * 4 / dt / 1e9), flush=True)
