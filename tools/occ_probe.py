import sys, os
sys.path.insert(0, ".")
from polardecoding_b200 import Engine
e = Engine("CASCL_1024_L8", real="f32")
B = int(e.wave_frames()) * 4
e.simulate_batch(2.0, 0, B)
best = 1e9
for rep in range(3):
    acc, _ = e.simulate_batch(2.0, 1 << 22, B)
    best = min(best, e.last_kernel_ms()[0])
print("ctas/sm cap %s: wave %d B=%d decode %.3f ms -> %.3f Mframes/s" % (os.environ.get("POLARGPU_LIST_CTAS_PER_SM"), e.wave_frames(), B, best, B / best / 1e3))
