import sys, time, torch
sys.path.insert(0, ".")
from polardecoding_b200 import Engine
for prog, n in (("SC_1024", 1024), ("SC_128", 128)):
    eng = Engine(prog, real="f32", seed=1024)
    wave = eng.wave_frames()
    for mult in (2, 3, 8):
        B = wave * mult
        llr = torch.empty(B * n, dtype=torch.float32, device="cuda"); truth = torch.empty(B * (n // 32), dtype=torch.int32, device="cuda"); info = torch.empty(B, dtype=torch.int32, device="cuda")
        eng.channel_device(2.0, 1 << 32, B, llr.data_ptr(), truth.data_ptr()); eng.sync()
        for variant, args in (("count+info", (truth.data_ptr(), None, info.data_ptr())), ("count only", (truth.data_ptr(), None, None))):
            eng.decode_count_device(llr.data_ptr(), B, *args); eng.sync()
            k = eng.last_kernel_ms()[0]
            st = torch.cuda.ExternalStream(eng.stream_ptr())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(20):
                eng.decode_count_device(llr.data_ptr(), B, *args)
            e1.record(st); torch.cuda.synchronize()
            print("%s B=%d (%d waves) %s: single launch %.3f ms = %.1f M/s; 20 back to back %.3f ms each = %.1f M/s" % (prog, B, mult, variant, k, B / k / 1e3, e0.elapsed_time(e1) / 20, B / (e0.elapsed_time(e1) / 20) / 1e3))
        del llr, truth, info
    eng.close()
