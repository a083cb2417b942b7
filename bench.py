#!/usr/bin/env python
"""bench.py -- throughput of the polar decoding hot path on B200 (contract in the task statement).

  python bench.py [--gpus N] [--steps K] [--warmup W]          (N>1: launched by torch.distributed.run, one rank per GPU)
  python bench.py --impl reference ...                          (the reference's own CPU code on the host cores)

Metric (BASELINE.json): decoded information Gbit/s for CA-SCL L=8 N=1024 (CRC-24, K=512) -- `value` -- and for BP N=1024,
100 sweeps -- the `bp_1024` object of the same line.  A step = one pass of the decode kernel over one batch of synthetic
frames (Philox channel kernel at the Eb/N0 of SURVEY 8d) already resident in HBM; the batch (256 MB / 128 MB of LLRs) is
larger than the 126 MB L2, so successive steps re-read it from HBM.  `e2e` = the same metric through the C-ABI call a host
makes (pg_decode_llr_packed) with pinned HOST buffers: H2D copy of the LLRs, decode, D2H copy of the packed decisions and
per-frame flags, every step.  Timing: CUDA events on the library's stream, barrier + synchronize on both sides, max over ranks.
Roofline: the path is ALU-bound: achieved = frames x algorithmic lane-ops per frame / kernel time against
148 SMs x 128 lanes x sm_max_mhz (MEASURED_PEAKS.json); HBM GB/s of the streamed LLRs is reported beside it."""
import argparse
import ctypes as C
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line: NCCL's banner / debug log (it writes to stdout by default) goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
if os.environ.get("NCCL_DEBUG", "VERSION").upper() in ("VERSION", "WARN"):   # the version banner ignores NCCL_DEBUG_FILE
    os.environ["NCCL_DEBUG"] = "NONE"

N, K_INFO = 1024, 512
OPS_CASCL = 30016 * 27 + 35906 * 2 + 11185 * 11 + 533 * 160 + 36000 + 28000   # ~1.15 M lane-ops / frame
OPS_BP_SWEEP = 573440                                                          # per sweep, N=1024 (20 stage passes; the kernel executes 18)
OPS_BP_SWEEP_EXEC = OPS_BP_SWEEP * 18 // 20                                    # the two passes per sweep whose outputs nothing reads are not executed
# the other configs[] of BASELINE.json (CHK = 27 ops, g = 2, PHI = 11, 2L-sort = 160, partial-sum XOR = 1)
OPS_SC_128 = 448 * 29                                                          # 13.0 k
OPS_SC_1024 = 5120 * 29                                                        # 148 k
OPS_BP_128_SWEEP = 1792 * 28                                                   # 7 stages x 128 x 2 passes; 5.0 M per 100 sweeps
OPS_SCL_1024 = 29572 * 27 + 35400 * 2 + 11000 * 11 + 509 * 160 + 36000         # ~1.11 M (no CRC)
EBN0_CASCL, EBN0_BP = 2.0, 2.5
# frames per step: sized so that the driver's 20 timed steps last >= 2 s on one B200 (6 GB / 200 MB of fp32 LLRs per GPU), whole waves
B_CASCL, B_BP = 3 << 19, 3 << 14


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return p, "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md recipe)"""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu):
        self.gpu, self.lines, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------ CPU arms (oracle/_ref = the compiled reference)
def _cpu_worker(args):
    kind, prog, nframes, seed, ebn0 = args
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import Oracle, RefHarness, awgn_llr
    rng = np.random.default_rng(seed)
    llr = awgn_llr(rng, N, nframes, ebn0)           # all-zero codeword + AWGN: decoding work does not depend on the payload
    dec = RefHarness(prog) if kind == "reference" else Oracle(prog)
    t = time.perf_counter()
    out = dec.decode(llr)
    dt = time.perf_counter() - t
    out = out[0] if isinstance(out, tuple) else out
    return dt, int((out != 0).any(1).sum())


def cpu_arm(prog, ebn0, frames_per_core, cores=None):
    """frames/s of the reference decoder on `cores` host cores (one process per core), bounded sample"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import have_ref, build_port
    kind = "reference" if have_ref(prog) else "port"
    if kind == "port":
        build_port()
    cores = cores or len(os.sched_getaffinity(0)) or 1
    ctx = mp.get_context("spawn")
    t = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(kind, prog, frames_per_core, 1000 + i, ebn0) for i in range(cores)])
    wall = time.perf_counter() - t
    busy = max(r[0] for r in res)                    # slowest worker = time for all cores to finish their share
    fps = cores * frames_per_core / busy
    return {"kind": kind, "cores": cores, "frames": cores * frames_per_core, "seconds": busy, "wall": wall, "fps": fps,
            "block_errors": sum(r[1] for r in res)}


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fpc = 400                                         # frames per core per step: ~2 s of CASCL_1024_L8 on one core
    for _ in range(a.warmup):
        cpu_arm("CASCL_1024_L8", EBN0_CASCL, 4)
    t0 = time.perf_counter()
    fps, r = [], None
    for _ in range(a.steps):
        r = cpu_arm("CASCL_1024_L8", EBN0_CASCL, fpc)
        fps.append(r["fps"])
    ms = (time.perf_counter() - t0) * 1e3 / max(1, a.steps)
    f = sum(fps) / len(fps)
    rb = cpu_arm("BP_1024", EBN0_BP, 100)
    val = f * K_INFO / 1e9
    line = {"impl": "reference", "metric": "decoded info Gbps: CA-SCL L=8 N=1024 (K=512, CRC-24)", "value": val, "unit": "Gbit/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "CASCL_1024_L8 decode (CASCL(), N=1024 K=512 r=24 L=8) at Eb/N0 %.1f dB" % EBN0_CASCL,
                       "frames_per_step": r["frames"], "host": "reference C decoder, one process per core"},
            "frames_per_s": f,
            "cpu_baseline": {"value": val, "unit": "Gbit/s", "cores": r["cores"], "kind": r["kind"],
                             "sample": "%d frames per core per step, %d steps, unmodified CASCL() compiled -O2 from the reference source" % (fpc, a.steps)},
            "bp_1024": {"value": rb["fps"] * K_INFO / 1e9, "unit": "Gbit/s", "frames_per_s": rb["fps"], "cores": rb["cores"], "kind": rb["kind"],
                        "sample": "%d frames per core, BP() 100 sweeps" % 100},
            "e2e": {"value": val, "unit": "Gbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ GPU arm
def bind_to_gpu_numa_node(torch, local):
    """One rank per GPU on a multi-socket host: run on (and first-touch the pinned e2e buffers from) the CPUs next to
    this rank's GPU, otherwise eight ranks pull their host LLRs through one socket's memory controllers."""
    try:
        pr = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        cpus = set()
        for part in open("/sys/bus/pci/devices/%s/local_cpulist" % bdf).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            node = open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip()
            return {"gpu": bdf, "numa_node": int(node), "cpus": len(cpus)}
    except Exception:
        pass
    return None


def timed_steps(torch, dist, world, ext_stream, fn, steps, warmup, count=None):
    for _ in range(warmup):
        fn()
    if count:
        count("start")
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext_stream)
    for _ in range(steps):
        fn()
    e1.record(ext_stream)
    torch.cuda.synchronize()
    if count:
        count("stop")
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        ms = float(t.item())
    return ms


def wall_steps(torch, dist, world, fn, steps, warmup, count=None):
    for _ in range(warmup):
        fn()
    if count:
        count("start")
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    if count:
        count("stop")
    ms = (time.perf_counter() - t0) * 1e3
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        ms = float(t.item())
    return ms


def run_gpu_arm(a):
    import torch
    import torch.distributed as dist
    from polardecoding_b200 import Engine
    from polardecoding_b200.capi import comm_unique_id

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(torch, local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pk, pk_src = peaks()
    peak_ops = 148 * 128 * pk["sm_max_mhz"] * 1e6
    sampler = ClockSampler(local)
    launches = 0
    res = {}

    traffic_db, traffic_src = {}, None
    for cand in ("r2_traffic.json", "r1b_traffic.json"):
        try:
            traffic_db, traffic_src = json.load(open(os.path.join(ROOT, "profiles", cand))), "profiles/" + cand
            break
        except Exception:
            pass

    def bench_one(prog, ebn0, B, ops_per_frame, tkey=None, real="f32", e2e=True, n=N, k_info=K_INFO, ops_sweep=OPS_BP_SWEEP, **over):
        nonlocal launches
        eng = Engine(prog, real=real, device=local, rank=rank, nranks=world, seed=1024, data_mode=0, **over)
        if world > 1:  # the library's own NCCL communicator: the id travels over torch.distributed
            ids = [comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            eng.comm_init(ids[0])
        wave = eng.wave_frames()                      # frames a full grid decodes concurrently
        B = max(1, round(B / wave)) * wave             # whole waves: every SM stays busy until the launch ends
        st = torch.cuda.ExternalStream(eng.stream_ptr())
        esz = 8 if real == "f64" else 4
        llr = torch.empty(B * n, dtype=torch.float64 if real == "f64" else torch.float32, device="cuda")
        truth = torch.empty(B * (n // 32), dtype=torch.int32, device="cuda")
        info = torch.empty(B, dtype=torch.int32, device="cuda")
        first = (1 << 32) + rank * B                                  # disjoint Philox frame ranges per rank
        eng.channel_device(ebn0, first, B, llr.data_ptr(), truth.data_ptr())
        eng.sync()
        mark = [0]

        def count(what):  # kernels of this library launched inside the timed regions only
            nonlocal launches
            if what == "start":
                mark[0] = eng.launches()
            else:
                launches += eng.launches() - mark[0]
        eng.counters_read(reset=True)
        ms = timed_steps(torch, dist, world, st, lambda: eng.decode_count_device(llr.data_ptr(), B, truth.data_ptr(), None, info.data_ptr()), a.steps, a.warmup, count)
        cnt = eng.counters_read(reset=True)
        if world > 1:
            cnt = eng.allreduce_counters(cnt)                          # the one collective of the path: final counters
        per_step = ms / a.steps
        fps = world * B / (per_step * 1e-3)
        steps_total = a.steps + a.warmup
        sweeps = cnt.bp_sweeps / max(1, cnt.frames)
        ops = ops_per_frame if ops_per_frame else ops_sweep * sweeps
        byts = n * esz + 2 * (n // 8) + 4                              # algorithmic bytes per frame: LLRs in, truth in, decisions + word out
        peak = peak_ops / 2 if real == "f64" else peak_ops             # fp64: 64 lanes per SM
        out = {"frames_per_s": fps, "gbps": fps * k_info / 1e9, "ms_per_step": per_step, "frames_per_step": world * B,
               "timed_s": ms / 1e3, "fer": cnt.err_blocks / max(1, cnt.frames), "frames_counted": int(cnt.frames), "tie_frames": int(cnt.tie_frames),
               "sweeps_per_frame": sweeps,
               "roofline": {"bound": "alu", "achieved": (fps / world) * ops / 1e12, "peak": peak / 1e12, "unit": "Tlaneop/s",
                            "frac": (fps / world) * ops / peak, "traffic": None, "ops_per_frame": ops,
                            "peak_source": pk_src + (" sm_max_mhz x 148 SMs x 64 fp64 lanes" if real == "f64" else " sm_max_mhz x 148 SMs x 128 lanes"),
                            "frac_of_fp32_lane_peak": (fps / world) * ops / peak_ops,
                            "hbm_gbs": (fps / world) * byts / 1e9, "hbm_frac": (fps / world) * byts / 1e9 / pk["hbm_gbs"]}}
        tr = traffic_db.get(tkey) if tkey else None
        if tr:  # DRAM bytes of this kernel from the committed ncu --set full capture, scaled to this launch's frame count
            out["roofline"]["traffic"] = tr["dram_bytes_per_launch"] * B / tr["frames_per_launch"]
            out["roofline"]["traffic_source"] = "%s (ncu dram__bytes_read+write, %d-frame launch)" % (traffic_src, tr["frames_per_launch"])
            out["roofline"]["algorithmic_bytes"] = B * byts
            out["roofline"]["traffic_over_algorithmic"] = out["roofline"]["traffic"] / (B * byts)
        assert cnt.frames == world * B * steps_total, (cnt.frames, world, B, steps_total)
        if not e2e:
            eng.close()
            del llr, truth, info
            return out
        # ---- e2e: host LLRs (pinned) -> C ABI -> host decisions, every step.  The host buffer holds Be <= 2^18 frames (1 GB);
        # a step sends it B/Be times (one C-ABI call each, every call copies its inputs H2D and its results D2H)
        reps = max(1, -(-B // (1 << 18)))
        Be = -(-B // reps // wave) * wave
        h_llr = torch.empty(Be * n, dtype=llr.dtype).pin_memory()
        h_llr.copy_(llr[: Be * n])
        h_out = torch.empty(Be * (n // 32), dtype=torch.int32).pin_memory()
        h_flags = torch.empty(Be, dtype=torch.int32).pin_memory()

        def e2e_step():
            for _ in range(reps):
                eng.decode_llr_host_ptr(h_llr.data_ptr(), real == "f64", Be, h_out.data_ptr(), h_flags.data_ptr())
        ms_e = wall_steps(torch, dist, world, e2e_step, a.steps, max(1, min(a.warmup, 2)), count)
        fps_e = world * Be * reps / (ms_e / a.steps * 1e-3)
        out["e2e"] = {"value": fps_e * k_info / 1e9, "unit": "Gbit/s", "h2d_bytes_per_step": Be * reps * n * esz, "d2h_bytes_per_step": Be * reps * ((n // 8) + 4),
                      "frames_per_s": fps_e, "frames_per_step": world * Be * reps, "calls_per_step": reps, "timed_s": ms_e / 1e3}
        if e2e == "f16":   # optional input format PG_LLR_F16: the same frames rounded to binary16, half the host-to-device bytes
            h16 = torch.empty(Be * n, dtype=torch.float16).pin_memory()
            h16.copy_(llr[: Be * n])

            def e2e16_step():
                for _ in range(reps):
                    eng.decode_llr_host_ptr(h16.data_ptr(), 2, Be, h_out.data_ptr(), h_flags.data_ptr())
            ms_h = wall_steps(torch, dist, world, e2e16_step, a.steps, 1, count)
            fps_h = world * Be * reps / (ms_h / a.steps * 1e-3)
            out["e2e_f16"] = {"value": fps_h * k_info / 1e9, "unit": "Gbit/s", "h2d_bytes_per_step": Be * reps * n * 2, "d2h_bytes_per_step": Be * reps * ((n // 8) + 4),
                              "frames_per_s": fps_h, "note": "optional PG_LLR_F16 input (quantised LLRs: FER-level parity, tests/test_gpu_parity.py::test_fp16_llr_input)"}
        eng.close()
        del llr, truth, info
        return out

    def bench_simulate(prog, ebn0, seconds):
        """pg_simulate, the Monte-Carlo loop itself (channel + decode + count + the NCCL counter exchange): frames/s with a frame
        budget (device-side accumulation, ONE all-reduce) and with the reference's exact stopping rule (one all-reduce per round, on
        its own stream, the next round already running)."""
        eng = Engine(prog, real="f32", device=local, rank=rank, nranks=world, seed=1024, data_mode=0)
        if world > 1:
            ids = [comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            eng.comm_init(ids[0])
        wave = eng.wave_frames()
        eng.simulate(ebn0, 0, max_frames=world * wave * 8)                  # warm-up: buffers, NCCL channels
        res = {}
        rate = 12e6 * world                                                # frames/s guess used only to size the runs
        for name, kw in (("frame_budget", {"max_frames": int(rate * seconds)}), ("exact_stop", {"target_err_blocks": int(rate * seconds * 0.0038), "exact_stop": True})):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            c = eng.simulate(ebn0, 1 << 33, **kw)
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            rounds, ar = eng.simulate_stats()
            res[name] = {"frames": int(c.frames), "err_blocks": int(c.err_blocks), "seconds": dt, "frames_per_s": c.frames / dt, "gbps": c.frames / dt * K_INFO / 1e9,
                         "rounds": rounds, "allreduces": ar}
        eng.close()
        return res

    sampler.start()
    legs = set(a.legs.split(",")) if a.legs else None           # development aid (profiling one kernel): the driver runs all legs

    def want(name):
        return legs is None or name in legs
    res["cascl"] = bench_one("CASCL_1024_L8", EBN0_CASCL, B_CASCL, OPS_CASCL, tkey="cascl", e2e="f16" if legs is None else False)
    if want("bp"):
        res["bp"] = bench_one("BP_1024", EBN0_BP, B_BP, OPS_BP_SWEEP * 100, tkey="bp", e2e=legs is None)
    if legs is None:
        res["bp_stop"] = bench_one("BP_1024", EBN0_BP, B_BP * 4, 0, bp_early_stop=1)
        res["bp_gm"] = bench_one("BP_1024", EBN0_BP, B_BP * 4, 0, e2e=False, bp_early_stop=3)      # optional codeword ("G-matrix") stop rule
        res["bp_h2"] = bench_one("BP_1024", EBN0_BP, B_BP, OPS_BP_SWEEP * 100, real="h2", e2e=False)   # optional packed-half mode (FER-only parity)
    # the bit-exact (fp64) instantiation of both kernels: the configuration that meets north_star's "bit-exact decisions" to the letter
    if want("cascl64"):
        res["cascl64"] = bench_one("CASCL_1024_L8", EBN0_CASCL, B_CASCL // 3, OPS_CASCL, tkey="cascl64", real="f64", e2e=legs is None)
    if want("bp64"):
        res["bp64"] = bench_one("BP_1024", EBN0_BP, B_BP // 2, OPS_BP_SWEEP * 100, tkey="bp64", real="f64", e2e=legs is None)
    if legs is not None:
        sampler.stop()
        if rank == 0:
            print(json.dumps({k: {"frames_per_s": v["frames_per_s"], "frames_per_step": v["frames_per_step"], "roofline": v["roofline"]} for k, v in res.items()}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    res["bp64_stop"] = bench_one("BP_1024", EBN0_BP, B_BP * 2, 0, real="f64", bp_early_stop=1)
    # the other configs[] of BASELINE.json (device-resident inputs; the N = 128 programs have K = 64)
    res["sc_128"] = bench_one("SC_128", 2.0, 1 << 22, OPS_SC_128, e2e=False, n=128, k_info=64)
    res["bp_128"] = bench_one("BP_128", 2.5, 1 << 19, OPS_BP_128_SWEEP * 100, e2e=False, n=128, k_info=64)
    res["sc_1024"] = bench_one("SC_1024", 2.0, 3 << 17, OPS_SC_1024, e2e=False)   # 1.5 GB of LLRs (> L2); one frame per lane reads 32 rows per request
    res["scl_1024"] = bench_one("SCL_1024", 2.0, 3 << 18, OPS_SCL_1024, e2e=False)
    # the author's deepest list size (myResult_1024/CASCL_L32.dat): one frame per warp; ops ~ 4 x the L = 8 count except the 64-candidate selection
    res["cascl_l32"] = bench_one("CASCL_1024_L8", EBN0_CASCL, 3 << 17, 4 * (OPS_CASCL - 533 * 160) + 533 * 6 * 160, e2e=False, list_size=32)
    sim = bench_simulate("CASCL_1024_L8", EBN0_CASCL, 2.0)
    clocks = sampler.stop()

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        c = cpu_arm("CASCL_1024_L8", EBN0_CASCL, 2400)                # bounded sample: ~11 s on every core
        cb = cpu_arm("BP_1024", EBN0_BP, 480)                         # ~10 s on every core
        cpu = {"value": c["fps"] * K_INFO / 1e9, "unit": "Gbit/s", "cores": c["cores"], "kind": c["kind"], "frames_per_s": c["fps"],
               "sample": "%d frames of CASCL_1024_L8 per core at %.1f dB (%.1f s), unmodified reference CASCL() compiled -O2" % (2400, EBN0_CASCL, c["seconds"]),
               "bp_1024": {"value": cb["fps"] * K_INFO / 1e9, "frames_per_s": cb["fps"], "sample": "480 frames per core, BP() 100 sweeps (%.1f s)" % cb["seconds"]}}
    if rank == 0:
        r = res["cascl"]
        line = {"metric": "decoded info Gbps: CA-SCL L=8 N=1024 (K=512, CRC-24)", "value": r["gbps"], "unit": "Gbit/s", "n_gpus": world,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "CASCL_1024_L8 (N=1024 K=512 r=24 L=8) at Eb/N0 %.1f dB, PN-63 payload, Philox AWGN" % EBN0_CASCL,
                           "frames_per_step": r["frames_per_step"], "l2": "inputs larger than L2 (%d MB of LLRs per GPU per step)" % (r["frames_per_step"] // world * N * 4 >> 20),
                           "arith": "fp32 throughput mode; fp64 parity mode is bit-exact with the reference (tests/test_gpu_parity.py)",
                           "partition": "rank r decodes its own Philox frame range; one NCCL all-reduce of the final counters",
                           "host_numa": numa},
                "frames_per_s": r["frames_per_s"], "fer": r["fer"], "tie_frames": r["tie_frames"], "frames_counted": r["frames_counted"],
                "timed_s": r["timed_s"], "roofline": r["roofline"], "e2e": r["e2e"], "e2e_f16": r.get("e2e_f16"), "gpu_launches": launches, "clocks": clocks,
                "bp_1024": {"value": res["bp"]["gbps"], "unit": "Gbit/s", "frames_per_s": res["bp"]["frames_per_s"], "ms_per_step": res["bp"]["ms_per_step"],
                            "frames_per_step": res["bp"]["frames_per_step"], "sweeps": 100, "fer": res["bp"]["fer"], "timed_s": res["bp"]["timed_s"],
                            "roofline": dict(res["bp"]["roofline"], frac_on_executed_work=res["bp"]["roofline"]["frac"] * OPS_BP_SWEEP_EXEC / OPS_BP_SWEEP,
                                             note="frac counts the reference's 20 stage passes per sweep; the kernel executes 18 (the two whose outputs nothing reads are skipped): frac_on_executed_work"),
                            "e2e": res["bp"]["e2e"],
                            "fixed_point_stop": {"value": res["bp_stop"]["gbps"], "frames_per_s": res["bp_stop"]["frames_per_s"],
                                                 "sweeps_per_frame": res["bp_stop"]["sweeps_per_frame"], "fer": res["bp_stop"]["fer"],
                                                 "roofline": res["bp_stop"]["roofline"], "e2e": res["bp_stop"]["e2e"], "note": "same decisions as 100 sweeps (bit-exact stop)"},
                            "gmatrix_stop": {"value": res["bp_gm"]["gbps"], "frames_per_s": res["bp_gm"]["frames_per_s"],
                                             "sweeps_per_frame": res["bp_gm"]["sweeps_per_frame"], "fer": res["bp_gm"]["fer"],
                                             "note": "optional flag (bp_early_stop bit 1): stop when the decisions form a codeword; not in the reference, FER-level parity only"},
                            "half2_mode": {"value": res["bp_h2"]["gbps"], "frames_per_s": res["bp_h2"]["frames_per_s"], "fer": res["bp_h2"]["fer"],
                                           "frac_of_fp32_lane_roofline": res["bp_h2"]["roofline"]["frac"],
                                           "note": "PG_REAL_H2, optional flag: two frames per __half2, a numerically different decoder judged on FER only (not the headline)"}}}
        def leg(x, **extra):
            d = {"value": x["gbps"], "unit": "Gbit/s", "frames_per_s": x["frames_per_s"], "ms_per_step": x["ms_per_step"], "frames_per_step": x["frames_per_step"],
                 "timed_s": x["timed_s"], "fer": x["fer"], "roofline": x["roofline"]}
            if "e2e" in x:
                d["e2e"] = x["e2e"]
            d.update(extra)
            return d
        line["f64_parity_mode"] = {"note": "same kernels instantiated in double: decisions bit-exact with the reference (tests/test_gpu_parity.py); roofline.frac is "
                                           "against the fp64 peak (64 lanes per SM), frac_of_fp32_lane_peak against the line's headline peak; e2e sends fp64 LLRs (8 KiB per frame)",
                                   "cascl_1024_l8": leg(res["cascl64"]),
                                   "bp_1024": leg(res["bp64"], sweeps=100,
                                                  fixed_point_stop=leg(res["bp64_stop"], sweeps_per_frame=res["bp64_stop"]["sweeps_per_frame"],
                                                                       note="bit-exact AND early-stopped: the sweeps after the fixed point change nothing (BP_1024.c:393 runs them anyway)"))}
        line["configs"] = {"note": "the other configs[] of BASELINE.json, device-resident inputs, fp32; ops_per_frame from SURVEY 8d",
                           "sc_128": leg(res["sc_128"]), "bp_128": leg(res["bp_128"], sweeps=100), "sc_1024": leg(res["sc_1024"]), "scl_1024": leg(res["scl_1024"]),
                           "cascl_1024_l32": leg(res["cascl_l32"], note="L = 32 (SURVEY 8f.2): ops_per_frame is an estimate, 4 x the L = 8 count with a 64-candidate selection")}
        line["simulate"] = dict(sim, note="pg_simulate = channel + decode + count + counter exchange, wall clock, max over ranks",
                                collective=("ncclAllReduce (sum) of %d x 8 u64 per round on a second stream; ONE at the end of a frame-budget run" % world) if world > 1 else "none (1 rank)",
                                frac_of_kernel_rate={k: v["frames_per_s"] / r["frames_per_s"] for k, v in sim.items()})
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--legs", default=None, help="development aid: only these device-resident legs (cascl always; bp, cascl64, bp64), short JSON")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_gpu_arm(a)


if __name__ == "__main__":
    main()
