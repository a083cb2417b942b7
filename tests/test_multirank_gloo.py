"""CPU, world_size 2 (gloo): the multi-rank path of the Monte-Carlo loop.  The rules that pg_simulate applies on
every rank -- frame-space partition, merge of the per-rank round counters in global frame order, truncation at the
target-th block error (the reference's stopping rule, SC_128.c:169) -- are exported by libpolargpu.so as pure host
functions; here two processes drive them with a synthetic per-frame error pattern, exchanging counters with a
gloo all-reduce exactly where pg_simulate uses NCCL, and must reproduce the single-process sequential answer."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def frame_info(first, count, seed=5):
    """deterministic synthetic per-frame result words keyed by GLOBAL frame index (as Philox keys the real ones)"""
    idx = np.arange(first, first + count, dtype=np.uint64)
    h = (idx * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed)) >> np.uint64(40)
    err = (h % np.uint64(23) == 0)
    bits = np.where(err, (h % np.uint64(37)) + np.uint64(1), 0).astype(np.uint32)
    tie = ((h % np.uint64(101)) == 0).astype(np.uint32) << 16
    return (bits | tie).astype(np.uint32)


def sequential(first, target, max_frames):
    info = frame_info(first, 1 << 16)
    frames = blocks = bits = ties = 0
    for w in info:
        if (target and blocks >= target) or (max_frames and frames >= max_frames):
            break
        frames += 1
        if w & 0xFFFF:
            blocks += 1
            bits += int(w & 0xFFFF)
        if w & (1 << 16):
            ties += 1
    return frames, blocks, bits, ties


def run_rank(rank, world, port, first, target, max_frames, chunk0, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from polardecoding_b200 import load_library, PgCounters
    lib = load_library()
    lib.pg_partition.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.pg_merge_round.argtypes = [C.POINTER(PgCounters), C.c_int, C.c_uint64, C.c_int, C.POINTER(PgCounters), C.POINTER(C.c_int), C.POINTER(C.c_uint64)]
    lib.pg_truncate_info.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64, C.POINTER(PgCounters)]
    acc = PgCounters()
    nxt, chunk = first, chunk0
    while True:
        budget = (max_frames - acc.frames) if max_frames else (1 << 62)
        if max_frames and budget == 0:
            break
        s, c = C.c_uint64(), C.c_uint64()
        lib.pg_partition(nxt, chunk, world, rank, budget, C.byref(s), C.byref(c))
        info = frame_info(s.value, c.value)
        rnd = (PgCounters * world)()
        mine = PgCounters()
        lib.pg_truncate_info(info.ctypes.data, len(info), 1 << 62, C.byref(mine))   # = full count of my chunk
        t = torch.zeros(world * 8, dtype=torch.int64)
        t[rank * 8: rank * 8 + 8] = torch.tensor([mine.frames, mine.err_blocks, mine.err_bits, mine.tie_frames, mine.crc_fail, mine.bp_sweeps, 0, 0])
        dist.all_reduce(t)                                                          # pg_simulate: ncclAllReduce
        for q in range(world):
            rnd[q].frames, rnd[q].err_blocks, rnd[q].err_bits, rnd[q].tie_frames = [int(v) for v in t[q * 8: q * 8 + 4]]
        cut, need = C.c_int(), C.c_uint64()
        lib.pg_merge_round(rnd, world, target, 1, C.byref(acc), C.byref(cut), C.byref(need))
        if cut.value >= 0:
            part = PgCounters()
            if cut.value == rank:
                lib.pg_truncate_info(info.ctypes.data, len(info), need.value, C.byref(part))
            p = torch.tensor([part.frames, part.err_blocks, part.err_bits, part.tie_frames], dtype=torch.int64)
            dist.all_reduce(p)
            acc.frames += int(p[0]); acc.err_blocks += int(p[1]); acc.err_bits += int(p[2]); acc.tie_frames += int(p[3])
            break
        if (target and acc.err_blocks >= target) or (max_frames and acc.frames >= max_frames):
            break
        nxt += world * chunk
        chunk = min(chunk * 2, 4096)
    out[rank] = (acc.frames, acc.err_blocks, acc.err_bits, acc.tie_frames)
    dist.destroy_process_group()


def run_rank_pipelined(rank, world, port, first, target, max_frames, chunk0, out):
    """The control flow of the pipelined pg_simulate (csrc/api.cu): round i+1 is planned from the frames PLANNED so far and is
    computed (and all-reduced) before round i is merged; rounds in flight past the stopping point are discarded.  With a frame
    budget only, every rank walks its fixed schedule and ONE all-reduce combines the ranks at the end."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from polardecoding_b200 import load_library, PgCounters
    lib = load_library()
    lib.pg_partition.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.pg_merge_round.argtypes = [C.POINTER(PgCounters), C.c_int, C.c_uint64, C.c_int, C.POINTER(PgCounters), C.POINTER(C.c_int), C.POINTER(C.c_uint64)]
    lib.pg_truncate_info.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64, C.POINTER(PgCounters)]
    acc = PgCounters()
    allreduces = 0

    def count(info):
        c = PgCounters()
        lib.pg_truncate_info(info.ctypes.data, len(info), 1 << 62, C.byref(c))
        return [c.frames, c.err_blocks, c.err_bits, c.tie_frames, c.crc_fail, c.bp_sweeps, 0, 0]

    if target == 0:                                   # frame budget only
        CH = 4096
        nxt, left, tot = first, max_frames, np.zeros(8, dtype=np.int64)
        while left:
            chunk = min(CH, (left + world - 1) // world)
            s, c = C.c_uint64(), C.c_uint64()
            lib.pg_partition(nxt, chunk, world, rank, left, C.byref(s), C.byref(c))
            tot += np.array(count(frame_info(s.value, c.value)), dtype=np.int64)
            rnd = min(left, world * chunk)
            left -= rnd
            nxt += rnd
        t = torch.from_numpy(tot)
        dist.all_reduce(t)
        allreduces += 1
        out[rank] = (int(t[0]), int(t[1]), int(t[2]), int(t[3]), allreduces)
        dist.destroy_process_group()
        return

    state = {"next": first, "chunk": chunk0, "planned": 0}
    queue = []

    def enqueue():
        budget = (max_frames - state["planned"]) if max_frames else (1 << 62)
        s, c = C.c_uint64(), C.c_uint64()
        lib.pg_partition(state["next"], state["chunk"], world, rank, budget, C.byref(s), C.byref(c))
        info = frame_info(s.value, c.value)
        t = torch.zeros(world * 8, dtype=torch.int64)
        t[rank * 8: rank * 8 + 8] = torch.tensor(count(info))
        dist.all_reduce(t)                            # the exchange stream's ncclAllReduce of this round
        queue.append((info, t))
        state["planned"] += min(budget, world * state["chunk"])
        state["next"] += world * state["chunk"]
        state["chunk"] = min(state["chunk"] * 2, 4096)

    enqueue()
    allreduces += 1
    while True:
        while len(queue) < 3 and (not max_frames or state["planned"] < max_frames):   # pg_ctx::kRing - 1 rounds in flight
            enqueue()
            allreduces += 1
        info, t = queue.pop(0)
        rnd = (PgCounters * world)()
        for q in range(world):
            rnd[q].frames, rnd[q].err_blocks, rnd[q].err_bits, rnd[q].tie_frames = [int(v) for v in t[q * 8: q * 8 + 4]]
        cut, need = C.c_int(), C.c_uint64()
        lib.pg_merge_round(rnd, world, target, 1, C.byref(acc), C.byref(cut), C.byref(need))
        if cut.value >= 0:
            part = PgCounters()
            if cut.value == rank:
                lib.pg_truncate_info(info.ctypes.data, len(info), need.value, C.byref(part))
            p = torch.tensor([part.frames, part.err_blocks, part.err_bits, part.tie_frames], dtype=torch.int64)
            dist.all_reduce(p)
            allreduces += 1
            acc.frames += int(p[0]); acc.err_blocks += int(p[1]); acc.err_bits += int(p[2]); acc.tie_frames += int(p[3])
            break
        if acc.err_blocks >= target or (max_frames and acc.frames >= max_frames):
            break
        if not queue:
            enqueue()
            allreduces += 1
    out[rank] = (acc.frames, acc.err_blocks, acc.err_bits, acc.tie_frames, allreduces)
    dist.destroy_process_group()


@pytest.mark.parametrize("target,max_frames,chunk0", [(50, 0, 64), (7, 0, 512), (0, 3000, 128), (1000, 2500, 100)])
def test_two_ranks_reproduce_sequential_stop(target, max_frames, chunk0):
    world = 2
    port = 29500 + (os.getpid() + target + max_frames + chunk0) % 2000
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(run_rank, args=(world, port, 12345, target, max_frames, chunk0, out), nprocs=world, join=True)
    want = sequential(12345, target, max_frames)
    assert out[0] == out[1] == want, (dict(out), want)


@pytest.mark.parametrize("target,max_frames,chunk0", [(50, 0, 64), (7, 0, 512), (0, 3000, 128), (0, 20001, 64), (1000, 2500, 100), (3, 100, 64)])
def test_two_ranks_pipelined_loop_reproduces_sequential_stop(target, max_frames, chunk0):
    """same answer as the sequential walk although a round runs ahead of the merge; a frame budget alone needs one all-reduce"""
    world = 2
    port = 31500 + (os.getpid() + target + max_frames + chunk0) % 2000
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(run_rank_pipelined, args=(world, port, 12345, target, max_frames, chunk0, out), nprocs=world, join=True)
    want = sequential(12345, target, max_frames)
    assert out[0][:4] == out[1][:4] == want, (dict(out), want)
    assert out[0][4] == out[1][4]
    if target == 0:
        assert out[0][4] == 1
