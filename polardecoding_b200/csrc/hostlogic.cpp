// Pure host logic of the Monte-Carlo loop (no CUDA): how the global frame index space is partitioned
// over ranks and how per-rank round results are merged so that the reference's stopping rule
// ("run until the BLE-th block error", /root/reference/SC_128.c:169, CASCL_1024_L8.c:239) keeps its
// meaning when frames are decoded in batches on several GPUs.  Exported through the C ABI so that the
// multi-rank path can be tested on CPU (tests/test_multirank_gloo.py) and re-used by other hosts.
#include "../../include/polargpu.h"

#include <algorithm>
#include <cstring>

extern "C" int pg_partition(uint64_t round_first, uint64_t chunk, int nranks, int rank, uint64_t budget,
                            uint64_t *start, uint64_t *count)
{
    if (!start || !count || nranks < 1 || rank < 0 || rank >= nranks || chunk == 0) return PG_ERR_ARG;
    const uint64_t lo = (uint64_t)rank * chunk;
    *start = round_first + lo;
    *count = (lo >= budget) ? 0 : std::min<uint64_t>(chunk, budget - lo);
    return PG_OK;
}

extern "C" int pg_truncate_info(const uint32_t *frame_info, size_t nframes, uint64_t need, pg_counters *part)
{
    if (!part || (nframes && !frame_info)) return PG_ERR_ARG;
    std::memset(part, 0, sizeof(*part));
    uint64_t seen = 0;
    for (size_t i = 0; i < nframes && seen < need; i++) {
        const uint32_t w = frame_info[i];
        part->frames++;
        if (w & 0xFFFFu) { part->err_blocks++; part->err_bits += (w & 0xFFFFu); seen++; }
        if (w & (1u << 16)) part->tie_frames++;
        if (w & (1u << 17)) part->crc_fail++;
        part->bp_sweeps += (w >> 24);
    }
    return PG_OK;
}

static void add(pg_counters *a, const pg_counters *b)
{
    a->frames += b->frames; a->err_blocks += b->err_blocks; a->err_bits += b->err_bits;
    a->tie_frames += b->tie_frames; a->crc_fail += b->crc_fail; a->bp_sweeps += b->bp_sweeps;
}

extern "C" int pg_merge_round(const pg_counters *round, int nranks, uint64_t target, int exact_stop, pg_counters *acc,
                              int *cut_rank, uint64_t *need)
{
    if (!round || !acc || !cut_rank || !need || nranks < 1) return PG_ERR_ARG;
    *cut_rank = -1;
    *need = 0;
    for (int q = 0; q < nranks; q++) {
        if (exact_stop && target && acc->err_blocks + round[q].err_blocks >= target) {
            *cut_rank = q;
            *need = target - acc->err_blocks;
            return PG_OK;
        }
        add(acc, &round[q]);
    }
    return PG_OK;
}
