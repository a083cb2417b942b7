// Pure host logic of the Monte-Carlo loop (no CUDA): how the global frame index space is partitioned
// over ranks and how per-rank round results are merged so that the reference's stopping rule
// ("run until the BLE-th block error", /root/reference/SC_128.c:169, CASCL_1024_L8.c:239) keeps its
// meaning when frames are decoded in batches on several GPUs.  Exported through the C ABI so that the
// multi-rank path can be tested on CPU (tests/test_multirank_gloo.py) and re-used by other hosts.
#include "../../include/polargpu.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

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

// CRC parity table in the reference's file format (/root/reference/CRC_6.dat: K rows of r integers 0/1, row i = coefficients
// c0..c(r-1) of D^(r+i) mod g(D), i.e. row i of the systematic generator Gc that CASCL_1024_sys.c:49-561 spells out as a literal;
// UTF-16 with a byte-order mark as the author's editor saved it, or plain ASCII; blanks, CR/LF free-form).
// Row 0 is D^r mod g(D) = g(D) - D^r, so the file determines the polynomial; every further row must be D times its predecessor
// modulo g(D), otherwise the file is not a CRC table and the call fails.
extern "C" int pg_crc_table_load(const char *path, int K, int r, uint64_t *crc_poly, uint32_t *rows)
{
    if (!path || K < 1 || r < 1 || r > 32 || !crc_poly) return PG_ERR_ARG;
    FILE *f = std::fopen(path, "rb");
    if (!f) return PG_ERR_ARG;
    std::vector<unsigned char> raw;
    unsigned char buf[4096];
    size_t got;
    while ((got = std::fread(buf, 1, sizeof(buf), f)) > 0) raw.insert(raw.end(), buf, buf + got);
    std::fclose(f);
    std::vector<int> v;
    const bool le = raw.size() >= 2 && raw[0] == 0xFF && raw[1] == 0xFE, be = raw.size() >= 2 && raw[0] == 0xFE && raw[1] == 0xFF;
    const size_t step = (le || be) ? 2 : 1;
    for (size_t i = (le || be) ? 2 : 0; i + step <= raw.size(); i += step) {
        const unsigned ch = (step == 1) ? raw[i] : (le ? (raw[i] | (raw[i + 1] << 8)) : (raw[i + 1] | (raw[i] << 8)));
        if (ch == '0' || ch == '1') v.push_back((int)(ch - '0'));
        else if (ch != ' ' && ch != '\t' && ch != '\r' && ch != '\n' && ch != 0xFEFF) return PG_ERR_ARG;
    }
    if (v.size() != (size_t)K * (size_t)r) return PG_ERR_ARG;
    std::vector<uint32_t> w((size_t)K, 0u);
    for (int i = 0; i < K; i++)
        for (int b = 0; b < r; b++)
            if (v[(size_t)i * r + b]) w[i] |= 1u << b;
    const uint64_t low = w[0], mask = (r == 32) ? 0xFFFFFFFFull : ((1ull << r) - 1ull);
    if (!(low & 1ull)) return PG_ERR_ARG;  // g(D) must contain the constant term
    uint64_t cur = low;
    for (int i = 1; i < K; i++) {
        cur <<= 1;
        if ((cur >> r) & 1ull) cur = (cur & mask) ^ low;
        if ((uint32_t)cur != w[i]) return PG_ERR_ARG;
    }
    *crc_poly = low | (1ull << r);
    if (rows) std::memcpy(rows, w.data(), sizeof(uint32_t) * (size_t)K);
    return PG_OK;
}
