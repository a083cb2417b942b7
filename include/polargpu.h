/* polargpu.h -- C ABI of the B200-native batched polar-code decoding engine (libpolargpu.so).
 *
 * The reference (CHEBSB/PolarDecoding) has no library or FFI layer: each SC_x/SCL_x/CASCL_x/BP_x .c file is a
 * program whose main() builds a frame, calls ONE decoder function on it and counts errors.  The
 * entry points below are what a binding for that hot path would bind; each cites what it replaces:
 *
 *   pg_decode_llr        void SCdecode(double *y,int *u_hat)        /root/reference/SC_128.c:395  (SC_1024.c:434, SC_128_fag.c:410)
 *                        void SCLdecode(double *y,int *u_hat)       /root/reference/SCL_1024.c:547 (SCL_128.c:508, SCL_128_fag.c:526)
 *                        void CASCL(double *y,int *u_hat)           /root/reference/CASCL_1024_L8.c:601 (CASCL_128.c:539, CASCL_1024_sys.c:1125)
 *                        void BP(double *y,int *u_hat)              /root/reference/BP_1024.c:372 (BP_128.c:334, BP_128_fag.c:349)
 *                        void BPr(double *y,int *u_hat,int *u)      /root/reference/BPr_128.c:373 (pg_bpr_* below)
 *                        -- batched: B frames per call instead of one, LLRs instead of (y, global std)
 *   pg_channel           the encode + channel block of main()       /root/reference/SC_128.c:171-202, CASCL_1024_L8.c:245-292
 *                        (Ranq1/normal, SC_128.c:236-267, replaced by counter-based Philox4x32-10)
 *   pg_simulate          the Monte-Carlo loop of main()             /root/reference/SC_128.c:164-222, CASCL_1024_L8.c:234-312
 *   pg_info_set          the information-set block of main()        /root/reference/SC_128.c:139-147
 *
 * Conventions: plain pointers and sizes, caller owns every buffer, the context owns device scratch.
 * All functions return 0 on success or a negative pg_status; pg_last_error() gives the text.
 * There is NO CPU fallback: without a usable CUDA device pg_create() fails with PG_ERR_NO_DEVICE.
 * One context = one GPU + one stream; calls on one context must be serialised by the caller
 * (like the reference's decoders, which use process-global state, a context is not re-entrant).
 *
 * Bit order: position p of a frame is code bit / graph row p of the reference's Lee graph
 * (G = F^{(x)n}, no bit reversal).  The Kao-graph ("_fag") programs differ from the Lee programs only
 * by an internal relabelling, so they map to the same calls.
 * Packed bit vectors: bit p of a frame is bit (p & 31) of 32-bit word (p >> 5). */
#ifndef POLARGPU_H
#define POLARGPU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pg_ctx pg_ctx;

enum pg_status {
    PG_OK = 0,
    PG_ERR_ARG = -1,        /* bad parameter */
    PG_ERR_NO_DEVICE = -2,  /* no CUDA device / wrong architecture (sm_100 required) */
    PG_ERR_CUDA = -3,       /* CUDA runtime error, see pg_last_error */
    PG_ERR_NCCL = -4,       /* NCCL error */
    PG_ERR_UNSUPPORTED = -5 /* parameter combination without a compiled kernel */
};

enum pg_decoder { PG_DEC_SC = 0, PG_DEC_SCL = 1, PG_DEC_CASCL = 2, PG_DEC_BP = 3 };
enum pg_real { PG_REAL_F64 = 0, /* parity mode: bit-exact with the reference's double arithmetic */
               PG_REAL_F32 = 1, /* throughput mode */
               PG_REAL_H2 = 2   /* BP only, optional: messages as packed half (two frames per __half2); a numerically
                                   different decoder, judged on FER only; LLR buffers stay float */ };
enum pg_llr_format { PG_LLR_F32 = 0, PG_LLR_F64 = 1, PG_LLR_F16 = 2 }; /* the `llr_is_f64` argument of the streaming calls */
enum pg_data { PG_DATA_PN63 = 0,  /* the reference's PN-63 payload, phase m = frame*(K%63) mod 63 (SC_128.c:126-138,180,214) */
               PG_DATA_PHILOX = 1 /* random payload from the Philox stream */ };

typedef struct pg_params {
    int N;                 /* block length: 128 or 1024 (any power of two 32..1024) */
    int K;                 /* payload bits */
    int crc_bits;          /* r: 0, 6 (CASCL_128.c:18) or 24 (CASCL_1024_L8.c:19); <= 32 */
    uint64_t crc_poly;     /* bit e set <=> D^e in g(D), incl. D^r and 1; see PG_CRC24_POLY / PG_CRC6_POLY */
    int crc_systematic;    /* 0: w = v*g (CASCL_1024_L8.c:251-266); 1: parity-first systematic (CASCL_1024_sys.c:776-789) */
    int decoder;           /* enum pg_decoder */
    int list_size;         /* L: 1 (SC) or 2,4,8,16,32 */
    int iter_max;          /* BP sweeps: 100 (BP_1024.c:16), 1..255 (the per-frame word holds the executed sweeps in 8 bits); ignored otherwise */
    int bp_early_stop;     /* bit 0: stop a frame once a sweep leaves every message bit-identical (decisions unchanged, parity-safe);
                              bit 1: also stop when the hard decisions form a codeword (u G = x, "G-matrix" rule; not in the
                              reference, FER-level equivalence only; not available with PG_REAL_H2) */
    int real;              /* enum pg_real */
    int data_mode;         /* enum pg_data */
    int count_from;        /* first index of I[] that enters the error count: r for CASCL_1024_sys.c:821, else 0 */
    int device;            /* CUDA device ordinal */
    uint64_t seed;         /* Philox key */
    int rank, nranks;      /* frame-space partition for pg_simulate (see there) */
    float llr_clip;        /* > 0: every channel LLR is clipped to [-llr_clip, +llr_clip] before decoding (optional receiver model,
                              SURVEY 8f.3: what a fixed-point front end does; not in the reference, so 0 = off in every preset).
                              Decisions equal the reference decoder's on the clipped LLRs (tests/test_gpu_parity.py) */
} pg_params;

#define PG_CRC24_POLY 0x1B2B117ull /* D^24+D^23+D^21+D^20+D^17+D^15+D^13+D^12+D^8+D^4+D^2+D+1 */
#define PG_CRC6_POLY 0x61ull       /* D^6+D^5+1 */

/* counters of one Eb/N0 point; all u64 so that multi-rank sums cannot overflow */
typedef struct pg_counters {
    uint64_t frames;       /* "run" */
    uint64_t err_blocks;   /* "error block" */
    uint64_t err_bits;     /* "Error bit" */
    uint64_t tie_frames;   /* list decoders: frames in which an exact PM tie straddled the list boundary ("Oops!", SCL_1024.c:621) */
    uint64_t crc_fail;     /* CA-SCL: frames in which no path passed the CRC */
    uint64_t bp_sweeps;    /* BP: sweeps executed, summed over frames */
    uint64_t reserved[2];
} pg_counters;

/* fill *p with the defaults of one of the reference programs: "SC_128", "SC_1024", "SC_128_fag", "SCL_128",
 * "SCL_1024", "SCL_128_fag", "CASCL_128", "CASCL_1024_L8", "CASCL_1024_sys", "BP_128", "BP_1024", "BP_128_fag", "BPr_128";
 * plus "CASCL_128_sys": the systematic CRC-6 variant of CASCL_128 whose parity table is the reference's CRC_6.dat and whose
 * result files are result_128_fag/CAL8_0.dat / CAL32_0.dat (its source is not in the reference repository) */
int pg_params_preset(pg_params *p, const char *program);

int pg_create(const pg_params *p, pg_ctx **out);
void pg_destroy(pg_ctx *ctx);
const char *pg_last_error(const pg_ctx *ctx); /* ctx may be NULL: error of the last failed pg_create on this thread */

/* information(+CRC) positions in reliability order I[0..K+r) and the membership mask inI[0..N) (either may be NULL) */
int pg_info_set(const pg_ctx *ctx, int *I_out, uint8_t *inI_out);

/* ---- streaming mode: decode B frames of caller-supplied LLRs -------------------------------------------
 * llr      [B][N] row-major, natural bit order; element type by llr_is_f64 = one of PG_LLR_F32 (0), PG_LLR_F64 (1) or
 *          PG_LLR_F16 (2: IEEE binary16, 2 KiB per N = 1024 frame -- half the bytes a streaming host has to move; a quantised
 *          input, so judged on FER like PG_REAL_H2, not bit-exact); converted to the context's arithmetic type on the
 *          device.  HOST pointer (pinned or pageable); copied H2D inside the call.
 * u_hat    [B][N] one byte per bit (0/1), HOST, may be NULL.
 * flags    [B] per-frame: bit0 tie frame, bit1 no CRC pass; bits 8..15 BP sweeps executed. HOST, may be NULL. */
int pg_decode_llr(pg_ctx *ctx, const void *llr, int llr_is_f64, size_t B, uint8_t *u_hat, uint32_t *flags);
/* as pg_decode_llr, but the decisions come back packed ([B][N/32] words, HOST): 32x less device-to-host traffic */
int pg_decode_llr_packed(pg_ctx *ctx, const void *llr, int llr_is_f64, size_t B, uint32_t *u_hat_packed, uint32_t *flags);
/* decode with the truth supplied by the caller -- the shape of BPr(double *y, int *u_hat, int *u) (BPr_128.c:373): u_true is
 * [B][N] bytes (HOST); wrong counted bits per frame go to frame_err (may be NULL), counters are ADDED to *acc, u_hat may be
 * NULL.  With pg_bpr_config active and a BP context this also accumulates the BPR statistic. */
int pg_decode_llr_counted(pg_ctx *ctx, const void *llr, int llr_is_f64, size_t B, const uint8_t *u_true, uint8_t *u_hat,
                          pg_counters *acc, uint16_t *frame_err);
/* same with DEVICE pointers and packed output (u_hat_packed: [B][N/32] words); no copies, asynchronous on the ctx stream.
 * d_frame_info ([B], DEVICE, may be NULL) receives the kernels' RAW per-frame word (layout: PG_INFO_* below), not the
 * host calls' compact `flags`: the call launches nothing but the decode kernel.  PG_INFO_TO_FLAGS(w) converts one word. */
int pg_decode_llr_device(pg_ctx *ctx, const void *d_llr, int llr_is_f64, size_t B, uint32_t *d_u_hat_packed, uint32_t *d_frame_info);

/* per-frame word written by the decode kernels (d_frame_info of the *_device calls, frame_info of pg_truncate_info) */
#define PG_INFO_ERRBITS_MASK 0xFFFFu      /* bits 0..15: wrong counted bits (only when a truth vector was supplied), saturating */
#define PG_INFO_TIE (1u << 16)            /* list decoders: exact PM tie across the list boundary in this frame */
#define PG_INFO_CRC_FAIL (1u << 17)       /* CA-SCL: no path passed the CRC */
#define PG_INFO_SWEEPS_SHIFT 24           /* bits 24..31: BP sweeps executed (iter_max <= 255) */
/* the compact `flags` word of pg_decode_llr / pg_decode_llr_packed: bit0 tie, bit1 no CRC pass, bits 8..15 BP sweeps */
#define PG_INFO_TO_FLAGS(w) ((((w) >> 16) & 3u) | (((w) >> PG_INFO_SWEEPS_SHIFT) << 8))

/* device-resident variant with the on-device error count: d_truth_packed ([B][N/32], e.g. from pg_channel_device) is
 * compared on the counted positions; block/bit errors, tie and CRC-fail frames are ADDED to the context's device
 * counters (read with pg_counters_read; no other call touches them).  d_u_hat_packed / d_frame_info may be NULL.
 * Asynchronous on the ctx stream. */
int pg_decode_count_device(pg_ctx *ctx, const void *d_llr, int llr_is_f64, size_t B, const uint32_t *d_truth_packed,
                           uint32_t *d_u_hat_packed, uint32_t *d_frame_info);
int pg_counters_read(pg_ctx *ctx, pg_counters *out, int reset); /* synchronises the ctx stream */

/* ---- fused channel: payload, CRC, polar encode, BPSK, AWGN, LLR for frames [first_frame, first_frame+B) ----
 * Noise comes from Philox4x32-10 with key = seed and counter = (global frame index, position/4): the result for a
 * frame does not depend on B, on the batch split or on the rank that produces it.
 * llr_out [B][N] (context's arithmetic type: float or double), u_out [B][N] bytes; HOST pointers, may be NULL. */
int pg_channel(pg_ctx *ctx, double ebn0_db, uint64_t first_frame, size_t B, void *llr_out, uint8_t *u_out);

/* same into DEVICE buffers (d_llr: [B][N] of the context's type; d_u_packed: [B][N/32] words; either may be NULL) */
int pg_channel_device(pg_ctx *ctx, double ebn0_db, uint64_t first_frame, size_t B, void *d_llr, uint32_t *d_u_packed);

/* ---- Monte-Carlo point: channel + decode + on-device error count -------------------------------------
 * Simulates frames first_frame + rank*chunk + i ... in rounds of nranks*chunk frames (chunk = frames one rank
 * processes per round) until, over all ranks, at least target_err_blocks block errors were seen or max_frames
 * frames were run (either limit may be 0 = unlimited, not both).  With exact_stop != 0 the counters are truncated at
 * the frame, in global frame order, on which the target-th block error fell -- the reference's stopping rule
 * (SC_128.c:169) -- otherwise whole rounds are counted.  With nranks > 1 every rank must make the same call; the
 * per-round counters are combined with one NCCL all-reduce (pg_comm_init first).  out = global counters. */
int pg_simulate(pg_ctx *ctx, double ebn0_db, uint64_t first_frame, uint64_t target_err_blocks, uint64_t max_frames,
                int exact_stop, pg_counters *out);
/* The loop is pipelined: the next round is already queued on the GPU while the host (and NCCL, on a second stream) finish the
 * previous one; with a frame budget only (target_err_blocks == 0) the counters accumulate on the device and the ranks are combined
 * by ONE all-reduce at the end.  pg_simulate_stats: rounds launched and all-reduces issued by the last pg_simulate call. */
int pg_simulate_stats(const pg_ctx *ctx, uint64_t *rounds, uint64_t *allreduces);

/* fixed-size batch, no stopping rule, no collective: frames [first_frame, first_frame+B) on this GPU; counters ADDED to *acc.
 * frame_err (HOST, [B], may be NULL): per frame, number of wrong counted bits (0 = frame correct).  Used by bench and tests. */
int pg_simulate_batch(pg_ctx *ctx, double ebn0_db, uint64_t first_frame, size_t B, pg_counters *acc, uint16_t *frame_err);

/* ---- BPR statistic (BPr_128.c:418-568): per-stage hard decision + re-encode error counts -------------------
 * Enables sampling at the given 1-based sweep counts (ns <= 8) for subsequent BP calls; E (HOST, ns x (n+1) u64) is read
 * back with pg_bpr_read (sums since the last pg_bpr_reset). */
int pg_bpr_config(pg_ctx *ctx, const int *sample_sweeps, int ns);
int pg_bpr_read(pg_ctx *ctx, uint64_t *E);
int pg_bpr_reset(pg_ctx *ctx);

/* ---- multi-GPU: one process (or thread) per GPU; only counters are exchanged ---------------------------------
 * pg_comm_unique_id fills 128 bytes (ncclUniqueId) on one rank; the caller distributes them (MPI, torch.distributed,
 * a file ...) and every rank calls pg_comm_init(ctx, id).  rank/nranks come from pg_params. */
int pg_comm_unique_id(void *id128);
int pg_comm_init(pg_ctx *ctx, const void *id128);
int pg_allreduce_counters(pg_ctx *ctx, pg_counters *c); /* in place, sum over ranks */

/* ---- the host-side rules pg_simulate applies, exported so that other hosts and CPU tests can use them --------
 * pg_partition: frames of `rank` in the round that starts at global frame round_first: rank q owns
 *   [round_first + q*chunk, +chunk), clipped to `budget` frames left in the run (count may be 0).
 * pg_merge_round: walk the ranks' counters of one round in global frame order, adding them to *acc.  If exact_stop and
 *   the target-th block error falls into rank q's chunk: stop BEFORE adding q, return *cut_rank = q and *need = block
 *   errors still missing; rank q then calls pg_truncate_info on its per-frame results and every rank adds that part.
 * pg_truncate_info: counters of the shortest prefix of frame_info[] that holds `need` block errors (frame_info words
 *   as written by the kernels: bits 0..15 wrong bits, bit 16 tie, bit 17 CRC fail, bits 24..31 BP sweeps). */
int pg_partition(uint64_t round_first, uint64_t chunk, int nranks, int rank, uint64_t budget, uint64_t *start, uint64_t *count);
int pg_merge_round(const pg_counters *round, int nranks, uint64_t target, int exact_stop, pg_counters *acc, int *cut_rank, uint64_t *need);
int pg_truncate_info(const uint32_t *frame_info, size_t nframes, uint64_t need, pg_counters *part);

/* ---- CRC parity table in the reference's file format (/root/reference/CRC_6.dat: K rows of r integers, row i = coefficients
 * c0..c(r-1) of D^(r+i) mod g(D); UTF-16 with BOM or ASCII).  Validates that the file is the table of ONE polynomial, returns it
 * in the pg_params.crc_poly convention and, if rows != NULL, the K parity rows as r-bit words (bit b = coefficient of D^b).
 * Host only, no device needed.  The systematic encoder/CRC check (crc_systematic = 1) derives the same rows from crc_poly. */
int pg_crc_table_load(const char *path, int K, int r, uint64_t *crc_poly, uint32_t *rows);

/* ---- introspection for benches ------------------------------------------------------------------------ */
/* frames one full grid of the decode kernel holds at a time (resident CTAs x frames per CTA): batch sizes that are a
 * multiple of this keep every SM busy until the end of a launch */
uint64_t pg_wave_frames(const pg_ctx *ctx);
int pg_sync(pg_ctx *ctx);                              /* cudaStreamSynchronize on the ctx stream */
void *pg_stream(pg_ctx *ctx);                          /* the cudaStream_t the kernels are launched on */
uint64_t pg_kernel_launches(const pg_ctx *ctx);        /* kernels launched by this context so far */
/* device time (ms, CUDA events on the ctx stream) of the last decode kernel and of the last channel kernel */
int pg_last_kernel_ms(pg_ctx *ctx, float *decode_ms, float *channel_ms);
const char *pg_version(void);
int pg_device_count(void);                             /* usable (sm_100) devices 0..n-1; 0 without a driver or device */

#ifdef __cplusplus
}
#endif
#endif
