/* Drop-in GPU counterparts of the reference simulators (one binary per reference program).
 *
 * Each reference file (/root/reference/SC_128.c, SCL_1024.c, CASCL_1024_L8.c, BP_1024.c, ...) is a
 * program: compile-time N/K/L/BLE/iterMax, an Eb/N0 sweep, Fn on stdin, result lines on stdout.
 * This host program keeps that contract -- same defaults per program, same stdin (validated if
 * present, optional), byte-identical result-line formats -- and replaces the per-frame work
 * (frame construction, decoder call, error count; e.g. CASCL_1024_L8.c:239-306) by calls into
 * libpolargpu.so (include/polargpu.h).  Host code is plain C; there is no CPU decoding path.
 *
 * The program identity comes from -DPOLAR_PROGRAM="..." (host/Makefile builds one binary per
 * reference program), from --program NAME, or from argv[0].
 *
 *   --ebn0 a:step:b | --ebn0 x      sweep (default: the reference file's for-loop bounds)
 *   --ble n                         block errors to stop at (default: the reference's)
 *   --max-frames n                  frame budget per point instead of / in addition to --ble
 *   --seed s                        default: the reference's rule (const 1024, or time()%10000 / %1000)
 *   --rng philox|ref                philox (default): counter-based channel fused on the GPU;
 *                                   ref: the reference's Ranq1 + polar-method generator on the host, in the
 *                                   reference's order, so that run/error columns reproduce its captures exactly
 *   --real f32|f64                  arithmetic (default f32 with philox, f64 with ref)
 *   --L n  --iters n  --early-stop  list size / BP sweeps / bit-exact fixed-point stop
 *   --gmatrix-stop                  BP: also stop when the decisions form a codeword (not in the reference; same FER, fewer sweeps)
 *   --gpus n                        partition the frame space over n GPUs (one forked process + one ctx per GPU, NCCL counters)
 *   --verbose                       throughput and tie/CRC statistics on stderr (stdout stays drop-in)
 */
#define _POSIX_C_SOURCE 200809L
#include <math.h>
#include <poll.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <time.h>
#include <unistd.h>

#include "../../include/polargpu.h"

#ifndef POLAR_PROGRAM
#define POLAR_PROGRAM ""
#endif

typedef struct prog_def {
    const char *name;
    double e0, e1;       /* sweep bounds of the reference's for loop, step 0.5 */
    int ble;             /* error-block target */
    int seed_mod;        /* 0: const SEED = 1024; else time()%seed_mod */
    int fmt;             /* output format id */
} prog_def;

enum { F_SC, F_SCL_1E1, F_SCL_1E2, F_SCL_FAG, F_CASCL_1E3, F_CASCL_1E4, F_CASCL_SYS, F_CASCL_128_SYS, F_BP_128, F_BP_1024, F_BPR };

static const prog_def PROGS[] = {
    {"SC_128", 1.0, 4.0, 100, 0, F_SC},           /* SC_128.c:164,169,218-221 */
    {"SC_1024", 1.0, 3.5, 50, 0, F_SC},           /* SC_1024.c:203,208,257-260 */
    {"SC_128_fag", 1.0, 4.0, 500, 0, F_SC},       /* SC_128_fag.c:181,186 */
    {"SCL_128", 1.0, 2.5, 50, 0, F_SCL_1E1},      /* SCL_128.c:184,189,238 */
    {"SCL_1024", 1.0, 3.0, 50, 0, F_SCL_1E2},     /* SCL_1024.c:223,228,277 */
    {"SCL_128_fag", 1.0, 3.0, 200, 10000, F_SCL_FAG}, /* SCL_128_fag.c:121,205,256-259 */
    {"CASCL_128", 1.0, 3.0, 200, 10000, F_CASCL_1E3},   /* CASCL_128.c:124,194,257 */
    {"CASCL_1024_L8", 1.0, 1.5, 200, 10000, F_CASCL_1E4}, /* CASCL_1024_L8.c:164,234,308 */
    {"CASCL_1024_sys", 2.5, 2.5, 200, 10000, F_CASCL_SYS}, /* CASCL_1024_sys.c:681,765,832-835 */
    {"CASCL_128_sys", 1.0, 3.5, 200, 10000, F_CASCL_128_SYS}, /* systematic CRC-6 (CRC_6.dat); sweep and line format of result_128_fag/CAL8_0.dat */
    {"BP_128", 1.0, 4.0, 200, 1000, F_BP_128},    /* BP_128.c:96,163,217 */
    {"BP_1024", 1.0, 3.5, 200, 1000, F_BP_1024},  /* BP_1024.c:136,203,255-257 */
    {"BP_128_fag", 1.0, 4.0, 200, 1000, F_BP_128}, /* BP_128_fag.c:98,179 */
    {"BPr_128", 1.0, 4.0, 200, 1000, F_BPR},      /* BPr_128.c:105,171,227-258 */
};

static void die(const char *msg, pg_ctx *ctx)
{
    fprintf(stderr, "polar_sim: %s: %s\n", msg, pg_last_error(ctx));
    exit(2);
}

/* ---- the reference's generator (Ranq1 + Marsaglia polar, SC_128.c:236-267), host side of --rng ref ---- */
typedef struct { uint64_t v; } ranq1;
static void ranq1_seed(ranq1 *g, uint64_t seed)
{
    uint64_t v = seed ^ 4101842887655102017ULL;
    v ^= v >> 21; v ^= v << 35; v ^= v >> 4;
    g->v = v * 2685821657736338717ULL;
}
static double ranq1_next(ranq1 *g)
{
    uint64_t v = g->v;
    v ^= v >> 21; v ^= v << 35; v ^= v >> 4;
    g->v = v;
    return (double)(v * 2685821657736338717ULL) * 5.42101086242752217E-20;
}
static void normal_pair(ranq1 *g, double sd, double *n1, double *n2)
{
    double x1, x2, s;
    do {
        x1 = 2 * ranq1_next(g) - 1;
        x2 = 2 * ranq1_next(g) - 1;
        s = x1 * x1 + x2 * x2;
    } while (s >= 1.0);
    *n1 = sd * x1 * sqrt(-2 * log(s) / s);
    *n2 = sd * x2 * sqrt(-2 * log(s) / s);
}

/* stdin: the reference reads N*N ints of Fn (SC_128.c:149-158).  Accept it, check it, do not need it. */
static void consume_stdin(int N)
{
    if (isatty(STDIN_FILENO)) return;
    {   /* nothing waiting on stdin (an idle pipe, a closed descriptor): the matrix is optional here, do not block */
        struct pollfd pf = {STDIN_FILENO, POLLIN, 0};
        if (poll(&pf, 1, 0) <= 0 || !(pf.revents & (POLLIN | POLLHUP))) return;
    }
    long cnt = 0, bad = 0, wrong = 0;
    int v;
    while (cnt < (long)N * N && scanf("%d", &v) == 1) {
        if (v != 0 && v != 1) { printf("Illegal input!\n"); bad++; }
        else {
            const long i = cnt / N, j = cnt % N;
            if (v != (((i & j) == j) ? 1 : 0)) wrong++;
        }
        cnt++;
    }
    if (cnt == (long)N * N && (wrong || bad))
        fprintf(stderr, "polar_sim: note: the matrix on stdin is not F^{(x)n}; the GPU encoder uses F^{(x)n} (as every reference run does)\n");
}

static const int BPR_SAMPLES[6] = {3, 6, 10, 20, 40, 80}; /* i0..i5, BPr_128.c:18-23 */
static uint64_t g_bprE[6 * 16];                             /* E[sample][stage] of the current Eb/N0 point */

static void print_point(const prog_def *pd, const pg_params *p, double snr, const pg_counters *c)
{
    const int errBlock = (int)c->err_blocks, run = (int)c->frames, errbit = (int)c->err_bits, K = p->K, L = p->list_size;
    switch (pd->fmt) {
    case F_SC:
        printf("bSNR = %.2lf\terror block = %d\trun = %d\tBLER = %lf\n", snr, errBlock, run, ((double)errBlock) / run);
        printf("Error bit = %d\tBER = %lf\n", errbit, ((double)errbit) / K / run);
        break;
    case F_SCL_1E1:
        printf("L = %d\tbSNR = %.2lf\terror block = %d\trun = %d\tBLER = %lf * 10^-1\n", L, snr, errBlock, run, ((double)errBlock) * 10 / run);
        break;
    case F_SCL_1E2:
        printf("L = %d\tbSNR = %.2lf\terror block = %d\trun = %d\tBLER = %lfe-2\n", L, snr, errBlock, run, ((double)errBlock) * 100 / run);
        break;
    case F_SCL_FAG:
        printf("L = %d\tbSNR = %.2lf\terror block = %d\trun = %d\tBLER = %lf * 10^-1\n", L, snr, errBlock, run, ((double)errBlock) * 10 / run);
        printf("Error bit = %d\tBER = %lf\n", errbit, ((double)errbit) / K / run);
        break;
    case F_CASCL_1E3:
        printf("L = %d\tbSNR = %.2lf\terror block = %d\trun = %d\tBLER = %lfe-3\n", L, snr, errBlock, run, ((double)errBlock) / (run / 1000.0));
        break;
    case F_CASCL_1E4:
        printf("L = %d\tbSNR = %.2lf\terror block = %d\trun = %d\tBLER = %lfe-4\n", L, snr, errBlock, run, ((double)errBlock) / (run / 10000.0));
        break;
    case F_CASCL_SYS:
        printf("bSNR = %.2lf\trun = %d\tBLER = %lfe-3\t", snr, run, ((double)errBlock) / (run / 1000.0));
        printf("Error bit = %d\tBER = %lfe-3\n", errbit, ((double)errbit) / (K) / (run / 1000.0));
        break;
    case F_CASCL_128_SYS: /* result_128_fag/CAL8_0.dat */
        printf("L = %d\tbSNR = %.2lf\terror block = %d\trun = %d\tBLER = %lfe-3\n", L, snr, errBlock, run, ((double)errBlock) / (run / 1000.0));
        printf("Error bit = %d\tBER = %lfe-3\n", errbit, ((double)errbit) / (K) / (run / 1000.0));
        break;
    case F_BP_128:
        printf("bSNR = %.2lf\terror block = %d\trun = %d\t", snr, errBlock, run);
        printf("BLER = %lfe-2\tBER = %lfe-2\n", ((double)errBlock) * 100 / run, ((double)errbit) * 100 / K / run);
        break;
    case F_BP_1024:
        printf("bSNR = %.2lf\terror block = %d\trun = %d\t", snr, errBlock, run);
        printf("BLER = %lf * 10^-3\n", ((double)errBlock) * 1000 / run);
        break;
    case F_BPR: { /* BPr_128.c:227-258 */
        int nst = 0;
        while ((1 << nst) < p->N) nst++;
        printf("bSNR = %.2lf\terror block = %d\trun = %d\t", snr, errBlock, run);
        for (int t = 0; t < 6; t++) {
            printf(t == 0 ? "\nAfter %2d iterations:\n" : "After %2d iterations:\n", BPR_SAMPLES[t]);
            for (int i = 0; i <= nst; i++) printf("%lf\t", ((double)g_bprE[t * (nst + 1) + i]) / run);
            printf("\n");
        }
        printf("BLER = %lfe-2\tBER = %lfe-2\tK * BER = %lf\n", ((double)errBlock) * 100 / run, ((double)errbit) * 100 / K / run, ((double)errbit) / run);
        break;
    }
    }
    fflush(stdout);
}

/* ---- --rng ref: frames built on the host exactly as the reference's main() does, decoded on the GPU ---- */
typedef struct { ranq1 g; int m; } ref_state;

static void ref_point(pg_ctx *ctx, const pg_params *p, double snr, uint64_t ble, uint64_t max_frames, ref_state *st, pg_counters *out, int bpr)
{
    const int N = p->N, K = p->K, r = p->crc_bits, nI = K + r;
    int *I = (int *)malloc(sizeof(int) * (size_t)nI);
    int PN[63], reg[6] = {0, 0, 0, 0, 0, 0};
    const double sd = pow(10, snr / ((double)-20)); /* SC_128.c:167 */
    const size_t CH = 512;
    double *llr = (double *)malloc(sizeof(double) * CH * (size_t)N);
    uint8_t *u = (uint8_t *)calloc(CH * (size_t)N, 1), *uh = (uint8_t *)malloc(CH * (size_t)N);
    uint32_t *fl = (uint32_t *)malloc(sizeof(uint32_t) * CH);
    ref_state *after = (ref_state *)malloc(sizeof(ref_state) * CH);
    int *w = (int *)malloc(sizeof(int) * (size_t)(nI + 64)), *x = (int *)malloc(sizeof(int) * (size_t)N);
    if (pg_info_set(ctx, I, NULL)) die("pg_info_set", ctx);
    for (int i = 0; i < 63; i++) { /* SC_128.c:126-138 */
        const int b = (i == 0) ? 1 : (i < 6) ? 0 : (reg[4] ^ reg[5]);
        PN[i] = b;
        reg[5] = reg[4]; reg[4] = reg[3]; reg[3] = reg[2]; reg[2] = reg[1]; reg[1] = reg[0]; reg[0] = b;
    }
    memset(out, 0, sizeof(*out));
    int done = 0;
    while (!done) {
        size_t nb = CH;
        if (max_frames && max_frames - out->frames < nb) nb = (size_t)(max_frames - out->frames);
        if (nb == 0) break;
        for (size_t f = 0; f < nb; f++) {
            uint8_t *uf = u + f * (size_t)N;
            memset(uf, 0, (size_t)N);
            memset(w, 0, sizeof(int) * (size_t)(nI + 64));
            if (r == 0) {
                for (int i = 0; i < K; i++) w[i] = PN[(st->m + i) % 63];
            } else if (!p->crc_systematic) { /* CASCL_1024_L8.c:251-266 */
                for (int i = 0; i < K; i++)
                    if (PN[(st->m + i) % 63])
                        for (int e = 0; e <= r; e++)
                            if ((p->crc_poly >> e) & 1) w[i + e] ^= 1;
            } else { /* CASCL_1024_sys.c:778-789: parity = v(D) D^r mod g(D) */
                int *rem = (int *)calloc((size_t)(nI + 64), sizeof(int));
                for (int i = 0; i < K; i++) { w[r + i] = PN[(st->m + i) % 63]; rem[r + i] = w[r + i]; }
                for (int i = nI - 1; i >= r; i--)
                    if (rem[i])
                        for (int e = 0; e <= r; e++)
                            if ((p->crc_poly >> e) & 1) rem[i - r + e] ^= 1;
                for (int i = 0; i < r; i++) w[i] = rem[i];
                free(rem);
            }
            for (int i = 0; i < nI; i++) uf[I[i]] = (uint8_t)w[i];
            for (int i = 0; i < N; i++) x[i] = uf[i];
            for (int d = 1; d < N; d <<= 1) /* x = u F^{(x)n}: what the row-XOR of Fn computes (SC_128.c:183-191) */
                for (int j = 0; j < N; j++)
                    if (!(j & d)) x[j] ^= x[j + d];
            for (int i = 0; i < N; i += 2) { /* SC_128.c:194-202, LLR as in SC_128.c:418 */
                double n1, n2;
                normal_pair(&st->g, sd, &n1, &n2);
                const double y0 = (x[i] == 0) ? 1 + n1 : -1 + n1, y1 = (x[i + 1] == 0) ? 1 + n2 : -1 + n2;
                llr[f * (size_t)N + i] = 2 * y0 / sd / sd;
                llr[f * (size_t)N + i + 1] = 2 * y1 / sd / sd;
            }
            st->m += K % 63;
            if (st->m >= 63) st->m -= 63;
            after[f] = *st;
        }
        if (bpr) { /* BPr(y, u_hat, u): the truth goes in with the frame (BPr_128.c:213) */
            pg_counters scratch;
            memset(&scratch, 0, sizeof(scratch));
            memset(fl, 0, sizeof(uint32_t) * nb);
            if (pg_bpr_reset(ctx) || pg_decode_llr_counted(ctx, llr, 1, nb, u, uh, &scratch, NULL)) die("pg_decode_llr_counted", ctx);
        } else if (pg_decode_llr(ctx, llr, 1, nb, uh, fl)) die("pg_decode_llr", ctx);
        size_t used = nb;
        for (size_t f = 0; f < nb; f++) {
            int bad = 0;
            for (int i = p->count_from; i < nI; i++)
                if (u[f * (size_t)N + I[i]] != uh[f * (size_t)N + I[i]]) { bad = 1; out->err_bits++; }
            out->err_blocks += (uint64_t)bad;
            out->frames++;
            out->tie_frames += fl[f] & 1u;
            out->crc_fail += (fl[f] >> 1) & 1u;
            if (ble && out->err_blocks >= ble) { *st = after[f]; done = 1; used = f + 1; break; } /* the reference stops here (SC_128.c:169) */
        }
        if (bpr) { /* the statistic covers exactly the frames the sequential program would have run */
            uint64_t E[6 * 16];
            pg_counters scratch;
            memset(&scratch, 0, sizeof(scratch));
            if (used < nb && (pg_bpr_reset(ctx) || pg_decode_llr_counted(ctx, llr, 1, used, u, NULL, &scratch, NULL))) die("pg_decode_llr_counted", ctx);
            if (pg_bpr_read(ctx, E)) die("pg_bpr_read", ctx);
            for (int i = 0; i < 6 * 16; i++) g_bprE[i] += E[i];
        }
        if (max_frames && out->frames >= max_frames) done = 1;
    }
    free(I); free(llr); free(u); free(uh); free(fl); free(after); free(w); free(x);
}

/* ---- --gpus n: one PROCESS per GPU (fork before any CUDA call); rank 0 creates the NCCL id and hands it to the
 * others through pipes; every rank makes the same pg_simulate calls, the library exchanges the counters, rank 0 prints ---- */
static int spawn_ranks(long gpus, unsigned char id[128], int *have_id)
{
    int (*fds)[2] = (int (*)[2])malloc(sizeof(int[2]) * (size_t)gpus);
    for (long g = 1; g < gpus; g++)
        if (pipe(fds[g])) { perror("pipe"); exit(3); }
    int rank = 0;
    for (long g = 1; g < gpus; g++) {
        const pid_t pid = fork();
        if (pid < 0) { perror("fork"); exit(3); }
        if (pid == 0) { rank = (int)g; break; }
    }
    if (rank == 0) {
        if (pg_comm_unique_id(id)) { fprintf(stderr, "polar_sim: %s\n", pg_last_error(NULL)); exit(3); }
        for (long g = 1; g < gpus; g++) {
            close(fds[g][0]);
            if (write(fds[g][1], id, 128) != 128) { perror("write"); exit(3); }
            close(fds[g][1]);
        }
    } else {
        close(fds[rank][1]);
        if (read(fds[rank][0], id, 128) != 128) { fprintf(stderr, "polar_sim: rank %d got no NCCL id\n", rank); exit(3); }
        close(fds[rank][0]);
        if (!freopen("/dev/null", "w", stdout)) exit(3);   /* only rank 0 prints the result lines */
    }
    *have_id = 1;
    free(fds);
    return rank;
}

int main(int argc, char **argv)
{
    const char *name = POLAR_PROGRAM;
    double e0 = -1, e1 = -1, estep = 0.5;
    long ble = -1, L = -1, iters = -1, gpus = 1;
    long long seed = -1, maxf = 0;
    int use_ref = 0, real = -1, early = 0, verbose = 0;
    const char *crc_file = NULL;
    double llr_clip = 0;
    if (!name[0]) { const char *b = strrchr(argv[0], '/'); name = b ? b + 1 : argv[0]; }
    for (int i = 1; i < argc; i++) {
        const char *a = argv[i], *v = (i + 1 < argc) ? argv[i + 1] : NULL;
        if (!strcmp(a, "--program") && v) { name = v; i++; }
        else if (!strcmp(a, "--ebn0") && v) {
            if (sscanf(v, "%lf:%lf:%lf", &e0, &estep, &e1) != 3) { e0 = e1 = atof(v); estep = 0.5; }
            i++;
        }
        else if (!strcmp(a, "--ble") && v) { ble = atol(v); i++; }
        else if (!strcmp(a, "--max-frames") && v) { maxf = atoll(v); i++; }
        else if (!strcmp(a, "--seed") && v) { seed = atoll(v); i++; }
        else if (!strcmp(a, "--rng") && v) { use_ref = !strcmp(v, "ref"); i++; }
        else if (!strcmp(a, "--real") && v) { real = !strcmp(v, "f64") ? PG_REAL_F64 : (!strcmp(v, "h2") ? PG_REAL_H2 : PG_REAL_F32); i++; }
        else if (!strcmp(a, "--L") && v) { L = atol(v); i++; }
        else if (!strcmp(a, "--iters") && v) { iters = atol(v); i++; }
        else if (!strcmp(a, "--gpus") && v) { gpus = atol(v); i++; }
        else if (!strcmp(a, "--crc-file") && v) { crc_file = v; i++; }
        else if (!strcmp(a, "--llr-clip") && v) { llr_clip = atof(v); i++; }
        else if (!strcmp(a, "--early-stop")) early |= 1;
        else if (!strcmp(a, "--gmatrix-stop")) early |= 2;
        else if (!strcmp(a, "--verbose")) verbose = 1;
        else { fprintf(stderr, "polar_sim: unknown option %s\n", a); return 2; }
    }
    const prog_def *pd = NULL;
    for (size_t i = 0; i < sizeof(PROGS) / sizeof(PROGS[0]); i++)
        if (!strcmp(PROGS[i].name, name)) pd = &PROGS[i];
    if (!pd) { fprintf(stderr, "polar_sim: unknown program '%s' (build one binary per program, or use --program)\n", name); return 2; }

    pg_params p;
    if (pg_params_preset(&p, pd->name)) { fprintf(stderr, "polar_sim: no preset for %s\n", pd->name); return 2; }
    if (L > 0) p.list_size = (int)L;
    if (iters > 0) p.iter_max = (int)iters;
    if (crc_file) { /* the CRC parity table in the reference's file format (CRC_6.dat): it determines g(D) */
        uint64_t poly = 0;
        if (p.crc_bits == 0 || pg_crc_table_load(crc_file, p.K, p.crc_bits, &poly, NULL)) {
            fprintf(stderr, "polar_sim: %s is not a %d x %d CRC parity table\n", crc_file, p.K, p.crc_bits);
            return 2;
        }
        p.crc_poly = poly;
    }
    p.bp_early_stop = early;
    p.llr_clip = (float)llr_clip;
    p.real = (real >= 0) ? real : (use_ref ? PG_REAL_F64 : PG_REAL_F32);
    if (e0 < 0) { e0 = pd->e0; e1 = pd->e1; }
    if (ble < 0) ble = maxf ? 0 : pd->ble;
    if (gpus < 1 || (use_ref && gpus != 1)) { fprintf(stderr, "polar_sim: --rng ref is sequential by construction: one GPU\n"); return 2; }

    /* the reference's seeding rule (e.g. CASCL_1024_L8.c:164-165; const SEED = 1024 in SC_128.c:35) */
    unsigned long long SEED = (seed >= 0) ? (unsigned long long)seed
                              : pd->seed_mod ? ((unsigned long long)time(NULL)) % (unsigned long long)pd->seed_mod : 1024ULL;
    p.seed = SEED;

    if (!strcmp(pd->name, "SC_128_fag")) { /* its debug line, SC_128_fag.c:150-152 */
        for (int i = 0; i < p.N; i = 3 * i + 5) {
            int rev = 0, t = i;
            for (int j = 6; j >= 0; j--) { if (t % 2 == 1) rev += 1 << j; t /= 2; }
            printf("bRev[%d] = %d\t", i, rev);
        }
        printf("\n");
    }
    if (pd->fmt == F_CASCL_SYS) printf("SEED = %ld\terror block = %d\tL = %d\n", (long)SEED, (int)ble, p.list_size); /* CASCL_1024_sys.c:682 */
    else if (pd->seed_mod) printf("SEED = %ld\n", (long)SEED);
    consume_stdin(p.N);

    if (gpus > 1 && pd->fmt == F_BPR) { fprintf(stderr, "polar_sim: BPr_128 supports one GPU\n"); return 2; }
    if (gpus > 1) { /* a rank whose device is missing would leave the others waiting in the NCCL rendezvous: check before forking.
                       The probe runs in a child so that the parent forks its ranks without a CUDA context of its own. */
        fflush(stdout);
        pid_t pr = fork();
        if (pr == 0) _exit(pg_device_count() >= (int)gpus ? 0 : 1);
        int st = 0;
        if (pr < 0 || waitpid(pr, &st, 0) < 0 || !WIFEXITED(st) || WEXITSTATUS(st) != 0) {
            fprintf(stderr, "polar_sim: --gpus %ld needs that many sm_100 devices\n", gpus);
            return 3;
        }
    }
    fflush(stdout);
    int rank = 0, have_id = 0;
    unsigned char id[128];
    if (gpus > 1) {
        {   /* stdout carries only the reference's result lines: keep NCCL's version banner off it */
            const char *dbg = getenv("NCCL_DEBUG");
            /* the VERSION and WARN levels both print the banner to stdout; INFO/TRACE are the user's explicit choice */
            if (!dbg || !strcasecmp(dbg, "VERSION") || !strcasecmp(dbg, "WARN")) setenv("NCCL_DEBUG", "NONE", 1);
        }
        rank = spawn_ranks(gpus, id, &have_id);
    }
    pg_ctx *ctx = NULL;
    {
        pg_params q = p;
        q.device = rank; q.rank = rank; q.nranks = (int)gpus;
        if (pg_create(&q, &ctx)) { fprintf(stderr, "polar_sim: pg_create (GPU %d): %s\n", rank, pg_last_error(NULL)); return 3; }
        if (have_id && pg_comm_init(ctx, id)) die("pg_comm_init", ctx);
        if (pd->fmt == F_BPR && pg_bpr_config(ctx, BPR_SAMPLES, 6)) die("pg_bpr_config", ctx);
    }

    ref_state st;
    ranq1_seed(&st.g, SEED);
    st.m = 0;
    uint64_t next_frame = 0; /* global frame counter: the PN phase keeps running across Eb/N0 points as in the reference */
    for (double bSNR_dB = e0; bSNR_dB <= e1; bSNR_dB += estep) {
        pg_counters c;
        struct timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        const int bpr = (pd->fmt == F_BPR);
        memset(g_bprE, 0, sizeof(g_bprE));
        if (use_ref) {
            ref_point(ctx, &p, bSNR_dB, (uint64_t)ble, (uint64_t)maxf, &st, &c, bpr);
        } else {
            /* BPR counters are summed on the device over whole rounds, so the BPr program counts whole rounds too
               (error block >= BLE) to keep E, run and the error counts on the same set of frames */
            if (bpr && pg_bpr_reset(ctx)) die("pg_bpr_reset", ctx);
            if (pg_simulate(ctx, bSNR_dB, next_frame, (uint64_t)ble, (uint64_t)maxf, bpr ? 0 : 1, &c)) die("pg_simulate", ctx);
            if (bpr) {
                if (pg_bpr_read(ctx, g_bprE)) die("pg_bpr_read", ctx);
            }
        }
        clock_gettime(CLOCK_MONOTONIC, &t1);
        next_frame += c.frames;
        print_point(pd, &p, bSNR_dB, &c);
        if (verbose && rank == 0) {
            const double s = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
            fprintf(stderr, "# %s Eb/N0 %.2f: %llu frames in %.3f s = %.3f Mframes/s = %.4f Gbit/s info; tie frames %llu, CRC-fail frames %llu, BP sweeps/frame %.2f\n",
                    pd->name, bSNR_dB, (unsigned long long)c.frames, s, (double)c.frames / s / 1e6, (double)c.frames * p.K / s / 1e9,
                    (unsigned long long)c.tie_frames, (unsigned long long)c.crc_fail, c.frames ? (double)c.bp_sweeps / (double)c.frames : 0.0);
        }
    }
    pg_destroy(ctx);
    if (rank == 0 && gpus > 1) {
        int status = 0, bad = 0;
        while (wait(&status) > 0) bad |= !(WIFEXITED(status) && WEXITSTATUS(status) == 0);
        return bad ? 4 : 0;
    }
    return 0;
}
