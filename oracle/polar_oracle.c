/* TEST INFRASTRUCTURE (oracle) -- see polar_oracle.h.  CPU restatement, array form, of the
 * decoders and Monte-Carlo loops of CHEBSB/PolarDecoding.  Not part of the product. */
#include "polar_oracle.h"
#include "../include/polar_q_table.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef PO_REAL
#define PO_REAL double
#endif
typedef PO_REAL real;

/* ---------------------------------------------------------------- code construction */

static int ilog2(int v) { int k = 0; while ((1 << k) < v) k++; return k; }

int po_code_init(po_code *c, int N, int K, int r, uint64_t crc_poly, int crc_systematic)
{
    int i, k = 0, q[PO_MAX_N];
    if (N < 2 || N > 1024 || (N & (N - 1)) || K < 1 || r < 0 || r > PO_MAX_CRC || K + r > N) return -1;
    memset(c, 0, sizeof(*c));
    c->N = N; c->n = ilog2(N); c->K = K; c->r = r; c->nI = K + r;
    c->crc_poly = crc_poly; c->crc_systematic = crc_systematic;
    /* reliability sequence for length N = the 1024 table restricted to indices < N */
    for (i = 0; i < POLAR_Q_TABLE_LEN; i++)
        if (polar_q_table_1024[i] < N) q[k++] = polar_q_table_1024[i];
    /* SC_128.c:143-147: the nI most reliable, kept in reliability order */
    for (i = 0; i < c->nI; i++) { c->I[i] = q[N - c->nI + i]; c->inI[c->I[i]] = 1; }
    return 0;
}

#define CRC24_POLY ((1ull<<24)|(1ull<<23)|(1ull<<21)|(1ull<<20)|(1ull<<17)|(1ull<<15)|(1ull<<13)|(1ull<<12)|(1ull<<8)|(1ull<<4)|(1ull<<2)|(1ull<<1)|1ull)
#define CRC6_POLY ((1ull<<6)|(1ull<<5)|1ull)

int po_code_preset(po_code *c, const char *prog, int *L_out, int *iters_out)
{
    int L = 1, it = 0, rc = -1;
    if (!strcmp(prog, "SC_128") || !strcmp(prog, "SC_128_fag")) rc = po_code_init(c, 128, 64, 0, 0, 0);
    else if (!strcmp(prog, "SC_1024")) rc = po_code_init(c, 1024, 512, 0, 0, 0);
    else if (!strcmp(prog, "SCL_128") || !strcmp(prog, "SCL_128_fag")) { L = 8; rc = po_code_init(c, 128, 64, 0, 0, 0); }
    else if (!strcmp(prog, "SCL_1024")) { L = 8; rc = po_code_init(c, 1024, 512, 0, 0, 0); }
    else if (!strcmp(prog, "CASCL_128")) { L = 8; rc = po_code_init(c, 128, 64, 6, CRC6_POLY, 0); }
    else if (!strcmp(prog, "CASCL_1024_L8")) { L = 8; rc = po_code_init(c, 1024, 512, 24, CRC24_POLY, 0); }
    else if (!strcmp(prog, "CASCL_1024_sys")) { L = 8; rc = po_code_init(c, 1024, 512, 24, CRC24_POLY, 1); }
    else if (!strcmp(prog, "BP_128") || !strcmp(prog, "BP_128_fag")) { it = 100; rc = po_code_init(c, 128, 64, 0, 0, 0); }
    else if (!strcmp(prog, "BP_1024")) { it = 100; rc = po_code_init(c, 1024, 512, 0, 0, 0); }
    else if (!strcmp(prog, "BPr_128")) { it = 90; rc = po_code_init(c, 128, 64, 0, 0, 0); }
    if (L_out) *L_out = L;
    if (iters_out) *iters_out = it;
    return rc;
}

/* ---------------------------------------------------------------- primitives */

/* eight-level table for ln(1+e^-x) (SC_128.c:293-300, SCL_1024.c:489-496) */
static real tbl(real a)
{
    if (a < (real)0.196) return (real)0.65;
    if (a < (real)0.433) return (real)0.55;
    if (a < (real)0.71) return (real)0.45;
    if (a < (real)1.05) return (real)0.35;
    if (a < (real)1.508) return (real)0.25;
    if (a < (real)2.252) return (real)0.15;
    if (a < (real)4.5) return (real)0.05;
    return (real)0;
}

/* SC_128.c:284-315.  delta is formed as T(|a+b|) then -= T(|a-b|); when |a-b| >= 4.5 nothing is
 * subtracted, which equals subtracting 0.  Sign convention: sign(0) = +1. */
static real chk(real a, real b)
{
    real sa = (real)fabs((double)(a + b)), da = (real)fabs((double)(a - b));
    real delta = tbl(sa);
    real aa = (real)fabs((double)a), ab = (real)fabs((double)b);
    real m;
    int sgn;
    delta -= tbl(da);
    sgn = ((a >= 0) ? 1 : -1) * ((b >= 0) ? 1 : -1);
    m = (aa > ab) ? ab : aa;
    return (real)sgn * m + delta;
}

/* SCL_1024.c:481-502 */
static real phi(real lambda, int u)
{
    real a = (real)fabs((double)lambda);
    real res = tbl(a);
    if ((u == 0 && lambda < 0) || (u == 1 && lambda > 0)) res += a;
    return res;
}

double po_chk(double a, double b) { return (double)chk((real)a, (real)b); }
double po_phi(double lambda, int u) { return (double)phi((real)lambda, u); }

/* ---------------------------------------------------------------- SC / SCL state */

typedef struct path {
    real l[PO_MAX_LOGN + 1][PO_MAX_N];    /* l[s][p]; l[n] = channel */
    uint8_t b[PO_MAX_LOGN + 1][PO_MAX_N]; /* partial sums b[s][p]; b[0] = decided u */
    real pm;
} path;

/* LLRs that bit j needs and that are not yet known: g-layer at stage t=ctz(j), then f-layers
 * t-1..0 (equals the set getLLR() evaluates lazily, SC_128.c:345-365) */
static void llr_for_bit(path *P, int n, int j)
{
    int t, s, p, d;
    if (j == 0) t = n;
    else { t = 0; while (!((j >> t) & 1)) t++; }
    if (t < n) {
        d = 1 << t;
        for (p = j; p < j + d; p++) /* lower-left nodes: SC_128.c:355-359 */
            P->l[t][p] = (P->b[t][p - d] == 0) ? P->l[t + 1][p] + P->l[t + 1][p - d]
                                               : P->l[t + 1][p] - P->l[t + 1][p - d];
    }
    for (s = t - 1; s >= 0; s--) {
        d = 1 << s;
        for (p = j; p < j + d; p++) /* upper-left nodes: SC_128.c:353-354 */
            P->l[s][p] = chk(P->l[s + 1][p], P->l[s + 1][p + d]);
    }
}

/* partial-sum push after u_j is known (updateBit, SC_128.c:368-392): every butterfly whose two left
 * bits are now known gets its right bits */
static void push_bit(path *P, int n, int j)
{
    int s, i, d, base;
    for (s = 0; s < n && ((j >> s) & 1); s++) {
        d = 1 << s;
        base = j + 1 - 2 * d;
        for (i = 0; i < d; i++) {
            P->b[s + 1][base + i] = P->b[s][base + i] ^ P->b[s][base + i + d];
            P->b[s + 1][base + i + d] = P->b[s][base + i + d];
        }
    }
}

/* clone everything a continuation of path src from bit j on can still read (the reference copies all
 * n*N nodes, CASCL_1024_L8.c:473-500; unread entries do not influence the result) */
static void clone_path(path *dst, const path *src, int n, int N, int j)
{
    int s, lo, len;
    for (s = 0; s < n; s++) {
        len = 2 << s;
        if (len > N) len = N;
        lo = (j / len) * len;
        memcpy(&dst->l[s][lo], &src->l[s][lo], (size_t)len * sizeof(real));
        memcpy(&dst->b[s][lo], &src->b[s][lo], (size_t)len);
    }
    memcpy(dst->b[0], src->b[0], (size_t)N);
    dst->pm = src->pm;
}

void po_sc_decode(const po_code *c, const double *llr, int *u_hat)
{
    static path P;
    int j;
    for (j = 0; j < c->N; j++) P.l[c->n][j] = (real)llr[j];
    for (j = 0; j < c->N; j++) {
        llr_for_bit(&P, c->n, j);
        P.b[0][j] = (c->inI[j] && !(P.l[0][j] >= 0)) ? 1 : 0; /* SC_128.c:426-431 */
        push_bit(&P, c->n, j);
    }
    for (j = 0; j < c->N; j++) u_hat[j] = P.b[0][j];
}

int po_crc_check(const po_code *c, const int *cw)
{
    int C[PO_MAX_N], i, e;
    for (i = 0; i < c->nI; i++) C[i] = cw[i];
    for (i = c->nI - 1; i >= c->r; i--)
        if (C[i]) /* subtract D^(i-r) g(D) */
            for (e = 0; e <= c->r; e++)
                if ((c->crc_poly >> e) & 1) C[i - c->r + e] ^= 1;
    for (i = c->r - 1; i >= 0; i--)
        if (C[i]) return 0;
    return 1;
}

static int cmp_real(const void *a, const void *b)
{
    real x = *(const real *)a, y = *(const real *)b;
    return (x < y) ? -1 : (x > y);
}

void po_scl_decode(const po_code *c, int L, int use_crc, const double *llr, int *u_hat, int *flags_out)
{
    static path *P = NULL;
    const int N = c->N, n = c->n;
    int act = 1, j, k, i, flags = 0, best;
    real cand[2 * PO_MAX_L], sorted[2 * PO_MAX_L], med;
    int surv[PO_MAX_L];
    if (!P) P = (path *)malloc(sizeof(path) * PO_MAX_L);
    for (j = 0; j < N; j++) P[0].l[n][j] = (real)llr[j];
    for (k = 1; k < L; k++) memcpy(P[k].l[n], P[0].l[n], (size_t)N * sizeof(real));
    P[0].pm = 0;
    for (j = 0; j < N; j++) {
        for (k = 0; k < act; k++) llr_for_bit(&P[k], n, j);
        if (!c->inI[j]) { /* frozen: PM only (SCL_1024.c:601-604,662-665) */
            for (k = 0; k < act; k++) { P[k].pm += phi(P[k].l[0][j], 0); P[k].b[0][j] = 0; push_bit(&P[k], n, j); }
            continue;
        }
        if (act < L) { /* list filling (SCL_1024.c:586-600): path k -> k (bit 0) and k+act (bit 1) */
            for (k = 0; k < act; k++) clone_path(&P[k + act], &P[k], n, N, j);
            for (k = 0; k < act; k++) {
                P[k + act].pm = P[k].pm + phi(P[k].l[0][j], 1);
                P[k].pm = P[k].pm + phi(P[k].l[0][j], 0);
                P[k].b[0][j] = 0; P[k + act].b[0][j] = 1;
                push_bit(&P[k], n, j); push_bit(&P[k + act], n, j);
            }
            act *= 2;
            continue;
        }
        /* full list (SCL_1024.c:610-661) */
        for (k = 0; k < L; k++) {
            cand[k] = P[k].pm + phi(P[k].l[0][j], 0);
            cand[k + L] = P[k].pm + phi(P[k].l[0][j], 1);
        }
        memcpy(sorted, cand, sizeof(real) * 2 * (size_t)L);
        qsort(sorted, 2 * (size_t)L, sizeof(real), cmp_real);
        med = sorted[L];
        if (sorted[L - 1] == med) {
            /* exact tie across the list boundary: the reference drops both tied candidates and goes on
             * with an uninitialised path.  Total order (value, candidate index) instead; flagged. */
            int order[2 * PO_MAX_L], a, bb, tmp;
            uint8_t keep[2 * PO_MAX_L];
            flags |= 1;
            for (a = 0; a < 2 * L; a++) order[a] = a;
            for (a = 1; a < 2 * L; a++) /* stable insertion sort by value */
                for (bb = a; bb > 0 && cand[order[bb]] < cand[order[bb - 1]]; bb--) { tmp = order[bb]; order[bb] = order[bb - 1]; order[bb - 1] = tmp; }
            memset(keep, 0, sizeof(keep));
            for (a = 0; a < L; a++) keep[order[a]] = 1;
            for (k = 0; k < L; k++) surv[k] = keep[k] ? (keep[k + L] ? 2 : 0) : (keep[k + L] ? 1 : -1);
        } else {
            for (k = 0; k < L; k++) {
                int s0 = cand[k] < med, s1 = cand[k + L] < med;
                surv[k] = s0 ? (s1 ? 2 : 0) : (s1 ? 1 : -1);
            }
        }
        i = 0;
        for (k = 0; k < L; k++) {
            if (surv[k] == 0) { P[k].pm = cand[k]; P[k].b[0][j] = 0; push_bit(&P[k], n, j); }
            else if (surv[k] == 1) { P[k].pm = cand[k + L]; P[k].b[0][j] = 1; push_bit(&P[k], n, j); }
            else if (surv[k] == 2) {
                while (surv[i] != -1) i++; /* lowest-index free slot, cursor never rewinds (SCL_1024.c:650) */
                clone_path(&P[i], &P[k], n, N, j);
                P[k].pm = cand[k]; P[k].b[0][j] = 0; push_bit(&P[k], n, j);
                P[i].pm = cand[k + L]; P[i].b[0][j] = 1; push_bit(&P[i], n, j);
                surv[i] = -2;
            }
        }
    }
    best = -1;
    if (use_crc) { /* CASCL_1024_L8.c:725-745 */
        int cw[PO_MAX_N];
        for (k = 0; k < act; k++) {
            for (i = 0; i < c->nI; i++) cw[i] = P[k].b[0][c->I[i]];
            if (po_crc_check(c, cw) && (best < 0 || P[k].pm < P[best].pm)) best = k;
        }
        if (best < 0) flags |= 2;
    }
    if (best < 0) { /* SCL_1024.c:667-674 */
        best = 0;
        for (k = 1; k < act; k++)
            if (P[k].pm < P[best].pm) best = k;
    }
    for (j = 0; j < N; j++) u_hat[j] = P[best].b[0][j];
    if (flags_out) *flags_out = flags;
}

/* ---------------------------------------------------------------- BP */

typedef struct bp_state {
    real l[PO_MAX_LOGN + 1][PO_MAX_N];
    real r[PO_MAX_LOGN + 1][PO_MAX_N];
} bp_state;

static void bp_init(const po_code *c, const double *llr, bp_state *S)
{
    int s, j;
    for (s = 0; s < c->n; s++) for (j = 0; j < c->N; j++) S->l[s][j] = 0;
    for (j = 0; j < c->N; j++) S->l[c->n][j] = (real)llr[j];
    for (s = 1; s <= c->n; s++) for (j = 0; j < c->N; j++) S->r[s][j] = 0;
    for (j = 0; j < c->N; j++) S->r[0][j] = c->inI[j] ? (real)0 : (real)999; /* BP_1024.c:387-392 */
}

/* one round trip: R pass s=0..n-1 then L pass s=n-1..0 (BP_1024.c:394-415); returns 1 if any l changed */
static int bp_sweep(const po_code *c, bp_state *S)
{
    const int N = c->N, n = c->n;
    int s, j, d, changed = 0;
    real a, b2;
    for (s = 0; s < n; s++) {
        d = 1 << s;
        for (j = 0; j < N; j++) {
            if (j & d) continue;
            a = chk(S->r[s][j], S->l[s + 1][j + d] + S->r[s][j + d]);
            b2 = S->r[s][j + d] + chk(S->r[s][j], S->l[s + 1][j]);
            S->r[s + 1][j] = a;
            S->r[s + 1][j + d] = b2;
        }
    }
    for (s = n - 1; s >= 0; s--) {
        d = 1 << s;
        for (j = 0; j < N; j++) {
            if (j & d) continue;
            a = chk(S->l[s + 1][j], S->l[s + 1][j + d] + S->r[s][j + d]);
            b2 = S->l[s + 1][j + d] + chk(S->r[s][j], S->l[s + 1][j]);
            if (s > 0 && (memcmp(&a, &S->l[s][j], sizeof(real)) || memcmp(&b2, &S->l[s][j + d], sizeof(real)))) changed = 1;
            S->l[s][j] = a;
            S->l[s][j + d] = b2;
        }
    }
    return changed;
}

void po_bp_decode(const po_code *c, int iters, const double *llr, int *u_hat, int *sweeps_out)
{
    static bp_state S;
    int it, j, fix = 0;
    bp_init(c, llr, &S);
    for (it = 0; it < iters; it++)
        if (!bp_sweep(c, &S) && !fix) fix = it + 1;
    for (j = 0; j < c->N; j++) /* BP_1024.c:417-425 */
        u_hat[j] = (c->inI[j] && !(S.l[0][j] + S.r[0][j] >= 0)) ? 1 : 0;
    if (sweeps_out) *sweeps_out = fix;
}

void po_bpr_decode(const po_code *c, int iters, const double *llr, const int *u_true, int *u_hat,
                   const int *samples, int ns, int *E)
{
    static bp_state S;
    static uint8_t bit[PO_MAX_N], nxt[PO_MAX_N];
    const int N = c->N, n = c->n;
    int it, j, s, k, q, d;
    bp_init(c, llr, &S);
    for (it = 0; it < iters; it++) {
        bp_sweep(c, &S);
        for (q = 0; q < ns; q++) {
            if (samples[q] != it + 1) continue;
            for (s = 0; s <= n; s++) { /* BPr_128.c:420-442: decide at stage s, un-encode back to stage 0 */
                for (j = 0; j < N; j++) bit[j] = (S.l[s][j] + S.r[s][j] >= 0) ? 0 : 1;
                for (k = s; k > 0; k--) {
                    d = 1 << (k - 1);
                    for (j = 0; j < N; j++) {
                        if (j & d) continue;
                        nxt[j + d] = bit[j + d];
                        nxt[j] = bit[j + d] ^ bit[j];
                    }
                    memcpy(bit, nxt, (size_t)N);
                }
                for (j = 0; j < c->K; j++)
                    if (bit[c->I[j]] != u_true[c->I[j]]) E[q * (n + 1) + s] += 1;
            }
        }
    }
    for (j = 0; j < N; j++)
        u_hat[j] = (c->inI[j] && !(S.l[0][j] + S.r[0][j] >= 0)) ? 1 : 0;
}

/* ---------------------------------------------------------------- reference random source, frames */

void po_rng_seed(po_rng *g, uint64_t seed)
{
    uint64_t v = seed ^ 4101842887655102017ull;
    v ^= v >> 21; v ^= v << 35; v ^= v >> 4;
    g->v = v * 2685821657736338717ull;
}

double po_rng_uniform(po_rng *g)
{
    uint64_t v = g->v;
    v ^= v >> 21; v ^= v << 35; v ^= v >> 4;
    g->v = v;
    /* the reference multiplies as (signed) long long and converts the unsigned product: same bits */
    return (double)(v * 2685821657736338717ull) * 5.42101086242752217E-20;
}

void po_rng_normal_pair(po_rng *g, double sigma, double *a, double *b)
{
    double x1, x2, s;
    do {
        x1 = 2 * po_rng_uniform(g) - 1;
        x2 = 2 * po_rng_uniform(g) - 1;
        s = x1 * x1 + x2 * x2;
    } while (s >= 1.0);
    *a = sigma * x1 * sqrt(-2 * log(s) / s);
    *b = sigma * x2 * sqrt(-2 * log(s) / s);
}

void po_pn63(int *pn)
{
    int reg[6] = {0, 0, 0, 0, 0, 0}, i, b;
    for (i = 0; i < 63; i++) {
        b = (i == 0) ? 1 : (i < 6) ? 0 : (reg[4] ^ reg[5]);
        pn[i] = b;
        reg[5] = reg[4]; reg[4] = reg[3]; reg[3] = reg[2]; reg[2] = reg[1]; reg[1] = reg[0]; reg[0] = b;
    }
}

void po_make_u(const po_code *c, const int *pn, int m, int *u)
{
    int w[PO_MAX_N], i, e;
    memset(u, 0, sizeof(int) * (size_t)c->N);
    memset(w, 0, sizeof(w));
    if (c->r == 0) {
        for (i = 0; i < c->K; i++) w[i] = pn[(m + i) % 63];
    } else if (!c->crc_systematic) { /* w(D) = v(D) g(D), CASCL_1024_L8.c:251-266 */
        for (i = 0; i < c->K; i++)
            if (pn[(m + i) % 63])
                for (e = 0; e <= c->r; e++)
                    if ((c->crc_poly >> e) & 1) w[i + e] ^= 1;
    } else { /* parity = v(D) D^r mod g(D) in w[0..r-1], data in w[r..] (CASCL_1024_sys.c:778-789) */
        int rem[PO_MAX_N];
        memset(rem, 0, sizeof(rem));
        for (i = 0; i < c->K; i++) { w[c->r + i] = pn[(m + i) % 63]; rem[c->r + i] = w[c->r + i]; }
        for (i = c->nI - 1; i >= c->r; i--)
            if (rem[i])
                for (e = 0; e <= c->r; e++)
                    if ((c->crc_poly >> e) & 1) rem[i - c->r + e] ^= 1;
        for (i = 0; i < c->r; i++) w[i] = rem[i];
    }
    for (i = 0; i < c->nI; i++) u[c->I[i]] = w[i];
}

void po_polar_encode(const po_code *c, const int *u, int *x)
{
    int s, j, d;
    memcpy(x, u, sizeof(int) * (size_t)c->N);
    for (s = 0; s < c->n; s++) {
        d = 1 << s;
        for (j = 0; j < c->N; j++)
            if (!(j & d)) x[j] ^= x[j + d];
    }
}

void po_simulate_ref(const po_code *c, int decoder, int L, int iters, double ebn0_db, int target,
                     int count_from, po_rng *g, int *m, po_point *out)
{
    int pn[63], u[PO_MAX_N], x[PO_MAX_N], uh[PO_MAX_N], i, bad;
    double y[PO_MAX_N], llr[PO_MAX_N], n1, n2;
    const double sigma = pow(10, ebn0_db / ((double)-20)); /* SC_128.c:167 */
    const int step = c->K % 63;
    po_pn63(pn);
    out->run = out->err_block = out->err_bit = 0;
    while (out->err_block < target) {
        po_make_u(c, pn, *m, u);
        po_polar_encode(c, u, x);
        for (i = 0; i < c->N; i += 2) { /* SC_128.c:194-202 */
            po_rng_normal_pair(g, sigma, &n1, &n2);
            y[i] = (x[i] == 0) ? 1 + n1 : -1 + n1;
            y[i + 1] = (x[i + 1] == 0) ? 1 + n2 : -1 + n2;
        }
        for (i = 0; i < c->N; i++) llr[i] = 2 * y[i] / sigma / sigma; /* SC_128.c:418 */
        if (decoder == 0) po_sc_decode(c, llr, uh);
        else if (decoder == 1) po_scl_decode(c, L, 0, llr, uh, NULL);
        else if (decoder == 2) po_scl_decode(c, L, 1, llr, uh, NULL);
        else po_bp_decode(c, iters, llr, uh, NULL);
        bad = 0;
        for (i = count_from; i < c->nI; i++)
            if (u[c->I[i]] != uh[c->I[i]]) { bad = 1; out->err_bit++; }
        out->err_block += bad;
        out->run++;
        *m += step;
        if (*m >= 63) *m -= 63;
    }
}
