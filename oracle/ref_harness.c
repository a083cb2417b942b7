/* TEST INFRASTRUCTURE (oracle) -- never linked into the product.
 *
 * Include-harness around ONE unmodified reference program.  Built by oracle/Makefile as
 *   gcc -O2 -shared -fPIC -DREF_FILE='"<ref>/CASCL_1024_L8.c"' -DREF_KIND=2 -DREF_DECODE=CASCL ...
 * into oracle/_ref/libref_<prog>.so.  The reference's main() is renamed and never run; this
 * file repeats the ~30 set-up lines of that main() (graph allocation + connectBCB wiring +
 * information set, e.g. CASCL_1024_L8.c:166-217) using the reference's OWN globals and
 * functions, and then calls the reference's OWN decoder on caller-supplied LLRs.
 *
 * Shared-LLR trick: the reference forms the channel LLR as
 * 2*y/std/std (SC_128.c:418); with std = 1.0 and y = llr/2 that expression returns llr
 * exactly (2*(llr/2) is exact in binary fp, /1.0 is exact).
 *
 * REF_KIND: 1 = SC (scalar node), 2 = list decoders (SCL / CA-SCL: PM arrays),
 *           3 = BP, 4 = BPr (extra truth argument).   REF_FAG: Kao graph (builds bRev).
 * The reference sources #define N K n r L ..., so nothing below may use those names as
 * identifiers. */
#define main ref_main_unused
#include REF_FILE
#undef main

#include <string.h>

static int g_ready = 0;

int ref_param_N(void) { return N; }
int ref_param_K(void) { return K; }
int ref_param_nI(void) { return (int)(sizeof(I) / sizeof(I[0])); }
#if REF_KIND == 2
int ref_param_L(void) { return L; }
#else
int ref_param_L(void) { return 1; }
#endif
#if REF_KIND == 3 || REF_KIND == 4
int ref_param_iter(void) { return iterMax; }
#else
int ref_param_iter(void) { return 0; }
#endif

/* mirror of the reference main()'s set-up block */
void ref_init(void)
{
    int s, p, cnt;
    if (g_ready) return;
    V = (node ***)calloc(n + 1, sizeof(node **));
    for (s = 0; s <= n; s++) {
        V[s] = (node **)calloc(N, sizeof(node *));
        for (p = 0; p < N; p++) V[s][p] = (node *)calloc(1, sizeof(node));
    }
    memset(initV, 0, sizeof(initV));
    for (p = 0; p < N; p++) { V[0][p]->pU = NULL; V[0][p]->pL = NULL; connectBCB(0, p); }
    for (s = 1; s < n; s++)
        for (p = 0; p < N; p++) connectBCB(s, p);
    for (p = 0; p < N; p++) { V[n][p]->cU = NULL; V[n][p]->cL = NULL; }
#if REF_KIND == 2
    PM = (double *)calloc(2 * L, sizeof(double));
    PMcand = (double *)calloc(2 * L, sizeof(double));
#endif
#ifdef REF_FAG
    for (p = 0; p < N; p++) {
        int t = p, q = 0;
        for (s = n - 1; s >= 0; s--) { if (t % 2 == 1) q += pow2(s); t /= 2; }
        bRev[p] = q;
    }
#endif
    cnt = ref_param_nI();
    for (p = 0; p < N; p++) inI[p] = 0;
    for (p = 0; p < cnt; p++) { I[p] = Q[N - cnt + p]; inI[I[p]] = 1; }
    g_ready = 1;
}

/* information(+CRC) positions in the reference's reliability order, and the membership mask */
void ref_info_set(int *idx_out, int *mask_out)
{
    int p, cnt = ref_param_nI();
    ref_init();
    if (idx_out) for (p = 0; p < cnt; p++) idx_out[p] = I[p];
    if (mask_out) for (p = 0; p < N; p++) mask_out[p] = inI[p];
}

/* decode nframes frames of N LLRs each with the reference decoder; u_hat is nframes x N ints.
 * truth (nframes x N) is only read by REF_KIND 4 (BPr), may be NULL otherwise. */
void ref_decode_llr(const double *llr, int nframes, int *u_hat, const int *truth)
{
    static double ybuf[N];
    int f, p;
    ref_init();
    std = 1.0;
    for (f = 0; f < nframes; f++) {
        for (p = 0; p < N; p++) ybuf[p] = llr[(size_t)f * N + p] / 2;
        for (p = 0; p < N; p++) u_hat[(size_t)f * N + p] = 0;
#if REF_KIND == 4
        REF_DECODE(ybuf, u_hat + (size_t)f * N, (int *)(truth + (size_t)f * N));
#else
        (void)truth;
        REF_DECODE(ybuf, u_hat + (size_t)f * N);
#endif
    }
}

/* the reference check-node operation and uniform/normal generators, for unit comparisons */
double ref_chk(double a, double b) { return CHK(a, b); }

#ifndef REF_NO_RNG
/* restart the reference generator; programs with `const SEED` (SC_*, SCL_128/1024) ignore seed */
void ref_rng_restart(unsigned long long seed)
{
#ifdef REF_SEED_MUTABLE
    SEED = seed;
#else
    (void)seed;
#endif
    RANI = 0;
}
unsigned long long ref_rng_seed(void) { return SEED; }
void ref_normal_pair(double sigma, double *a, double *b)
{
    std = sigma;
    normal();
    *a = n1;
    *b = n2;
}
#endif

#if REF_KIND == 2
/* path-metric increment of the reference for a given leaf LLR: plants lambda in V[0][0] path 0 */
double ref_phi(double lambda, int u)
{
    ref_init();
    V[0][0]->l[0] = lambda;
    return PHI(0, 0, u);
}
#endif

#if REF_KIND == 4
/* BPR statistic accumulated by BPr() (BPr_128.c:71,418-568): E[sample][stage] */
int ref_bpr_rows(void) { return (int)(sizeof(E) / sizeof(E[0])); }
void ref_bpr_get_E(int *out) { memcpy(out, E, sizeof(E)); }
void ref_bpr_reset_E(void) { memset(E, 0, sizeof(E)); }
#endif
