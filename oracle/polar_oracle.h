/* TEST INFRASTRUCTURE (oracle) -- CPU restatement of the reference algorithms.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (polardecoding_b200/) never does and has no CPU fallback.
 *
 * Parity status: PINNED.  tests/test_oracle_*.py check this restatement (a) frame-by-frame
 * against the compiled reference decoders (oracle/_ref/libref_*.so, built from the unmodified
 * /root/reference sources) on shared LLRs, (b) against the reference's own captured result files
 * (KAT K1/K2/K4/K5 of SURVEY.md section 4) through po_simulate_ref(), and (c) against golden
 * vectors in tests/golden/ that were produced by the compiled reference (tools/make_golden.py).
 *
 * Array formulation: Lee-graph node (s,p), s=0 is the u side, s=n the
 * channel side; the butterfly at stage s couples p and p+2^s (bit s of p clear).  Arithmetic is
 * IEEE double in the reference's operation order (PO_REAL=float builds the same code in fp32 to
 * measure rounding sensitivity). */
#ifndef POLAR_ORACLE_H
#define POLAR_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PO_MAX_N 2048
#define PO_MAX_LOGN 11
#define PO_MAX_L 32
#define PO_MAX_CRC 32

typedef struct po_code {
    int N, n;              /* block length and log2 */
    int K;                 /* data bits (PN-63 payload) */
    int r;                 /* CRC bits (0: none) */
    int nI;                /* K + r non-frozen positions */
    int crc_systematic;    /* 0: w(D)=v(D)g(D) (CASCL_1024_L8.c:245-266); 1: parity first, data after (CASCL_1024_sys.c:776-789) */
    uint64_t crc_poly;     /* bit e set <=> D^e in g(D), including D^r and D^0 */
    int I[PO_MAX_N];       /* non-frozen positions in RELIABILITY order: I[i]=Q[N-nI+i] (SC_128.c:143-147) */
    uint8_t inI[PO_MAX_N]; /* membership mask */
} po_code;

/* the reference's named configurations: "SC_128","SC_1024","SC_128_fag","SCL_128","SCL_1024",
 * "SCL_128_fag","CASCL_128","CASCL_1024_L8","CASCL_1024_sys","BP_128","BP_1024","BP_128_fag","BPr_128".
 * Returns 0 on success; *L_out / *iters_out get the list size / sweep count of that program. */
int po_code_preset(po_code *c, const char *prog, int *L_out, int *iters_out);
int po_code_init(po_code *c, int N, int K, int r, uint64_t crc_poly, int crc_systematic);

/* primitives */
double po_chk(double a, double b);            /* CHK, SC_128.c:284-315 */
double po_phi(double lambda, int u);          /* PHI, SCL_1024.c:481-502 */

/* decoders: llr[N] natural order (position p of the Lee graph = code bit p), u_hat[N] */
void po_sc_decode(const po_code *c, const double *llr, int *u_hat);                 /* SC_128.c:395-460 */
/* use_crc: CA-SCL final pick (CASCL_1024_L8.c:725-755) instead of min-PM (SCL_1024.c:667-674).
 * flags_out (may be NULL): bit0 = an exact PM tie straddled the list boundary at some bit ("Oops!",
 * SCL_1024.c:621) -- the reference's behaviour is then undefined and this restatement breaks the tie
 * by candidate index; bit1 = CA-SCL: no path passed the CRC. */
void po_scl_decode(const po_code *c, int L, int use_crc, const double *llr, int *u_hat, int *flags_out);
/* BP with the reference's round-trip sweep (BP_1024.c:372-427).  sweeps_out (may be NULL) = index of the
 * first sweep that left every l message bit-identical (1-based, counting that confirming sweep), or 0 if
 * none did within iters.  The decision is always that of the full iters sweeps. */
void po_bp_decode(const po_code *c, int iters, const double *llr, int *u_hat, int *sweeps_out);
/* BPr_128.c:373-580: BP plus the per-stage hard-decision statistic.  samples[ns] are the 1-based sweep
 * counts at which to sample (3,6,10,20,40,80); E is ns x (n+1), ADDED to. */
void po_bpr_decode(const po_code *c, int iters, const double *llr, const int *u_true, int *u_hat,
                   const int *samples, int ns, int *E);

/* reference random source and frame generation */
typedef struct po_rng { uint64_t v; } po_rng;
void po_rng_seed(po_rng *g, uint64_t seed);                       /* Ranq1 first-call block, SC_128.c:238-245 */
double po_rng_uniform(po_rng *g);                                 /* Ranq1, SC_128.c:246-249 */
void po_rng_normal_pair(po_rng *g, double sigma, double *a, double *b); /* normal(), SC_128.c:253-267 */
void po_pn63(int *pn);                                            /* SC_128.c:126-138 */
/* builds u[N] (frozen = 0) for PN phase m: data = PN[(m+i)%63], CRC-encoded if c->r>0 */
void po_make_u(const po_code *c, const int *pn, int m, int *u);
void po_polar_encode(const po_code *c, const int *u, int *x);     /* x = u F^{(x)n} */
int po_crc_check(const po_code *c, const int *cw);                /* CRcheck, CASCL_1024_L8.c:569-598; cw[nI] in I[] order */

/* Monte-Carlo loop of the reference main() (SC_128.c:164-222 and siblings): one Eb/N0 point, continuing
 * the generator and PN phase held in *g / *m, until `target` block errors.  decoder: 0 SC, 1 SCL, 2 CA-SCL, 3 BP.
 * count_from: first index i of I[] included in the error count (r for CASCL_1024_sys.c:821, else 0). */
typedef struct po_point { long run, err_block, err_bit; } po_point;
void po_simulate_ref(const po_code *c, int decoder, int L, int iters, double ebn0_db, int target,
                     int count_from, po_rng *g, int *m, po_point *out);

#ifdef __cplusplus
}
#endif
#endif
