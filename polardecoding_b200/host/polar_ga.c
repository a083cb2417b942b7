/* Host counterparts of the reference's density-evolution / Gaussian-approximation ANALYSIS programs
 * (no frames, no GPU work: deterministic O(iter * n * N) scalar recursions, < 1 s on one core; SURVEY.md
 * section 2 marks them "host"):
 *
 *   BPDEGA_128        /root/reference/BPDEGA_128.c        DE-GA of BP, N=128, 100 sweeps, union-bound BLER per Eb/N0
 *   BPRGA_128         /root/reference/BPRGA_128.c         DE-GA of BPR (decision at every stage, re-encoded), 28 sweeps
 *   BPRGA_1024        /root/reference/BPRGA_1024.c        same, N=1024, 2.5 dB, rows for sweeps 6..10
 *   BPRGA_128_allbit  /root/reference/BPRGA_128_allbit.c  sum over ALL nodes of a stage (Kao graph), 3.0 dB
 *   BPRGA_128_W       /root/reference/BPRGA_128_W.c       Kao graph; stage decisions W=l+r combined into the u domain through the
 *                                                         sparse matrices M_i read from stdin, with the table-corrected CHK
 *   BPRGA_128_M       /root/reference/BPRGA_128_M.c       same with W = ln(2/erfc(sqrt(l+r)/2) - 1) and p = (1 - tanh(W/2))/2
 *   BPRGA_1024_W      /root/reference/BPRGA_1024_W.c      N=1024, (int)(40/EbN0) sweeps, %e formats
 *   The three matrix programs read the reference's stdin format (n*N column weights, then the row indices of the ones of every
 *   column, stage by stage).  `--emit-m` prints that file from the closed form M_k[r][c] = 1 <=> r, c agree on the low n-k bits
 *   and the top k bits of c are a subset of the top k bits of r (the author generated it with MATLAB).
 *
 * Array form on the Lee graph (stage s couples p and p+2^s); the Kao-graph program is the same recursion with
 * every position bit-reversed, which only matters for the order in which its sums are accumulated.  The
 * arithmetic (phi, phi_inv, derivative_phi: BPRGA_128.c:213-285; the update rule :311-345; the BPR
 * propagation :346-376) is kept in the reference's operation order so that the printed tables are identical.
 * Program identity: -DPOLAR_PROGRAM, --program NAME or argv[0]. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/polar_q_table.h"

#ifndef POLAR_PROGRAM
#define POLAR_PROGRAM ""
#endif
#define MAXN 1024
#define MAXL 11

static int N, K, n;
static int I[MAXN], inI[MAXN];
static double l[MAXL][MAXN], r[MAXL][MAXN], uu[MAXN], E[MAXL];
static double sd;

static double phi(double x)
{
    if (x < 0) { printf("illegal input for phi function!\n"); return 1; }
    if (x <= 0.1910) return exp(0.1047 * x * x - 0.4992 * x);
    if (x <= 0.7420) return 0.9981 * exp(0.05315 * x * x - 0.4795 * x);
    if (x <= 9.2254) return exp(-0.4527 * pow(x, 0.86) + 0.0218);
    if (x <= 15) return exp(-0.2832 * x - 0.4254);
    if (x <= 25) return exp(-0.26725134794 * x - 0.6646297809);
    return sqrt(3.14159265 / x) * exp(-x / 4) * (1 - 10.0 / 7.0 / x);
}

static double derivative_phi(double x)
{
    if (x < 0) { printf("illegal input for phi's derivative'!\n"); return 1; }
    if (x <= 0.1910) return (0.2094 * x - 0.4992) * exp(0.1047 * x * x - 0.4992 * x);
    if (x <= 0.7420) return 0.9981 * (0.1063 * x - 0.4795) * exp(0.05315 * x * x - 0.4795 * x);
    if (x <= 9.2254) return -0.389322 * exp(0.0218 - 0.4527 * pow(x, 0.86)) / pow(x, 0.14);
    if (x <= 15) return -0.2832 * exp(-0.2832 * x - 0.4254);
    if (x <= 25) return -0.26725134794 * exp(-0.26725134794 * x - 0.6646297809);
    return exp(-x / 4) * sqrt(3.14159265 / x)
           * (-0.5 / x * (1 - 10.0 / 7.0 / x) - 0.25 * (1 - 10.0 / 7.0 / x) + 10.0 / 7.0 / x / x);
}

static double phi_inv(double x)
{
    double x0, x1, delta;
    if (x <= 1 && x >= 0.91253609394) return (0.4992 - sqrt(0.24920064 + 0.4188 * log(x))) / 0.2094;
    if (x >= 0.72005453218) return (0.4795 - sqrt(0.22992025 + 0.2126 * log(x / 0.9981))) / 0.1063;
    if (x >= 0.04792905738) return pow((log(x) - 0.0218) / -0.4527, 1 / 0.86);
    if (x >= 0.00934045792) return -(log(x) + 0.4254) / 0.2832;
    if (x >= 0.0006452237) return -(log(x) + 0.6646297809) / 0.26725134794;
    x1 = 25 - (phi(25) - x) / derivative_phi(25);
    delta = fabs(x1 - 25);
    while (delta >= 1e-3) {
        x0 = x1;
        x1 = x1 - (phi(x1) - x) / derivative_phi(x1);
        delta = fabs(x1 - x0);
    }
    return x1;
}

static int brev(int v)
{
    int o = 0, b;
    for (b = 0; b < n; b++) if ((v >> b) & 1) o |= 1 << (n - 1 - b);
    return o;
}

static void init_messages(void)
{
    int s, j;
    for (s = 0; s < n; s++) for (j = 0; j < N; j++) l[s][j] = 0;
    for (j = 0; j < N; j++) l[n][j] = 2 / sd / sd;           /* BPRGA_128.c:300 */
    for (s = 1; s <= n; s++) for (j = 0; j < N; j++) r[s][j] = 0;
    for (j = 0; j < N; j++) r[0][j] = inI[j] ? 0 : 999;     /* :304-309 */
}

static void sweep(void) /* BPRGA_128.c:311-345 */
{
    int s, j, d;
    for (s = 0; s < n; s++) {
        d = 1 << s;
        for (j = 0; j < N; j++) {
            if (j & d) continue;
            const double ru = r[s][j], rl = r[s][j + d];
            r[s + 1][j] = phi_inv(phi(ru) + phi(rl + l[s + 1][j + d]) - phi(ru) * phi(rl + l[s + 1][j + d]));
            r[s + 1][j + d] = phi_inv(phi(ru) + phi(l[s + 1][j]) - phi(ru) * phi(l[s + 1][j])) + rl;
        }
    }
    for (s = n - 1; s >= 0; s--) {
        d = 1 << s;
        for (j = 0; j < N; j++) {
            if (j & d) continue;
            const double lu = l[s + 1][j], ll = l[s + 1][j + d];
            l[s][j] = phi_inv(phi(lu) + phi(r[s][j + d] + ll) - phi(lu) * phi(r[s][j + d] + ll));
            l[s][j + d] = phi_inv(phi(r[s][j]) + phi(lu) - phi(r[s][j]) * phi(lu)) + ll;
        }
    }
}

/* BPR: take l+r at stage i, push the means back to stage 0 through i check/copy layers, union bound over the payload (:348-372) */
static void bpr_row(int label)
{
    int i, k, j, d;
    double tl, tu;
    printf("%2d\t", label);
    for (i = 0; i <= n; i++) {
        for (j = 0; j < N; j++) uu[j] = l[i][j] + r[i][j];
        for (k = i; k > 0; k--) {
            d = 1 << (k - 1);
            for (j = 0; j < N; j++) {
                if (j & d) continue;
                tl = phi(uu[j]);
                tu = phi(uu[j + d]);
                uu[j] = phi_inv(tl + tu - tl * tu); /* upper path; the lower path keeps its value */
            }
        }
        E[i] = 0;
        for (j = 0; j < K; j++) E[i] += erfc(sqrt(uu[I[j]]) / 2.0);
        E[i] = E[i] * 0.5;
        printf("%lf\t", E[i]);
    }
    printf("\n");
}

static void allbit_row(int label) /* BPRGA_128_allbit.c:366-380: every node of the stage, Kao position order */
{
    int i, j;
    printf("%2d\t", label);
    for (i = 0; i <= n; i++) {
        E[i] = 0;
        for (j = 0; j < N; j++) {
            const int p = brev(j);
            if (erfc(sqrt(l[i][p] + r[i][p]) / 2.0) < 0) printf("ERR!\n");
            E[i] += erfc(sqrt(l[i][p] + r[i][p]) / 2.0);
        }
        E[i] = E[i] * 0.5;
        printf("%lf\t", E[i]);
    }
    printf("\n");
}

/* ---------------------------------------------------------------- the matrix ("_W" / "_M") programs, Kao graph */
static int Mw[MAXL][MAXN], *M1[MAXL][MAXN];
static int frozen[MAXL][MAXN];
static double Wv[MAXN], pv[MAXN];

static double tbl(double a)
{
    if (a < 0.196) return 0.65;
    if (a < 0.433) return 0.55;
    if (a < 0.71) return 0.45;
    if (a < 1.05) return 0.35;
    if (a < 1.508) return 0.25;
    if (a < 2.252) return 0.15;
    if (a < 4.5) return 0.05;
    return 0;
}
static double CHKd(double a, double b) /* BPRGA_128_W.c: the same CHK as the decoders (SC_128.c:284-315) */
{
    double delta = tbl(fabs(a + b));
    const double A1 = fabs(a), A2 = fabs(b);
    const int sgn = ((a >= 0) ? 1 : -1) * ((b >= 0) ? 1 : -1);
    delta -= tbl(fabs(a - b));
    return (A1 > A2) ? sgn * A2 + delta : sgn * A1 + delta;
}

static void emit_m(void)
{
    int k, c, rr;
    for (k = 1; k <= n; k++)
        for (c = 0; c < N; c++) {
            int w = 0;
            for (rr = 0; rr < N; rr++) {
                const int low = (1 << (n - k)) - 1;
                if ((rr & low) == (c & low) && (((c & ~low) & ~(rr & ~low)) == 0)) w++;
            }
            printf("%d%c", w, c + 1 < N ? ' ' : '\n');
        }
    for (k = 1; k <= n; k++)
        for (c = 0; c < N; c++) {
            const int low = (1 << (n - k)) - 1;
            for (rr = 0; rr < N; rr++)
                if ((rr & low) == (c & low) && (((c & ~low) & ~(rr & ~low)) == 0)) printf("%d ", rr);
            printf("\n");
        }
}

static int read_m(void) /* BPRGA_128_W.c:129-147 */
{
    int k, i, j, t;
    for (k = 1; k <= n; k++)
        for (i = 0; i < N; i++) {
            if (scanf("%d", &t) != 1 || t < 0 || t > N) return -1;
            Mw[k][i] = t;
            M1[k][i] = (int *)calloc((size_t)(t > 0 ? t : 1), sizeof(int));
        }
    for (k = 1; k <= n; k++)
        for (i = 0; i < N; i++)
            for (j = 0; j < Mw[k][i]; j++) {
                if (scanf("%d", &t) != 1 || t < 0 || t >= N) return -1;
                M1[k][i][j] = t;
            }
    return 0;
}

static void kao_sweep(void) /* the update rule on the Kao graph: stage i couples j and j + 2^(n-1-i) */
{
    int i, j, d;
    for (i = 0; i < n; i++) {
        d = 1 << (n - 1 - i);
        for (j = 0; j < N; j++) {
            if (j & d) continue;
            const double ru = r[i][j], rl = r[i][j + d];
            r[i + 1][j] = phi_inv(phi(ru) + phi(rl + l[i + 1][j + d]) - phi(ru) * phi(rl + l[i + 1][j + d]));
            r[i + 1][j + d] = phi_inv(phi(ru) + phi(l[i + 1][j]) - phi(ru) * phi(l[i + 1][j])) + rl;
        }
    }
    for (i = n - 1; i >= 0; i--) {
        d = 1 << (n - 1 - i);
        for (j = 0; j < N; j++) {
            if (j & d) continue;
            const double lu = l[i + 1][j], ll = l[i + 1][j + d];
            l[i][j] = phi_inv(phi(lu) + phi(r[i][j + d] + ll) - phi(lu) * phi(r[i][j + d] + ll));
            l[i][j + d] = phi_inv(phi(r[i][j]) + phi(lu) - phi(r[i][j]) * phi(lu)) + ll;
        }
    }
}

/* the isFrozen marks exactly as the reference's set-up loop produces them (BPRGA_128_W.c:147-163 with connectBCB :262-273):
 * stage 0 is marked while the graph is being wired, so a butterfly sees its lower-left mark only if that node was
 * already visited */
static void mark_frozen(void)
{
    int i, j, d;
    static int init[MAXL][MAXN];
    memset(init, 0, sizeof(init));
    memset(frozen, 0, sizeof(frozen));
    for (i = 0; i < n; i++) {
        d = 1 << (n - 1 - i);
        for (j = 0; j < N; j++) {
            if (i == 0) frozen[0][j] = inI[brev(j)] ? 0 : 1;
            if (init[i][j]) continue;
            init[i][j] = 1; init[i][j + d] = 1;
            if (frozen[i][j + d] == 1 && frozen[i][j] == 1) { frozen[i + 1][j] = 1; frozen[i + 1][j + d] = 1; }
            else if (frozen[i][j + d] == 1) { frozen[i + 1][j] = 0; frozen[i + 1][j + d] = 1; }
            else { frozen[i + 1][j] = 0; frozen[i + 1][j + d] = 0; }
        }
    }
}

/* kind 4: BPRGA_128_W, 5: BPRGA_128_M, 6: BPRGA_1024_W */
static int run_matrix_program(int kind, double e0, double e1, int iterMax)
{
    int i, j, k, iter, niter;
    double bSNR_dB, bler, tempL;
    if (read_m()) { fprintf(stderr, "polar_ga: this program reads the M matrices on stdin (see --emit-m)\n"); return 2; }
    if (kind != 5) mark_frozen();
    for (bSNR_dB = e0; bSNR_dB <= e1; bSNR_dB += 0.5) {
        printf("bSNR = %.2lf\t", bSNR_dB);
        sd = pow(10, bSNR_dB / ((double)-20));
        for (i = 0; i < n; i++) for (j = 0; j < N; j++) l[i][j] = 0;
        for (j = 0; j < N; j++) l[n][j] = 2 / sd / sd;
        for (i = 1; i <= n; i++) for (j = 0; j < N; j++) r[i][j] = 0;
        for (j = 0; j < N; j++) r[0][j] = inI[brev(j)] ? 0 : 99;
        niter = (kind == 6) ? (int)(iterMax / bSNR_dB) : (int)(iterMax - 6 * bSNR_dB);
        printf("iterMax = %d\n", niter);
        for (iter = 0; (kind == 6) ? (iter < niter) : (iter < (iterMax - 6 * bSNR_dB)); iter++) {
            kao_sweep();
            if ((kind == 4 && iter > 1) || (kind == 5 && iter > 1 && iter < 8) || (kind == 6 && iter > 3)) {
                printf(kind == 6 ? "%2d  " : "%2d\t", iter + 1);
                E[0] = 0;
                for (j = 0; j < K; j++) E[0] += erfc(sqrt(l[0][brev(I[j])]) / 2.0);
                E[0] = E[0] * 0.5;
                printf(kind == 6 ? "%.5e  " : "%lf  ", E[0]);
                for (i = 1; i <= n; i++) {
                    for (j = 0; j < N; j++)
                        Wv[j] = (kind == 5) ? log(2.0 / erfc(sqrt(l[i][j] + r[i][j]) / 2.0) - 1) : l[i][j] + r[i][j];
                    for (j = 0; j < K; j++) {
                        const int c = brev(I[j]);
                        tempL = Wv[M1[i][c][0]];
                        for (k = 1; k < Mw[i][c]; k++)
                            if (kind == 5 || frozen[i][M1[i][c][k]] == 0) tempL = CHKd(tempL, Wv[M1[i][c][k]]);
                        pv[c] = (kind == 5) ? 0.5 * (1 - tanh(tempL / 2)) : 0.5 * erfc(sqrt(tempL) / 2.0);
                    }
                    E[i] = 0;
                    for (j = 0; j < K; j++) E[i] += pv[brev(I[j])];
                    printf(kind == 6 ? "%.5e  " : "%lf  ", E[i]);
                }
                printf("\n");
            }
        }
        bler = 0;
        for (i = 0; i < K; i++) bler += erfc(sqrt(l[0][brev(I[i])]) / 2.0);
        bler = bler * 0.5;
        if (kind == 6) printf("BLER = %e\t\tBER = %e\n", bler, bler / K);
        else printf("BLER = %lf\t\tBER = %lfe-2\n", bler, bler * 100 / K);
    }
    return 0;
}

int main(int argc, char **argv)
{
    const char *name = POLAR_PROGRAM;
    int i, k = 0, iterMax, iter, q[MAXN], kind, want_m = 0;
    double e0, e1, bSNR_dB, bler;
    if (!name[0]) { const char *b = strrchr(argv[0], '/'); name = b ? b + 1 : argv[0]; }
    for (i = 1; i + 1 < argc; i++) if (!strcmp(argv[i], "--program")) name = argv[i + 1];
    for (i = 1; i < argc; i++) if (!strcmp(argv[i], "--emit-m")) want_m = 1;
    if (!strcmp(name, "BPDEGA_128")) { N = 128; K = 64; iterMax = 100; e0 = 1.0; e1 = 5; kind = 0; }
    else if (!strcmp(name, "BPRGA_128")) { N = 128; K = 64; iterMax = 28; e0 = 1.0; e1 = 4; kind = 1; }
    else if (!strcmp(name, "BPRGA_1024")) { N = 1024; K = 512; iterMax = 30; e0 = 2.5; e1 = 2.5; kind = 2; }
    else if (!strcmp(name, "BPRGA_128_allbit")) { N = 128; K = 64; iterMax = 30; e0 = 3.0; e1 = 3; kind = 3; }
    else if (!strcmp(name, "BPRGA_128_W")) { N = 128; K = 64; iterMax = 32; e0 = 2.0; e1 = 4; kind = 4; }
    else if (!strcmp(name, "BPRGA_128_M")) { N = 128; K = 64; iterMax = 32; e0 = 3.0; e1 = 4; kind = 5; }
    else if (!strcmp(name, "BPRGA_1024_W")) { N = 1024; K = 512; iterMax = 40; e0 = 2.0; e1 = 4; kind = 6; }
    else { fprintf(stderr, "polar_ga: unknown program '%s'\n", name); return 2; }
    for (n = 0; (1 << n) < N; n++) {}
    for (i = 0; i < POLAR_Q_TABLE_LEN; i++) if (polar_q_table_1024[i] < N) q[k++] = polar_q_table_1024[i];
    for (i = 0; i < K; i++) { I[i] = q[N - K + i]; inI[I[i]] = 1; }

    if (want_m) { emit_m(); return 0; }
    if (kind >= 4) return run_matrix_program(kind, e0, e1, iterMax);

    printf("iterMax = %d\n", iterMax);
    for (bSNR_dB = e0; bSNR_dB <= e1; bSNR_dB += 0.5) {
        sd = pow(10, bSNR_dB / ((double)-20));
        init_messages();
        for (iter = 0; iter < ((kind == 3) ? (iterMax - 4 * bSNR_dB) : iterMax); iter++) {
            sweep();
            if (kind == 1 && iter % 2 == 1) bpr_row(iter + 1);
            if (kind == 2 && iter >= 5 && iter <= 9) bpr_row(iter + 1);
            if (kind == 3 && iter < 10 && iter > 1) allbit_row(iter + 1);
        }
        bler = 0;
        for (i = 0; i < K; i++) bler += erfc(sqrt(l[0][I[i]]) / 2.0);
        bler = bler * 0.5;
        if (kind == 0) printf("bSNR = %.2lf\tBLER = %lf\t\tBER = %lf e-2\n", bSNR_dB, bler, bler * 100 / K);
        else printf("bSNR = %.2lf\tBLER = %lf\t\tBER = %lfe-2\n", bSNR_dB, bler, bler * 100 / K);
    }
    return 0;
}
