// Belief-propagation decoding on the polar factor graph, one frame per CTA, messages resident in
// shared memory (sm_100a).
//
// What it computes (reference): BP() /root/reference/BP_1024.c:372-427 (BP_128.c:334, BP_128_fag.c:349):
// iterMax round trips, each an R pass over stages 0..n-1 followed by an L pass over stages n-1..0, every
// butterfly (j, j+2^s) updating
//     r(s+1,j)   = CHK(r(s,j), l(s+1,j+d) + r(s,j+d))      r(s+1,j+d) = r(s,j+d) + CHK(r(s,j), l(s+1,j))
//     l(s,j)     = CHK(l(s+1,j), l(s+1,j+d) + r(s,j+d))    l(s,j+d)   = l(s+1,j+d) + CHK(r(s,j), l(s+1,j))
// with l(n,.) = channel LLR, r(0,.) = 999 on frozen positions and 0 elsewhere, decision l(0,j)+r(0,j) >= 0 -> 0.
// BPr() /root/reference/BPr_128.c:373-580 adds the per-stage hard-decision statistic (optional, see bpr_*).
//
// How: the N/2 butterflies of a stage are independent, stages are sequential -> one butterfly (or two) per
// thread per stage and a barrier between stages.  Only l(1..n-1) and r(1..n-1) are state ((n-1)*2*N values:
// 72 KB in fp32 for N=1024, three frames resident per SM); l(n) lives in registers, r(0) is a bit mask.
// Two stage passes of the reference produce values nobody reads -- r(n) (R pass, last stage) and, in every
// sweep but the last, l(0) -- and are not executed; l(0) is formed once after the last sweep.
// Fixed-point stop (optional): the sweep map is deterministic and r is a function of l, so once an L pass
// leaves l(1..n-1) bit-identical the remaining sweeps cannot change anything; the frame then stops with the
// decisions the full iterMax sweeps would give.  Frames are pulled from a device-side queue so CTAs that
// finish early start the next frame.
#include "engine.h"
#include "polar_common.cuh"

// CHK as the BP kernels use it: BP sits near three limits at once (issue slots, FMA pipe, ALU pipe), so the table steps are
// split between the pipes: all seven accumulate packed (FFMA2: fewer issue slots, more FMA-pipe time) and the indicators of the
// first four come from the ALU pipe (FSET) instead of the FMA pipe (FFMA.SAT).  Bit-identical for every choice; measured at
// N=1024, Mframes/s for (packed, ALU) = (0,0) 0.424, (2,0) 0.432/0.4385, (6,3) 0.443, (7,3) 0.431, (7,4) 0.446, (7,5) 0.440,
// (7,7) 0.405 (tools/ab_bp.py; the two (2,0) figures are two boxes).
#ifndef POLAR_BP_KP
#define POLAR_BP_KP 7
#endif
#ifndef POLAR_BP_KM
#define POLAR_BP_KM 4
#endif

namespace polar {

// CHK as the BP kernels use it
template <typename real> __device__ __forceinline__ real bchk(real a, real b) { return chk<real>(a, b); }
#if POLAR_BP_KP > 0 || POLAR_BP_KM > 0
template <> __device__ __forceinline__ float bchk<float>(float a, float b) { return chk_mix_f32<POLAR_BP_KP, POLAR_BP_KM>(a, b); }
#endif

template <typename real, int LOGN, int THREADS>
struct BpCfg {
    static constexpr int N = 1 << LOGN;
    static constexpr int W = (N + 31) / 32;
    static constexpr int BPT = (N / 2) / THREADS;  // butterflies per thread per stage
    static constexpr size_t MSG = (size_t)2 * (LOGN - 1) * N;  // l(1..n-1), r(1..n-1)
    static constexpr size_t SMEM = MSG * sizeof(real) + (size_t)(2 * W + 4) * 4;  // messages, decisions + 4 words, re-encode check words
};

// two neighbouring messages as one 64/128-bit shared-memory access
template <typename real> struct pair_t;
template <> struct pair_t<float> { using type = float2; };
template <> struct pair_t<double> { using type = double2; };
template <typename real>
__device__ __forceinline__ void ld2(const real *p, real &a, real &b)
{
    const typename pair_t<real>::type v = *reinterpret_cast<const typename pair_t<real>::type *>(p);
    a = v.x; b = v.y;
}
template <typename real>
__device__ __forceinline__ void st2(real *p, real a, real b)
{
    typename pair_t<real>::type v; v.x = a; v.y = b;
    *reinterpret_cast<typename pair_t<real>::type *>(p) = v;
}

// the same through 32-bit shared-window addresses (the generic-pointer form makes the compiler rebuild the window base --
// S2UR + UMOV + ULEA -- in every stage: 6 of ~140 instructions per stage pass)
__device__ __forceinline__ void lds2(uint32_t a, float &x, float &y) { asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(x), "=f"(y) : "r"(a) : "memory"); }
__device__ __forceinline__ void lds2(uint32_t a, double &x, double &y) { asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(a) : "memory"); }
__device__ __forceinline__ void sts2(uint32_t a, float x, float y) { asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(x), "f"(y) : "memory"); }
__device__ __forceinline__ void sts2(uint32_t a, double x, double y) { asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(x), "d"(y) : "memory"); }

template <int THREADS>
__device__ __forceinline__ void cta_sync()
{
    if (THREADS == 32) __syncwarp();
    else __syncthreads();
}

template <typename real, int LOGN, int THREADS>
__global__ void __launch_bounds__(THREADS) bp_decode_kernel(const BpArgs a)
{
    using C = BpCfg<real, LOGN, THREADS>;
    using RT = real_traits<real>;
    constexpr int N = C::N, W = C::W, BPT = C::BPT, n = LOGN;
    static_assert(BPT >= 1, "too many threads for this N");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    real *Lm = reinterpret_cast<real *>(smem_raw);        // Lm[(s-1)*N + j] = l(s,j), s=1..n-1
    real *Rm = Lm + (size_t)(n - 1) * N;                  // Rm[(s-1)*N + j] = r(s,j), s=1..n-1
    uint32_t *uh = reinterpret_cast<uint32_t *>(smem_raw + C::MSG * sizeof(real));  // W words + [W]=nerr, [W+1]=frame lo, [W+2]=frame hi
    // BPR statistic (only when a.bpr_ns > 0; the launch then adds N + 8*16*4 bytes): one byte per position, per-CTA counters
    uint32_t *xh = uh + W + 4;  // hard decisions on the channel side (G-matrix stop)
    uint8_t *bb = reinterpret_cast<uint8_t *>(xh + W);
    uint32_t *bE = reinterpret_cast<uint32_t *>(bb + N);
    const int tid = threadIdx.x;
#ifndef POLAR_BP_GENERIC_SMEM
    uint32_t sa = (uint32_t)__cvta_generic_to_shared(Lm);  // message index i (l at i, r at RO + i) lives at sa + i*sizeof(real)
    asm volatile("mov.b32 %0, %0;" : "+r"(sa));                 // opaque: keep it in a register instead of rebuilding it per stage
    constexpr int RO = (LOGN - 1) * N;
    auto LD = [&](int i, real &x, real &y) { lds2(sa + (uint32_t)i * (uint32_t)sizeof(real), x, y); };
    auto ST = [&](int i, real x, real y) { sts2(sa + (uint32_t)i * (uint32_t)sizeof(real), x, y); };
#else
    constexpr int RO = (LOGN - 1) * N;
    auto LD = [&](int i, real &x, real &y) { ld2<real>(Lm + i, x, y); };
    auto ST = [&](int i, real x, real y) { st2<real>(Lm + i, x, y); };
#endif

    auto r0 = [&](int j) -> real { return ((a.m.info[j >> 5] >> (j & 31)) & 1u) ? (real)0 : (real)999; };

    for (;;) {
        if (tid == 0) {
            const unsigned long long f = atomicAdd(a.queue, 1ull);
            uh[W + 1] = (uint32_t)f;
            uh[W + 2] = (uint32_t)(f >> 32);
        }
        cta_sync<THREADS>();
        const unsigned long long frame = (unsigned long long)uh[W + 1] | ((unsigned long long)uh[W + 2] << 32);
        if (frame >= a.B) break;
        const real *llr = reinterpret_cast<const real *>(a.llr) + frame * (size_t)N;

        real ch_up[BPT], ch_lo[BPT];
#pragma unroll
        for (int i = 0; i < BPT; i++) {
            const int q = (BPT == 2) ? (2 * tid + i) : (tid + i * THREADS);
            ch_up[i] = __ldg(llr + q);
            ch_lo[i] = __ldg(llr + q + N / 2);
        }
        for (int i = tid; i < (n - 1) * N; i += THREADS) Lm[i] = (real)0;  // BP_1024.c:378-380
        if (tid < W + 1) uh[tid] = 0;
        if (a.bpr_ns) for (int i = tid; i < 8 * 16; i += THREADS) bE[i] = 0;
        cta_sync<THREADS>();

        // BPR sample (BPr_128.c:418-442): at every stage i decide on l(i,.)+r(i,.), undo encoder stages i-1..0, count wrong payload bits
        auto bpr_sample = [&](int q) {
            const uint32_t *tw = a.truth + frame * (size_t)W;
            for (int i = 0; i <= n; i++) {
#pragma unroll
                for (int e = 0; e < BPT; e++) {
                    const int qq = tid + e * THREADS;
                    real su, sl;  // l+r of the two positions this thread owns at stage i
                    int ju, jl;
                    if (i == 0) {  // l(0,.) is not state: form it from l(1,.), r(0,.)
                        ju = 2 * qq; jl = ju + 1;
                        const real lu = Lm[ju], ll = Lm[jl], ru = r0(ju), rl = r0(jl);
                        su = bchk<real>(lu, ll + rl) + ru;
                        sl = (ll + bchk<real>(ru, lu)) + rl;
                    } else if (i == n) {  // r(n,.) is not state: form it from r(n-1,.), l(n,.) = channel
                        ju = (BPT == 2) ? (2 * tid + e) : qq;  // the positions whose channel LLRs this thread holds
                        jl = ju + N / 2;
                        const real *rin = Rm + (size_t)(n - 2) * N;
                        const real ru = (n == 1) ? r0(ju) : rin[ju], rl = (n == 1) ? r0(jl) : rin[jl];
                        su = ch_up[e] + bchk<real>(ru, ch_lo[e] + rl);
                        sl = ch_lo[e] + (rl + bchk<real>(ru, ch_up[e]));
                    } else {
                        ju = 2 * qq; jl = ju + 1;
                        su = Lm[(size_t)(i - 1) * N + ju] + Rm[(size_t)(i - 1) * N + ju];
                        sl = Lm[(size_t)(i - 1) * N + jl] + Rm[(size_t)(i - 1) * N + jl];
                    }
                    bb[ju] = (su >= (real)0) ? 0 : 1;
                    bb[jl] = (sl >= (real)0) ? 0 : 1;
                }
                cta_sync<THREADS>();
                for (int kk = i; kk > 0; kk--) {  // BPr_128.c:428-437
                    const int d = 1 << (kk - 1);
#pragma unroll
                    for (int e = 0; e < BPT; e++) {
                        const int qq = tid + e * THREADS;
                        const int j = ((qq >> (kk - 1)) << kk) | (qq & (d - 1));
                        bb[j] ^= bb[j + d];
                    }
                    cta_sync<THREADS>();
                }
                uint32_t wrong = 0;
                for (int j = tid; j < N; j += THREADS)
                    if ((a.m.cnt[j >> 5] >> (j & 31)) & 1u) wrong += (uint32_t)(bb[j] != ((__ldg(tw + (j >> 5)) >> (j & 31)) & 1u));
                if (wrong) atomicAdd(&bE[q * 16 + i], wrong);
                cta_sync<THREADS>();
            }
        };

        // Optional stop rule (a.early_stop & 2; NOT in the reference, judged on FER only): after a sweep, decide u from
        // l(0,.)+r(0,.) and x from l(n,.)+r(n,.) and stop when u G = x, i.e. when the decisions form a codeword.  Costs the two
        // stage passes the plain sweep skips (l(0) and r(n)) plus a packed re-encode by one warp.
        auto gmatrix_ok = [&]() -> bool {
            if (tid < W) { uh[tid] = 0; xh[tid] = 0; }
            cta_sync<THREADS>();
            const real *rin = Rm + (size_t)(n - 2) * N;  // r(n-1,.)
#pragma unroll
            for (int i = 0; i < BPT; i++) {
                {   // u side: the positions 2q, 2q+1
                    const int q = tid + i * THREADS, j = 2 * q;
                    const real lu = Lm[j], ll = Lm[j + 1], ru = r0(j), rl = r0(j + 1);
                    const real ou = bchk<real>(lu, ll + rl), ol = ll + bchk<real>(ru, lu);
                    const uint32_t iu = (a.m.info[j >> 5] >> (j & 31)) & 1u, il = (a.m.info[j >> 5] >> ((j & 31) + 1)) & 1u;
                    const uint32_t two = ((iu && !(ou + ru >= (real)0)) ? 1u : 0u) | ((il && !(ol + rl >= (real)0)) ? 2u : 0u);
                    if (two) atomicOr(&uh[j >> 5], two << (j & 31));
                }
                {   // x side: the positions whose channel LLRs this thread holds
                    const int ju = (BPT == 2) ? (2 * tid + i) : (tid + i * THREADS), jl = ju + N / 2;
                    const real ru = rin[ju], rl = rin[jl];
                    const real su = ch_up[i] + bchk<real>(ru, ch_lo[i] + rl);
                    const real sl = ch_lo[i] + (rl + bchk<real>(ru, ch_up[i]));
                    if (!(su >= (real)0)) atomicOr(&xh[ju >> 5], 1u << (ju & 31));
                    if (!(sl >= (real)0)) atomicOr(&xh[jl >> 5], 1u << (jl & 31));
                }
            }
            cta_sync<THREADS>();
            if (tid < 32) {  // x' = u F^{(x)n} on packed words, one word per lane
                uint32_t w = (tid < W) ? polar_word_stages(uh[tid]) : 0u;
#pragma unroll
                for (int d = 1; d < W; d <<= 1) {
                    const uint32_t o = __shfl_xor_sync(0xffffffffu, w, d);
                    if (!(tid & d)) w ^= o;
                }
                const int bad = (tid < W) && (w != xh[tid]);
                const int any = __any_sync(0xffffffffu, bad);
                if (tid == 0) uh[W + 3] = any ? 0u : 1u;
            }
            cta_sync<THREADS>();
            return uh[W + 3] != 0u;
        };

        int sweeps = 0;
        for (int it = 0; it < a.iters; it++) {
            // ---- R pass, stages 0..n-2 (stage n-1 would only produce r(n), which nothing reads)
            if (BPT == 2) {
                // a thread owns two NEIGHBOURING butterflies (2*tid, 2*tid+1): for s >= 1 their upper nodes j, j+1 and their
                // lower nodes j+d, j+d+1 are adjacent, so every message pair moves as one 64-bit access
                {   // s = 0: butterflies (4t,4t+1) and (4t+2,4t+3)
                    const int j = 4 * tid;
                    real l0, l1, l2, l3;
                    LD(j, l0, l1);
                    LD(j + 2, l2, l3);
                    const real ra = r0(j), rb = r0(j + 1), rc = r0(j + 2), rd = r0(j + 3);
                    ST(RO + j, bchk<real>(ra, l1 + rb), rb + bchk<real>(ra, l0));
                    ST(RO + j + 2, bchk<real>(rc, l3 + rd), rd + bchk<real>(rc, l2));
                    cta_sync<THREADS>();
                }
#pragma unroll  // compile-time stage: strides, masks and array offsets become immediates (+4 % at N=1024; the CTAs of an SM run in
                // near lockstep, so the larger code costs nothing in instruction fetch)
                for (int s = 1; s < n - 1; s++) {
                    const int d = 1 << s, q = 2 * tid;
                    const int j = ((q >> s) << (s + 1)) | (q & (d - 1));
                    const int rin = RO + (s - 1) * N + j, lin = s * N + j, rout = RO + s * N + j;
                    real ru0, ru1, rl0, rl1, lu0, lu1, ll0, ll1;
                    LD(rin, ru0, ru1);
                    LD(rin + d, rl0, rl1);
                    LD(lin, lu0, lu1);
                    LD(lin + d, ll0, ll1);
                    ST(rout, bchk<real>(ru0, ll0 + rl0), bchk<real>(ru1, ll1 + rl1));
                    ST(rout + d, rl0 + bchk<real>(ru0, lu0), rl1 + bchk<real>(ru1, lu1));
                    cta_sync<THREADS>();
                }
            } else
#pragma unroll
            for (int s = 0; s < n - 1; s++) {
                const int d = 1 << s;
                const real *rin = Rm + (size_t)(s - 1) * N;
                const real *lin = Lm + (size_t)s * N;       // l(s+1,.)
                real *rout = Rm + (size_t)s * N;            // r(s+1,.)
#pragma unroll
                for (int i = 0; i < BPT; i++) {
                    const int q = tid + i * THREADS;
                    const int j = ((q >> s) << (s + 1)) | (q & (d - 1));
                    const real ru = (s == 0) ? r0(j) : rin[j];
                    const real rl = (s == 0) ? r0(j + d) : rin[j + d];
                    const real lu = lin[j], ll = lin[j + d];
                    rout[j] = bchk<real>(ru, ll + rl);
                    rout[j + d] = rl + bchk<real>(ru, lu);
                }
                cta_sync<THREADS>();
            }
            // ---- L pass, stages n-1..1
            int changed = 0;
            if (BPT == 2) {
#pragma unroll
                for (int s = n - 1; s >= 1; s--) {
                    const int d = 1 << s, q = 2 * tid;
                    const int j = ((q >> s) << (s + 1)) | (q & (d - 1));
                    const int rin = RO + (s - 1) * N + j, lin = s * N + j, lout = (s - 1) * N + j;
                    real ru0, ru1, rl0, rl1, lu0, lu1, ll0, ll1;
                    if (s == n - 1) { lu0 = ch_up[0]; lu1 = ch_up[1]; ll0 = ch_lo[0]; ll1 = ch_lo[1]; }
                    else { LD(lin, lu0, lu1); LD(lin + d, ll0, ll1); }
                    LD(rin, ru0, ru1);
                    LD(rin + d, rl0, rl1);
                    const real ou0 = bchk<real>(lu0, ll0 + rl0), ou1 = bchk<real>(lu1, ll1 + rl1);
                    const real ol0 = ll0 + bchk<real>(ru0, lu0), ol1 = ll1 + bchk<real>(ru1, lu1);
                    if (a.early_stop & 1) {
                        real pu0, pu1, pl0, pl1;
                        LD(lout, pu0, pu1);
                        LD(lout + d, pl0, pl1);
                        changed |= (int)(!RT::same_bits(ou0, pu0)) | (int)(!RT::same_bits(ou1, pu1)) | (int)(!RT::same_bits(ol0, pl0)) | (int)(!RT::same_bits(ol1, pl1));
                    }
                    ST(lout, ou0, ou1);
                    ST(lout + d, ol0, ol1);
                    cta_sync<THREADS>();
                }
            } else
#pragma unroll
            for (int s = n - 1; s >= 1; s--) {
                const int d = 1 << s;
                const real *rin = Rm + (size_t)(s - 1) * N;  // r(s,.)
                const real *lin = Lm + (size_t)s * N;        // l(s+1,.) (unused for s = n-1)
                real *lout = Lm + (size_t)(s - 1) * N;       // l(s,.)
#pragma unroll
                for (int i = 0; i < BPT; i++) {
                    const int q = tid + i * THREADS;
                    const int j = ((q >> s) << (s + 1)) | (q & (d - 1));
                    const real lu = (s == n - 1) ? ch_up[i] : lin[j];
                    const real ll = (s == n - 1) ? ch_lo[i] : lin[j + d];
                    const real ru = rin[j], rl = rin[j + d];
                    const real ou = bchk<real>(lu, ll + rl);
                    const real ol = ll + bchk<real>(ru, lu);
                    if (a.early_stop & 1) changed |= (int)(!RT::same_bits(ou, lout[j])) | (int)(!RT::same_bits(ol, lout[j + d]));
                    lout[j] = ou;
                    lout[j + d] = ol;
                }
                cta_sync<THREADS>();
            }
            sweeps = it + 1;
            if (a.bpr_ns && a.truth)
                for (int q = 0; q < a.bpr_ns; q++)
                    if (a.bpr_samples[q] == sweeps) bpr_sample(q);
            if (a.early_stop & 1) {
                const int any = (THREADS == 32) ? __any_sync(0xffffffffu, changed) : __syncthreads_or(changed);
                if (!any) break;
            }
            if ((a.early_stop & 2) && n >= 2 && gmatrix_ok()) break;
        }
        if (a.early_stop & 2) {  // the decision words are rebuilt below from the same state
            if (tid < W) uh[tid] = 0;
            cta_sync<THREADS>();
        }
        if (a.bpr_ns && a.truth) {  // samples scheduled after a fixed-point stop see the same, final, state
            for (int q = 0; q < a.bpr_ns; q++)
                if (a.bpr_samples[q] > sweeps && a.bpr_samples[q] <= a.iters) bpr_sample(q);
            cta_sync<THREADS>();
            for (int i = tid; i < 8 * 16; i += THREADS)
                if (bE[i]) atomicAdd(a.bpr_E + i, (unsigned long long)bE[i]);
        }
        // ---- l(0,.) from the final state and the decision (BP_1024.c:410-413,417-425)
        {
            const real *lin = Lm;  // l(1,.)
#pragma unroll
            for (int i = 0; i < BPT; i++) {
                const int q = tid + i * THREADS;
                const int j = 2 * q;
                const real lu = (n == 1) ? ch_up[i] : lin[j];
                const real ll = (n == 1) ? ch_lo[i] : lin[j + 1];
                const real ru = r0(j), rl = r0(j + 1);
                const real ou = bchk<real>(lu, ll + rl);
                const real ol = ll + bchk<real>(ru, lu);
                const uint32_t iu = (a.m.info[j >> 5] >> (j & 31)) & 1u, il = (a.m.info[j >> 5] >> ((j & 31) + 1)) & 1u;
                const uint32_t bu = (iu && !(ou + ru >= (real)0)) ? 1u : 0u;
                const uint32_t bl = (il && !(ol + rl >= (real)0)) ? 1u : 0u;
                const uint32_t two = bu | (bl << 1);
                if (two) atomicOr(&uh[j >> 5], two << (j & 31));
            }
        }
        cta_sync<THREADS>();
        if (tid < W) {
            const uint32_t w = uh[tid];
            if (a.u_hat) a.u_hat[frame * (size_t)W + tid] = w;
            if (a.truth) {
                const uint32_t e = __popc((w ^ __ldg(a.truth + frame * (size_t)W + tid)) & a.m.cnt[tid]);
                if (e) atomicAdd(&uh[W], e);
            }
        }
        cta_sync<THREADS>();
        if (tid == 0) {
            const uint32_t nerr = uh[W];
            if (a.frame_info) a.frame_info[frame] = (nerr > 0xFFFFu ? 0xFFFFu : nerr) | ((uint32_t)(sweeps > 255 ? 255 : sweeps) << 24);
            if (a.counters) {
                atomicAdd(a.counters + CNT_FRAMES, 1ull);
                if (nerr) { atomicAdd(a.counters + CNT_ERR_BLOCKS, 1ull); atomicAdd(a.counters + CNT_ERR_BITS, (unsigned long long)nerr); }
                atomicAdd(a.counters + CNT_SWEEPS, (unsigned long long)sweeps);
            }
        }
        cta_sync<THREADS>();
    }
}

template <typename real, int LOGN, int THREADS>
struct BpDispatch {
    using C = BpCfg<real, LOGN, THREADS>;
    static cudaError_t plan(BpPlan *p)
    {
        auto kern = bp_decode_kernel<real, LOGN, THREADS>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
        if (e != cudaSuccess) return e;
        int nb = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, THREADS, C::SMEM);
        if (e != cudaSuccess) return e;
        p->smem = C::SMEM;
        p->ctas_per_sm = nb;
        p->threads = THREADS;
        return cudaSuccess;
    }
    static cudaError_t launch(const BpArgs &a, int grid, cudaStream_t st)
    {
        const size_t smem = C::SMEM + (a.bpr_ns ? (size_t)C::N + 8 * 16 * 4 : 0);
        if (a.bpr_ns) {
            cudaError_t e = cudaFuncSetAttribute(bp_decode_kernel<real, LOGN, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        bp_decode_kernel<real, LOGN, THREADS><<<grid, THREADS, smem, st>>>(a);
        return cudaGetLastError();
    }
};

// threads per frame: fp32 N/4 (two butterflies per thread), fp64 N/2 for N=1024 (one frame per SM: use more threads)
#define POLAR_BP_CASES(X) X(6, 32, 32) X(7, 32, 32) X(8, 64, 64) X(9, 128, 128) X(10, 256, 512)

cudaError_t bp_plan(int n, bool f64, BpPlan *plan)
{
#define X(NN, T32, T64) \
    if (n == NN) return f64 ? BpDispatch<double, NN, T64>::plan(plan) : BpDispatch<float, NN, T32>::plan(plan);
    POLAR_BP_CASES(X)
#undef X
    return cudaErrorInvalidValue;
}

cudaError_t launch_bp(const BpArgs &a, int n, bool f64, int grid, cudaStream_t st)
{
#define X(NN, T32, T64) \
    if (n == NN) return f64 ? BpDispatch<double, NN, T64>::launch(a, grid, st) : BpDispatch<float, NN, T32>::launch(a, grid, st);
    POLAR_BP_CASES(X)
#undef X
    return cudaErrorInvalidValue;
}

}  // namespace polar
