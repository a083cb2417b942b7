// BP decoding with packed-half messages: TWO frames per CTA, one in each half of a __half2 (sm_100a).
//
// Optional throughput mode behind a flag (pg_params.real = PG_REAL_H2; BASELINE.json north_star item 4: "optionally as
// packed half2 ... behind a flag").  Same sweep, same butterfly rules and the same 8-level table as bp_decode.cu
// (/root/reference/BP_1024.c:372-427), but every add / min / compare / table step is one packed instruction for two
// frames, and a frame's messages take 36 KB instead of 72 KB.  It is a DIFFERENT decoder numerically (11-bit
// mantissa; table thresholds and values rounded to half; the frozen prior 999 plus a message loses its fraction), so
// it is never used for parity: it is judged on FER only (tests/test_gpu_parity.py::test_bp_half2_mode_fer).
#include <cuda_fp16.h>

#include "engine.h"
#include "polar_common.cuh"

namespace polar {

__device__ __forceinline__ __half2 h2c(float v) { return __float2half2_rn(v); }

// table on two packed non-negative values: running sum of increments over the thresholds above x
__device__ __forceinline__ __half2 tbl8_h2(__half2 x)
{
    __half2 t = __hmul2(__hlt2(x, h2c(4.5f)), h2c(0.05f));
    t = __hfma2(__hlt2(x, h2c(2.252f)), h2c(0.10f), t);
    t = __hfma2(__hlt2(x, h2c(1.508f)), h2c(0.10f), t);
    t = __hfma2(__hlt2(x, h2c(1.05f)), h2c(0.10f), t);
    t = __hfma2(__hlt2(x, h2c(0.71f)), h2c(0.10f), t);
    t = __hfma2(__hlt2(x, h2c(0.433f)), h2c(0.10f), t);
    t = __hfma2(__hlt2(x, h2c(0.196f)), h2c(0.10f), t);
    return t;
}

__device__ __forceinline__ __half2 chk_h2(__half2 a, __half2 b)
{
    const __half2 delta = __hsub2(tbl8_h2(__habs2(__hadd2(a, b))), tbl8_h2(__habs2(__hsub2(a, b))));
    const __half2 m = __hmin2(__habs2(a), __habs2(b));
    const uint32_t ua = *reinterpret_cast<const uint32_t *>(&a), ub = *reinterpret_cast<const uint32_t *>(&b);
    const uint32_t um = *reinterpret_cast<const uint32_t *>(&m) ^ ((ua ^ ub) & 0x80008000u);
    return __hadd2(*reinterpret_cast<const __half2 *>(&um), delta);
}

struct alignas(8) h2x2 { __half2 a, b; };  // two neighbouring messages (of two frames each): one 64-bit access
__device__ __forceinline__ void ld2h(const __half2 *p, __half2 &x, __half2 &y)
{
    const h2x2 v = *reinterpret_cast<const h2x2 *>(p);
    x = v.a; y = v.b;
}
__device__ __forceinline__ void st2h(__half2 *p, __half2 x, __half2 y)
{
    h2x2 v; v.a = x; v.b = y;
    *reinterpret_cast<h2x2 *>(p) = v;
}
// the same through 32-bit shared-window addresses (see bp_decode.cu: keeps the window base out of every stage)
__device__ __forceinline__ void lds2h(uint32_t a, __half2 &x, __half2 &y)
{
    uint32_t u, v;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(u), "=r"(v) : "r"(a) : "memory");
    x = *reinterpret_cast<const __half2 *>(&u); y = *reinterpret_cast<const __half2 *>(&v);
}
__device__ __forceinline__ void sts2h(uint32_t a, __half2 x, __half2 y)
{
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a), "r"(*reinterpret_cast<const uint32_t *>(&x)), "r"(*reinterpret_cast<const uint32_t *>(&y)) : "memory");
}
__device__ __forceinline__ uint32_t bits_of(__half2 v) { return *reinterpret_cast<const uint32_t *>(&v); }

template <int THREADS>
__device__ __forceinline__ void cta_sync_h2()
{
    if (THREADS == 32) __syncwarp();
    else __syncthreads();
}

template <int LOGN, int THREADS>
struct BpH2Cfg {
    static constexpr int N = 1 << LOGN;
    static constexpr int W = (N + 31) / 32;
    static constexpr size_t MSG = (size_t)2 * (LOGN - 1) * N;  // half2 elements: l(1..n-1), r(1..n-1) of two frames
    static constexpr size_t SMEM = MSG * sizeof(__half2) + (size_t)(2 * W + 8) * 4;
    static_assert((N / 2) / THREADS == 2, "the half2 kernel owns two neighbouring butterflies per thread");
};

template <int LOGN, int THREADS>
__global__ void __launch_bounds__(THREADS) bp_decode_h2_kernel(const BpArgs a)
{
    using C = BpH2Cfg<LOGN, THREADS>;
    constexpr int N = C::N, W = C::W, n = LOGN;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half2 *Lm = reinterpret_cast<__half2 *>(smem_raw);  // Lm[(s-1)*N + j] = l(s,j) of frames (A,B)
    // r(s,j) follows at Lm[(n-1)*N + (s-1)*N + j] (RO below)
    uint32_t *uh = reinterpret_cast<uint32_t *>(smem_raw + C::MSG * sizeof(__half2));  // [0,W) frame A, [W,2W) frame B, then scratch words
    const int tid = threadIdx.x;
    const unsigned long long pairs = (a.B + 1) / 2;
    uint32_t sa = (uint32_t)__cvta_generic_to_shared(Lm);  // message index i (l at i, r at RO + i) lives at sa + 4 i
    asm volatile("mov.b32 %0, %0;" : "+r"(sa));                 // opaque: stays in a register
    constexpr int RO = (LOGN - 1) * N;
    auto LD = [&](int i, __half2 &x, __half2 &y) { lds2h(sa + 4u * (uint32_t)i, x, y); };
    auto ST = [&](int i, __half2 x, __half2 y) { sts2h(sa + 4u * (uint32_t)i, x, y); };

    auto r0 = [&](int j) -> __half2 { return ((a.m.info[j >> 5] >> (j & 31)) & 1u) ? h2c(0.f) : h2c(999.f); };

    for (;;) {
        if (tid == 0) {
            const unsigned long long p = atomicAdd(a.queue, 1ull);
            uh[2 * W + 2] = (uint32_t)p;
            uh[2 * W + 3] = (uint32_t)(p >> 32);
        }
        cta_sync_h2<THREADS>();
        const unsigned long long pair = (unsigned long long)uh[2 * W + 2] | ((unsigned long long)uh[2 * W + 3] << 32);
        if (pair >= pairs) break;
        const unsigned long long fA = 2 * pair, fB = (2 * pair + 1 < a.B) ? 2 * pair + 1 : 2 * pair;
        const float *llrA = reinterpret_cast<const float *>(a.llr) + fA * (size_t)N, *llrB = reinterpret_cast<const float *>(a.llr) + fB * (size_t)N;

        __half2 ch_up[2], ch_lo[2];
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const int q = 2 * tid + i;
            ch_up[i] = __floats2half2_rn(__ldg(llrA + q), __ldg(llrB + q));
            ch_lo[i] = __floats2half2_rn(__ldg(llrA + q + N / 2), __ldg(llrB + q + N / 2));
        }
        for (int i = tid; i < (n - 1) * N; i += THREADS) Lm[i] = h2c(0.f);
        if (tid < 2 * W + 2) uh[tid] = 0;
        cta_sync_h2<THREADS>();

        int sweeps = 0;
        for (int it = 0; it < a.iters; it++) {
            {   // R pass, s = 0
                const int j = 4 * tid;
                __half2 l0, l1, l2, l3;
                LD(j, l0, l1);
                LD(j + 2, l2, l3);
                const __half2 ra = r0(j), rb = r0(j + 1), rc = r0(j + 2), rd = r0(j + 3);
                ST(RO + j, chk_h2(ra, __hadd2(l1, rb)), __hadd2(rb, chk_h2(ra, l0)));
                ST(RO + j + 2, chk_h2(rc, __hadd2(l3, rd)), __hadd2(rd, chk_h2(rc, l2)));
                cta_sync_h2<THREADS>();
            }
#pragma unroll
            for (int s = 1; s < n - 1; s++) {
                const int d = 1 << s, q = 2 * tid;
                const int j = ((q >> s) << (s + 1)) | (q & (d - 1));
                const int rin = RO + (s - 1) * N + j, lin = s * N + j, rout = RO + s * N + j;
                __half2 ru0, ru1, rl0, rl1, lu0, lu1, ll0, ll1;
                LD(rin, ru0, ru1);
                LD(rin + d, rl0, rl1);
                LD(lin, lu0, lu1);
                LD(lin + d, ll0, ll1);
                ST(rout, chk_h2(ru0, __hadd2(ll0, rl0)), chk_h2(ru1, __hadd2(ll1, rl1)));
                ST(rout + d, __hadd2(rl0, chk_h2(ru0, lu0)), __hadd2(rl1, chk_h2(ru1, lu1)));
                cta_sync_h2<THREADS>();
            }
            int changed = 0;
#pragma unroll
            for (int s = n - 1; s >= 1; s--) {
                const int d = 1 << s, q = 2 * tid;
                const int j = ((q >> s) << (s + 1)) | (q & (d - 1));
                const int rin = RO + (s - 1) * N + j, lin = s * N + j, lout = (s - 1) * N + j;
                __half2 ru0, ru1, rl0, rl1, lu0, lu1, ll0, ll1;
                if (s == n - 1) { lu0 = ch_up[0]; lu1 = ch_up[1]; ll0 = ch_lo[0]; ll1 = ch_lo[1]; }
                else { LD(lin, lu0, lu1); LD(lin + d, ll0, ll1); }
                LD(rin, ru0, ru1);
                LD(rin + d, rl0, rl1);
                const __half2 ou0 = chk_h2(lu0, __hadd2(ll0, rl0)), ou1 = chk_h2(lu1, __hadd2(ll1, rl1));
                const __half2 ol0 = __hadd2(ll0, chk_h2(ru0, lu0)), ol1 = __hadd2(ll1, chk_h2(ru1, lu1));
                if (a.early_stop) {
                    __half2 pu0, pu1, pl0, pl1;
                    LD(lout, pu0, pu1);
                    LD(lout + d, pl0, pl1);
                    changed |= (int)((bits_of(ou0) ^ bits_of(pu0)) | (bits_of(ou1) ^ bits_of(pu1)) | (bits_of(ol0) ^ bits_of(pl0)) | (bits_of(ol1) ^ bits_of(pl1))) != 0;
                }
                ST(lout, ou0, ou1);
                ST(lout + d, ol0, ol1);
                cta_sync_h2<THREADS>();
            }
            sweeps = it + 1;
            if (a.early_stop) {
                const int any = (THREADS == 32) ? __any_sync(0xffffffffu, changed) : __syncthreads_or(changed);
                if (!any) break;
            }
        }
        // l(0,.) and the decisions of both frames
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const int j = 2 * (2 * tid + i);
            __half2 lu, ll;
            ld2h(Lm + j, lu, ll);
            const __half2 ru = r0(j), rl = r0(j + 1);
            const __half2 su = __hadd2(chk_h2(lu, __hadd2(ll, rl)), ru);
            const __half2 sl = __hadd2(__hadd2(ll, chk_h2(ru, lu)), rl);
            const uint32_t iu = (a.m.info[j >> 5] >> (j & 31)) & 1u, il = (a.m.info[j >> 5] >> ((j & 31) + 1)) & 1u;
            const float2 fu = __half22float2(su), fl = __half22float2(sl);
            const uint32_t a2 = ((iu && !(fu.x >= 0.f)) ? 1u : 0u) | ((il && !(fl.x >= 0.f)) ? 2u : 0u);
            const uint32_t b2 = ((iu && !(fu.y >= 0.f)) ? 1u : 0u) | ((il && !(fl.y >= 0.f)) ? 2u : 0u);
            if (a2) atomicOr(&uh[j >> 5], a2 << (j & 31));
            if (b2) atomicOr(&uh[W + (j >> 5)], b2 << (j & 31));
        }
        cta_sync_h2<THREADS>();
        if (tid < 2 * W) {
            const int which = tid / W, w = tid - which * W;
            const unsigned long long fr = which ? (2 * pair + 1) : fA;
            if (fr < a.B) {
                const uint32_t v = uh[tid];
                if (a.u_hat) a.u_hat[fr * (size_t)W + w] = v;
                if (a.truth) {
                    const uint32_t e = __popc((v ^ __ldg(a.truth + fr * (size_t)W + w)) & a.m.cnt[w]);
                    if (e) atomicAdd(&uh[2 * W + which], e);
                }
            }
        }
        cta_sync_h2<THREADS>();
        if (tid < 2) {
            const unsigned long long fr = 2 * pair + tid;
            if (fr < a.B) {
                const uint32_t nerr = uh[2 * W + tid];
                if (a.frame_info) a.frame_info[fr] = (nerr > 0xFFFFu ? 0xFFFFu : nerr) | ((uint32_t)(sweeps > 255 ? 255 : sweeps) << 24);
                if (a.counters) {
                    atomicAdd(a.counters + CNT_FRAMES, 1ull);
                    if (nerr) { atomicAdd(a.counters + CNT_ERR_BLOCKS, 1ull); atomicAdd(a.counters + CNT_ERR_BITS, (unsigned long long)nerr); }
                    atomicAdd(a.counters + CNT_SWEEPS, (unsigned long long)sweeps);
                }
            }
        }
        cta_sync_h2<THREADS>();
    }
}

template <int LOGN, int THREADS>
struct BpH2Dispatch {
    using C = BpH2Cfg<LOGN, THREADS>;
    static cudaError_t plan(BpPlan *p)
    {
        auto kern = bp_decode_h2_kernel<LOGN, THREADS>;
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
        bp_decode_h2_kernel<LOGN, THREADS><<<grid, THREADS, C::SMEM, st>>>(a);
        return cudaGetLastError();
    }
};

#define POLAR_BPH2_CASES(X) X(7, 32) X(8, 64) X(9, 128) X(10, 256)

cudaError_t bp_h2_plan(int n, BpPlan *plan)
{
#define X(NN, T) \
    if (n == NN) return BpH2Dispatch<NN, T>::plan(plan);
    POLAR_BPH2_CASES(X)
#undef X
    return cudaErrorInvalidValue;
}

cudaError_t launch_bp_h2(const BpArgs &a, int n, int grid, cudaStream_t st)
{
#define X(NN, T) \
    if (n == NN) return BpH2Dispatch<NN, T>::launch(a, grid, st);
    POLAR_BPH2_CASES(X)
#undef X
    return cudaErrorInvalidValue;
}

}  // namespace polar
