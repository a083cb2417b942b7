// SC / SC-List / CRC-aided SC-List decoding, one PATH per lane (sm_100a).
//
// What it computes (reference): SCdecode /root/reference/SC_128.c:395-460, SCLdecode
// /root/reference/SCL_1024.c:547-680, CASCL /root/reference/CASCL_1024_L8.c:601-761, with
// CHK (SC_128.c:284-315), the g function (SC_128.c:355-359), PHI (SCL_1024.c:481-502), the
// "< med" survivor rule and slot re-use order (SCL_1024.c:619-661) and CRcheck
// (CASCL_1024_L8.c:569-598).  How: nothing like the reference's pointer graph.
//
//  * A warp decodes 32/L frames side by side; lane = (frame, path slot k).  All lanes run the SAME
//    instruction stream (the SC schedule does not depend on data), so there is no divergence, and
//    every f/g evaluation of a path is lane-private: no barrier inside a path, ILP from the 4-wide
//    unrolled layers.  Slots that hold no path yet (list filling) carry PM=+inf and are exact copies
//    of path 0, so they compute finite values and never win a comparison.
//  * Array formulation: at bit j, t=ctz(j): one g-layer at stage t then f-layers t-1..0; stage s keeps
//    2^s live LLRs.  Stages 0..2 live in registers only (stage 2 feeds the four leaves of its group and
//    is cloned by shuffle), stage 3 goes through registers to the f step below it and is also stored,
//    stages 3..SMEM_TOP-1 are in shared memory, the rest in an L2/HBM global scratch, all laid out
//    [idx/4][lane][4] so that a warp's 128-bit accesses are contiguous/conflict-free.
//  * Lazy copy: every path owns a HOME array per stage and a packed pointer word saying where its
//    current stage-s data lives.  Stage s is rewritten by all paths at the same bits (multiples of
//    2^s), always into the home array, so a clone is a register shuffle of the pointer word -- the
//    reference copies the whole graph instead (copyPath/simpleCopy, 74 % of its run time).
//    Only the first layer of a chain (the g-layer) follows a pointer; the f-layers below it read what
//    their own lane just wrote, and the pointer fields of the whole chain are merged in once at its end.
//  * Partial sums are kept as packed bit vectors per stage (B[s], 2^s bits) with the same pointer
//    scheme; stages 2..5 are registers, 6..BITS_TOP-1 shared memory, the rest global scratch.
//    The final B[n] is the re-encoded codeword, u_hat = B[n] F^{(x)n}.
//  * The all-frozen prefix (one path, every decision 0) has no data dependence: the frame's L lanes
//    evaluate it as a parallel butterfly and add the leaf penalties in leaf order (see "the frozen prefix").
//  * Code size matters: the warps of an SM sit at unrelated program counters.  The first version
//    (64-90 KB of SASS, everything unrolled) spent most issue slots waiting for instruction fetch
//    (ncu: stall_no_instruction 4.0 per issue, profiles/r1_cascl_v1_summary.txt).  The hot loop is a few
//    looped bodies (generic f/g layers, the stage-3/2 steps, two leaf bodies): ~50 KB, stall 0.5-1.2.
//  * List pruning: lo = L-th and hi = (L+1)-th smallest of the frame's 2L candidates from a bitonic
//    merge network on shuffles (see leaf()).  If lo < hi in every frame of the warp, "PM < hi" is exactly
//    the reference's "PM < med"; otherwise (exact ties, or +inf slots while the list fills) a slow
//    path applies the total order (value, candidate index) and flags frames where the reference's
//    rule would have been ambiguous ("Oops!", SCL_1024.c:621).
#include "engine.h"
#include "polar_common.cuh"
#ifndef POLAR_EMU
#include <algorithm>
#endif

namespace polar {

template <typename real> struct alignas(16) vec4 { real v[4]; };

template <int L> struct ptr_word { using type = uint32_t; static constexpr int W = 4; };
template <> struct ptr_word<32> { using type = unsigned long long; static constexpr int W = 5; };

// TMH > TML: the LLR stages TML..TMH-1 live in TENSOR MEMORY (sm_100a: 128 lanes x 512 columns of 32 bits per SM, idle in a
// kernel without tensor-core work): a warp owns the 32 TMEM lanes of its quarter, lane = path as everywhere else, column = index
// within the stages.  CTAs then hold four independent warps (one per TMEM lane quarter) and allocate the columns once.
// Stages 3..TML-1 are shared memory, TMH.. the global scratch.
template <typename real, int LOGN, int L, int SMEM_TOP, int BITS_TOP, int TML = 0, int TMH = 0>
struct ListCfg {
    static constexpr int N = 1 << LOGN;
    static constexpr int W = (N + 31) / 32;
    static constexpr int FPW = 32 / L;
    static constexpr bool HAS_TM = TMH > TML;
    static_assert(!HAS_TM || (TML >= 3 && TMH <= LOGN - 1), "tensor-memory stages: above stage 2, the first scratch stage below the channel");
    static constexpr int TOP = HAS_TM ? TML : ((SMEM_TOP < LOGN) ? SMEM_TOP : LOGN);  // LLR stages 3..TOP-1 in smem
    static constexpr int GLO = HAS_TM ? TMH : TOP;                             // first LLR stage in the global scratch
    static constexpr int WARPS = HAS_TM ? 4 : 1;                               // warps per CTA (independent of each other)
    static constexpr int THREADS = 32 * WARPS;
    static constexpr int TM_USED = HAS_TM ? ((1 << TMH) - (1 << TML)) * (int)(sizeof(real) / 4) : 0;
    static constexpr int TM_COLS = !HAS_TM ? 0 : (TM_USED <= 32 ? 32 : TM_USED <= 64 ? 64 : TM_USED <= 128 ? 128 : TM_USED <= 256 ? 256 : 512);
    static_assert(TM_USED <= 512, "tensor-memory allocation: at most 512 columns");
    static constexpr int BTOP = (BITS_TOP < LOGN + 1) ? BITS_TOP : LOGN + 1;    // bit stages 6..BTOP-1 in smem
    static constexpr int BLO = (BTOP > 6) ? BTOP : 6;                           // first bit stage in global scratch
    static constexpr int SM_STAGE_REALS = 32 * ((1 << TOP) - 8);          // stages 3..TOP-1 (stage 2 lives in registers only)
    static constexpr int SM_BIT_WORDS = (BTOP > 6) ? 32 * ((1 << (BTOP - 5)) - 2) : 0;
    static constexpr int SM_SLOT_WORDS = (L > 1) ? 32 : 0;                      // clone-source table, one word per lane
    static constexpr size_t SMEM = (size_t)SM_STAGE_REALS * sizeof(real) + (size_t)SM_BIT_WORDS * 4 + (size_t)SM_SLOT_WORDS * 4;  // per warp
    static constexpr size_t SMEM_CTA = SMEM * WARPS + (HAS_TM ? 16 : 0);       // + the word tcgen05.alloc writes
    static constexpr size_t GS_REALS = 32 * (size_t)((1 << LOGN) - (1 << GLO));                    // stages GLO..LOGN-1
    static constexpr size_t GS_BIT_WORDS = (LOGN >= BLO) ? 32 * (size_t)((1 << (LOGN - 4)) - (1 << (BLO - 5))) : 0;  // BLO..LOGN
    static constexpr size_t GS_BYTES = GS_REALS * sizeof(real) + GS_BIT_WORDS * 4;                 // per warp
};

// explicit 128-bit accesses (the compiler otherwise splits some of these into scalar loads)
__device__ __forceinline__ vec4<float> ldv(const vec4<float> *p)
{
    const float4 t = *reinterpret_cast<const float4 *>(p);
    vec4<float> o; o.v[0] = t.x; o.v[1] = t.y; o.v[2] = t.z; o.v[3] = t.w;
    return o;
}
__device__ __forceinline__ vec4<double> ldv(const vec4<double> *p)
{
    const double2 t = reinterpret_cast<const double2 *>(p)[0], u = reinterpret_cast<const double2 *>(p)[1];
    vec4<double> o; o.v[0] = t.x; o.v[1] = t.y; o.v[2] = u.x; o.v[3] = u.y;
    return o;
}
__device__ __forceinline__ void stv(vec4<float> *p, const vec4<float> &o)
{
    *reinterpret_cast<float4 *>(p) = make_float4(o.v[0], o.v[1], o.v[2], o.v[3]);
}
__device__ __forceinline__ void stv(vec4<double> *p, const vec4<double> &o)
{
    reinterpret_cast<double2 *>(p)[0] = make_double2(o.v[0], o.v[1]);
    reinterpret_cast<double2 *>(p)[1] = make_double2(o.v[2], o.v[3]);
}
__device__ __forceinline__ float rmax(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double rmax(double a, double b) { return fmax(a, b); }

// ---- tensor memory: one V4 (four reals of one path) = 4 (float) or 8 (double) consecutive columns of the lane's TMEM row.
// tcgen05.ld / tcgen05.st are asynchronous: tm_wait_ld() before the loaded registers are read, tm_wait_st() before the stored
// columns are loaded again.  `tm` = allocation base + (32 * (warp % 4)) << 16; `v4` = index of the V4 within the stage.
#ifdef POLAR_EMU
__device__ __forceinline__ uint32_t tm_alloc(uint32_t *slot, int cols, int warp) { if (warp == 0 && emu::lane_id() == 0) *slot = emu::tmem_alloc((uint32_t)cols); return *slot; }
__device__ __forceinline__ void tm_free(uint32_t, int) {}
__device__ __forceinline__ void tm_wait_ld() {}
__device__ __forceinline__ void tm_wait_st() {}
template <typename real> __device__ __forceinline__ vec4<real> tm_ld(uint32_t tm, int v4)
{
    vec4<real> o;
    constexpr int C = (int)(sizeof(vec4<real>) / 4);
    uint32_t w[C];
    for (int e = 0; e < C; e++) w[e] = *emu::tmem_cell(tm + (uint32_t)(v4 * C), e);
    memcpy(&o, w, sizeof(o));
    return o;
}
template <typename real> __device__ __forceinline__ void tm_st(uint32_t tm, int v4, const vec4<real> &o)
{
    constexpr int C = (int)(sizeof(vec4<real>) / 4);
    uint32_t w[C];
    memcpy(w, &o, sizeof(o));
    for (int e = 0; e < C; e++) *emu::tmem_cell(tm + (uint32_t)(v4 * C), e) = w[e];
}
#else
__device__ __forceinline__ uint32_t tm_alloc(uint32_t *slot, int cols, int warp)  // all threads of the CTA call this
{
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(slot)), "r"(cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    return *reinterpret_cast<volatile uint32_t *>(slot);
}
__device__ __forceinline__ void tm_free(uint32_t base, int cols)  // warp 0, after a __syncthreads()
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
template <typename real> __device__ __forceinline__ vec4<real> tm_ld(uint32_t tm, int v4);
template <> __device__ __forceinline__ vec4<float> tm_ld<float>(uint32_t tm, int v4)
{
    uint32_t a, b, c, d;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(tm + (uint32_t)(v4 * 4)) : "memory");
    vec4<float> o;
    o.v[0] = __uint_as_float(a); o.v[1] = __uint_as_float(b); o.v[2] = __uint_as_float(c); o.v[3] = __uint_as_float(d);
    return o;
}
template <> __device__ __forceinline__ vec4<double> tm_ld<double>(uint32_t tm, int v4)
{
    uint32_t w[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "r"(tm + (uint32_t)(v4 * 8)) : "memory");
    vec4<double> o;
#pragma unroll
    for (int e = 0; e < 4; e++) o.v[e] = __hiloint2double((int)w[2 * e + 1], (int)w[2 * e]);
    return o;
}
__device__ __forceinline__ void tm_st(uint32_t tm, int v4, const vec4<float> &o)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(tm + (uint32_t)(v4 * 4)), "r"(__float_as_uint(o.v[0])),
                 "r"(__float_as_uint(o.v[1])), "r"(__float_as_uint(o.v[2])), "r"(__float_as_uint(o.v[3])) : "memory");
}
__device__ __forceinline__ void tm_st(uint32_t tm, int v4, const vec4<double> &o)
{
    uint32_t w[8];
#pragma unroll
    for (int e = 0; e < 4; e++) { w[2 * e] = (uint32_t)__double2loint(o.v[e]); w[2 * e + 1] = (uint32_t)__double2hiint(o.v[e]); }
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tm + (uint32_t)(v4 * 8)), "r"(w[0]), "r"(w[1]), "r"(w[2]),
                 "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
}
#endif
// L1 prefetch of one 16-byte element per lane (no register, no scoreboard)
__device__ __forceinline__ void prefetch_l1(const void *p)
{
#ifndef POLAR_EMU
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}
// Two more ways to keep the f-layer from re-reading a g-layer's output out of HBM were measured and are slower (CA-SCL 1024 L=8, fp32
// M frames/s, product 13.5): the g-layer fused with the f-layer below it (registers instead of a scratch round trip; 12.7, +6 %
// instructions; removed again, see profiles/r2_list_kernel_experiments.md) and the chunked chain below (POLAR_CHUNK: 12.7, HBM reads -13 %,
// long_scoreboard 3.0 -> 2.3 per issue, +7 % instructions).  The issue slots are worth more than the stalls.
#ifndef POLAR_CHUNK
#define POLAR_CHUNK 0        // > 0 (a multiple of 8): g-layers at stages >= POLAR_CHUNK_MIN are produced this many rows per half at a time, each
#endif                       // chunk consumed at once by the f-layer below (chunked chain, see the main loop)
#ifndef POLAR_CHUNK_MIN
#define POLAR_CHUNK_MIN 7
#endif
#ifndef POLAR_F32_CTAS
#define POLAR_F32_CTAS 32   // resident one-warp CTAs per SM the fp32 kernel is compiled for (64 registers per thread)
#endif
// fewer inline copies of the leaf body / of the four-CHK block in the hot loop (instruction cache): 0 never, 1 always, 2 fp64 only
#ifndef POLAR_FOLD_LEAF
#define POLAR_FOLD_LEAF 2
#endif
#ifndef POLAR_FOLD_F3
#define POLAR_FOLD_F3 1
#endif
#ifndef POLAR_FRZ4
#define POLAR_FRZ4 1     // straight-line body for leaf groups of four frozen bits
#endif
// Measured on B200 (CA-SCL 1024 L=8; fp32 / fp64 M frames/s, tools/ab.py): round-1 loop 13.44 / 5.03; POLAR_COMPACT 13.46 / 6.06 (the
// fp64 instantiation, whose CHK is twice as long, comes back under the instruction-cache cliff); POLAR_COMPACT + POLAR_VIRT 11.1 /
// 5.65 although the HBM traffic drops by 14 % and long_scoreboard from 3.0 to 2.0 per issue: the two extra loops push the fp32
// kernel from 3 050 to 3 900 SASS instructions, over the same cliff (3 740 without the frozen-group body: 12.2).  So POLAR_VIRT is
// off in the product; the CPU emulator build turns it on to keep its logic pinned to the oracle.
#ifndef POLAR_VIRT
#define POLAR_VIRT 0     // 1: two g-outputs of the top LLR stages are recomputed instead of stored (see f_virtual)
#endif
#ifndef POLAR_PF_MAX
#define POLAR_PF_MAX 0   // g-layers up to this stage get their source prefetched into the L1 one leaf group ahead (0: off)
#endif
// a V4 as seen by lane `src` (clone-pointer indirection for tensor memory, where a lane reaches its own row only)
template <typename real> __device__ __forceinline__ vec4<real> shfl_v4(const vec4<real> &v, int src)
{
    vec4<real> o;
#pragma unroll
    for (int e = 0; e < 4; e++) o.v[e] = __shfl_sync(0xffffffffu, v.v[e], src);
    return o;
}

// lo <- max of lo, hi <- min of hi over the L lanes of a frame (mask = those lanes).  Path metrics are non-negative (or +inf), so
// in fp32 their bit patterns order like unsigned integers and one REDUX per value replaces log2(L) shuffle + min/max steps.
template <int L>
__device__ __forceinline__ void frame_minmax(float &lo, float &hi, uint32_t mask)
{
#ifdef POLAR_REDUX
    lo = __uint_as_float(__reduce_max_sync(mask, __float_as_uint(lo)));
    hi = __uint_as_float(__reduce_min_sync(mask, __float_as_uint(hi)));
#else
#pragma unroll
    for (int d = 1; d < L; d <<= 1) {
        lo = fmaxf(lo, __shfl_xor_sync(0xffffffffu, lo, d));
        hi = fminf(hi, __shfl_xor_sync(0xffffffffu, hi, d));
    }
#endif
}
template <int L>
__device__ __forceinline__ void frame_minmax(double &lo, double &hi, uint32_t)
{
#pragma unroll
    for (int d = 1; d < L; d <<= 1) {
        lo = fmax(lo, __shfl_xor_sync(0xffffffffu, lo, d));
        hi = fmin(hi, __shfl_xor_sync(0xffffffffu, hi, d));
    }
}

// one CHK / g evaluation of four neighbouring nodes
// (Inline on purpose.  As a called function the four-CHK block costs eight argument moves and spills around every call:
// 13.4 -> 11.5 M frames/s.  What matters instead is the NUMBER of inline copies: the warps of an SM sit at unrelated program
// counters, and 98 KB of SASS instead of 68 KB raised stall_no_instruction from 8 % to 44 % of the samples -- so every f-layer
// of the hot loop goes through ONE loop body whose loads and stores switch on warp-uniform flags.)
template <typename real>
__device__ __forceinline__ vec4<real> f4(const vec4<real> &x, const vec4<real> &y)
{
    vec4<real> o;
#pragma unroll
    for (int e = 0; e < 4; e++) o.v[e] = chk_lean<real>(x.v[e], y.v[e]);
    return o;
}
template <typename real>
__device__ __forceinline__ vec4<real> g4(const vec4<real> &up, const vec4<real> &lo, uint32_t b)
{
    vec4<real> o;
#pragma unroll
    for (int e = 0; e < 4; e++) o.v[e] = lo.v[e] + real_traits<real>::flip(up.v[e], (b >> e) & 1u);  // SC_128.c:355-359
    return o;
}

template <typename real, int LOGN, int L, int SMEM_TOP, int BITS_TOP, int TML, int TMH>
__global__ void __launch_bounds__(ListCfg<real, LOGN, L, SMEM_TOP, BITS_TOP, TML, TMH>::THREADS,
                                  (TMH > TML) ? ((TML > 3) ? 7 : 8) : ((sizeof(real) == 4) ? POLAR_F32_CTAS : 24))
list_decode_kernel(const ListArgs a)
{
    using C = ListCfg<real, LOGN, L, SMEM_TOP, BITS_TOP, TML, TMH>;
    using RT = real_traits<real>;
    using PW = ptr_word<L>;
    using ptr_t = typename PW::type;
    using V4 = vec4<real>;
    constexpr int N = C::N, W = C::W, FPW = C::FPW, TOP = C::TOP, BTOP = C::BTOP, BLO = C::BLO, GLO = C::GLO;
    constexpr bool HAS_TM = C::HAS_TM;
    constexpr bool FOLD_LEAF = POLAR_FOLD_LEAF == 1 || (POLAR_FOLD_LEAF == 2 && sizeof(real) == 8);
    constexpr bool FOLD_F3 = POLAR_FOLD_F3 == 1 || (POLAR_FOLD_F3 == 2 && sizeof(real) == 8);
    constexpr int PWID = PW::W;
    constexpr ptr_t PMASK = (ptr_t)((1u << PWID) - 1);
    constexpr uint32_t LMASK = (L == 32) ? 0xffffffffu : ((1u << (L & 31)) - 1u);  // lanes of one frame
    const real INF = RT::inf();

#ifdef POLAR_EMU
    unsigned char *const smem_cta = emu::g.smem;
#else
    extern __shared__ __align__(16) unsigned char smem_cta[];
#endif
    // the warps of a CTA are independent decoders: each has its own shared-memory block, scratch block and TMEM lane quarter
    const int wi = (C::WARPS > 1) ? (int)(threadIdx.x >> 5) : 0;
    const unsigned gw = blockIdx.x * C::WARPS + wi, nwarps = gridDim.x * C::WARPS;
    unsigned char *const smem_raw = smem_cta + (size_t)wi * C::SMEM;
    V4 *const sm_stage = reinterpret_cast<V4 *>(smem_raw);  // stage s (3<=s<TOP): group i4 of lane pl at [8*(2^s-8) + i4*32 + pl]
    uint32_t *const sm_bits = reinterpret_cast<uint32_t *>(smem_raw + (size_t)C::SM_STAGE_REALS * sizeof(real));
    uint32_t *const sm_slot = sm_bits + C::SM_BIT_WORDS;  // [32]: lane id of the t-th both-survivor of each frame
    unsigned char *const gs_raw = reinterpret_cast<unsigned char *>(a.gscratch) + (size_t)gw * C::GS_BYTES;
    V4 *const gs_stage = reinterpret_cast<V4 *>(gs_raw);
    uint32_t *const gs_bits = reinterpret_cast<uint32_t *>(gs_raw + C::GS_REALS * sizeof(real));
    uint32_t tm_base = 0, tm = 0;  // tensor memory: allocation of the CTA / this warp's lane quarter of it
    if (HAS_TM) {
        tm_base = tm_alloc(reinterpret_cast<uint32_t *>(smem_cta + C::SMEM * C::WARPS), C::TM_COLS, wi);
        tm = tm_base + ((uint32_t)(wi & 3) << 21);  // lane field (bits 16..) = 32 * (warp % 4)
    }

    const int lane = (C::WARPS > 1) ? (int)(threadIdx.x & 31) : (int)threadIdx.x;
    const int k = lane & (L - 1);
    const int fbase = lane - k;
    const int fl = lane / L;
    const unsigned long long groups = (a.B + FPW - 1) / FPW;
    ptr_t kpat = 0;  // k in every pointer field
#pragma unroll
    for (int i = 0; i * PWID < (int)(8 * sizeof(ptr_t)) - PWID + 1; i++) kpat |= (ptr_t)k << (i * PWID);

    // home arrays (generic pointers: one code path for shared and global stages keeps the hot loop small)
    auto stage_at = [&](int s) -> V4 * {  // (never called for a tensor-memory stage)
        return (s < TOP) ? (sm_stage + 8 * ((1 << s) - 8)) : (gs_stage + 8 * (size_t)((1 << s) - (1 << GLO)));
    };
    auto in_tm = [&](int s) -> bool { return HAS_TM && s >= TML && s < TMH; };
    auto tmoff = [&](int s) -> int { return ((1 << s) - (1 << (HAS_TM ? TML : 0))) >> 2; };  // first V4 of stage s in the lane's TMEM row
    // bit array of stage s (s>=6): word w of physical lane pl at [w*32 + pl]
    auto bits_at = [&](int s) -> uint32_t * {
        return (s < BTOP) ? (sm_bits + 32 * ((1 << (s - 5)) - 2)) : (gs_bits + 32 * (size_t)((1 << (s - 5)) - (1 << (BLO - 5))));
    };

    for (unsigned long long g = gw; g < groups; g += nwarps) {
        unsigned long long frame = g * FPW + fl;
        const bool valid = frame < a.B;
        if (!valid) frame = a.B - 1;  // tail lanes decode a duplicate and write nothing
        const V4 *const ch4 = reinterpret_cast<const V4 *>(reinterpret_cast<const real *>(a.llr) + frame * (size_t)N);

        real s2[4] = {0, 0, 0, 0}, s1[2] = {0, 0}, pm;
        ptr_t ptr = 0, bptr = 0;  // fields: LLR stage s at (s-2)*PWID / bit stage s at (s-6)*PWID
        uint32_t Blow = 0, B5 = 0, ug = 0, flags = 0;
        pm = (k == 0 || L == 1) ? (real)0 : INF;

        auto pfield = [&](int s) -> int { return (L == 1) ? 0 : (int)((ptr >> ((s - 2) * PWID)) & PMASK); };
        auto bfield = [&](int s) -> int { return (L == 1) ? 0 : (int)((bptr >> ((s - 6) * PWID)) & PMASK); };
        // after a chain of layers the stages 3..top live in this lane's home arrays: one masked merge of k into those fields
        auto set_pfields = [&](int top) {
            if (L > 1) {
                const ptr_t fm = (~(ptr_t)0 >> ((int)(8 * sizeof(ptr_t)) - (top - 1) * PWID)) & ~PMASK;
                ptr = (ptr & ~fm) | (kpat & fm);
            }
        };
        auto set_bfield = [&](int s) {
            if (L > 1) bptr = (bptr & ~(PMASK << ((s - 6) * PWID))) | ((ptr_t)k << ((s - 6) * PWID));
        };

        // ---- f-layer producing stage s (4 <= s < LOGN) from stage s+1, into the HOME array -----------------
        // An f-layer always follows the layer that produced stage s+1 in the same chain, so its source is this lane's own
        // home array (or the channel): no pointer lookup.  Pointer fields are set once per chain (set_pfields).
        auto f_layer = [&](int s, bool coop, int r0 = 0, int rn = -1) {   // rows r0 .. r0+rn-1 of the output (default: all)
            const int cnt4 = 1 << (s - 2);
            if (coop) {
                // all lanes of the frame still hold the same path (no information bit yet): they split the layer and
                // write ONE array, slot 0's home; every pointer field of a stage >= 4 still says slot 0
                V4 *dst = stage_at(s) + fbase;
                const V4 *src = ch4;
                int stride = 1;
                if (s + 1 != LOGN) { src = stage_at(s + 1) + fbase; stride = 32; }
#pragma unroll 1
                for (int i4 = k; i4 < cnt4; i4 += L) stv(dst + i4 * 32, f4<real>(ldv(src + i4 * stride), ldv(src + (i4 + cnt4) * stride)));
                __syncwarp();
                return;
            }
            // one loop body for every storage class (see f4): the source is this lane's home array in the scratch / shared memory,
            // the channel, or its own tensor-memory row; the destination its home array or its tensor-memory row
            const bool tsrc = in_tm(s + 1), tdst = in_tm(s);
            const int osrc = tmoff(s + 1) + r0, odst = tmoff(s) + r0;
            V4 *dst = tdst ? nullptr : stage_at(s) + lane + r0 * 32;
            const V4 *src = ch4;
            int stride = 1;
            if (!tsrc && s + 1 != LOGN) { src = stage_at(s + 1) + lane; stride = 32; }
            const V4 *src2 = src + (cnt4 + r0) * stride;   // the row range is folded into the bases: the loop below counts from 0
            src += r0 * stride;
            auto load2 = [&](int i4, V4 &x, V4 &y) {
                if (tsrc) { x = tm_ld<real>(tm, osrc + i4); y = tm_ld<real>(tm, osrc + i4 + cnt4); tm_wait_ld(); }
                else { x = ldv(src + i4 * stride); y = ldv(src2 + i4 * stride); }
            };
            auto store = [&](int i4, const V4 &v) {
                if (tdst) tm_st(tm, odst + i4, v);
                else stv(dst + i4 * 32, v);
            };
            // scratch / channel operands come from L2 or HBM: fetch the next pair while the current four CHKs run
            V4 x, y, x1, y1;
            const int rows = (rn < 0) ? cnt4 : rn;
            load2(0, x, y);
#pragma unroll 1
            for (int i4 = 0; i4 < rows; i4 += 2) {  // an even number of rows; two steps per trip so that the operand registers ping-pong
                load2(i4 + 1, x1, y1);
                store(i4, f4<real>(x, y));
                if (i4 + 2 < rows) load2(i4 + 2, x, y);
                store(i4 + 1, f4<real>(x1, y1));
            }
            if (tdst) tm_wait_st();
        };

        // ---- g-layer producing stage t (4 <= t < LOGN) from stage t+1 (via the pointer word) and the partial sums B[t]
        auto g_layer = [&](int t, int r0 = 0, int rn = -1) {   // rows r0 .. r0+rn-1 of the output (a multiple of 8, r0 too; default: all)
            const int cnt4 = 1 << (t - 2);
            const uint32_t *bsrc = (t >= 6) ? bits_at(t) + fbase + bfield(t) + (r0 >> 3) * 32 : nullptr;
            uint32_t bw = (t == 4) ? ((Blow >> 12) & 0xFFFFu) : B5;
            // one loop body for every storage class.  Tensor-memory source: a lane reaches its own row only, so every lane loads
            // its OWN row and the values of the slot the pointer word names arrive by shuffle.
            const bool tsrc = in_tm(t + 1), tdst = in_tm(t);
            const int osrc = tmoff(t + 1) + r0, odst = tmoff(t) + r0;   // the row range is folded into the bases
            const int sl = fbase + pfield(t + 1);
            V4 *dst = tdst ? nullptr : stage_at(t) + lane + r0 * 32;
            const V4 *src = ch4;
            int stride = 1;
            if (!tsrc && t + 1 != LOGN) { src = stage_at(t + 1) + fbase + pfield(t + 1); stride = 32; }
            const V4 *src2 = src + (cnt4 + r0) * stride;
            src += r0 * stride;
            const int bend = ((rn < 0) ? cnt4 : rn) >> 2;
            // a g-layer is one add per node: memory bound (its operands come from the L2/HBM scratch).  Batches of four
            // node groups: eight independent 128-bit loads in flight per lane, 16 partial-sum bits per batch.
#pragma unroll 1
            for (int b = 0; b < bend; b++) {
                if (t >= 6 && !(b & 1)) { bw = *bsrc; bsrc += 32; }
                V4 up[4], lo[4];
                if (tsrc) {
#pragma unroll
                    for (int q = 0; q < 4; q++) { up[q] = tm_ld<real>(tm, osrc + 4 * b + q); lo[q] = tm_ld<real>(tm, osrc + 4 * b + q + cnt4); }
                    tm_wait_ld();
#pragma unroll
                    for (int q = 0; q < 4; q++) if (L > 1) { up[q] = shfl_v4<real>(up[q], sl); lo[q] = shfl_v4<real>(lo[q], sl); }
                } else {
#pragma unroll
                    for (int q = 0; q < 4; q++) { up[q] = ldv(src + q * stride); lo[q] = ldv(src2 + q * stride); }
                    src += 4 * stride; src2 += 4 * stride;
                }
                if (tdst) {
#pragma unroll
                    for (int q = 0; q < 4; q++) tm_st(tm, odst + 4 * b + q, g4<real>(up[q], lo[q], (bw >> (4 * q)) & 0xFu));
                } else {
#pragma unroll
                    for (int q = 0; q < 4; q++) stv(dst + q * 32, g4<real>(up[q], lo[q], (bw >> (4 * q)) & 0xFu));
                    dst += 4 * 32;
                }
                bw >>= 16;
            }
            if (tdst) tm_wait_st();
        };

        // ---- two g-outputs of the top stages are VIRTUAL: never stored, recomputed where they are consumed ----------------
        // A g-layer output is one add per node, but at the two top stages it is also the bulk of the scratch traffic: each value is
        // written once and read twice, both times long after the L2 has dropped it (a top g-layer takes a warp ~200 us).  So, for
        // T = LOGN-1 and S = LOGN-2 (stages 9 and 8 of N = 1024), with indices counting V4 groups:
        //   stage T, second half (g from the channel, at leaf N/2):        vT[m]  = ch[m + N/8] + flip(ch[m], B[T] nibble m)
        //   stage S, block 1 (g from the STORED first half of T, at N/4):  vS1[i] = sT[i + N/16] + flip(sT[i], B[S] nibble i)
        // are not written.  Their consumers -- the f-layers below them (at N/2 and N/4), the g-layer that opens stage S-1 at 3N/8
        // and the g-layer at S at 3N/4 -- read the channel / the stored half instead, which all paths of a frame share (one 16-byte
        // request serves the frame's L lanes), plus the path's partial-sum bits.  The pointer field of T and the bit pointers of
        // B[T], B[S] keep doing their job: B[S] of leaf N/4 lives until 3N/4, B[T] until the end.  (Stage S block 3, a g of a g,
        // is stored: recomputing it means eight channel loads per value.)
        // Off (virt == false) when the cooperative prefix runs its in-place butterfly in stage T's array.
        auto nib = [&](const uint32_t *bw, int i) -> uint32_t { return (bw[(i >> 3) * 32] >> ((i & 7) * 4)) & 0xFu; };
        const uint32_t *vbits = nullptr;   // B[T] (mode 0) or B[S] (mode 1) of this path
        const V4 *vsrc = nullptr;          // the channel (mode 0) or the stored first half of stage T (mode 1)
        int vstride = 1, vhalf = 0;
        struct VRaw { V4 up, lo; uint32_t nb; };
        auto v_begin = [&](int mode) {
            constexpr int T = LOGN - 1, S = LOGN - 2;
            if (mode == 0) { vbits = bits_at(T >= 6 ? T : 6) + fbase + bfield(T >= 6 ? T : 6); vsrc = ch4; vstride = 1; vhalf = N >> 3; }
            else { vbits = bits_at(S >= 6 ? S : 6) + fbase + bfield(S >= 6 ? S : 6); vsrc = stage_at(T) + fbase + pfield(T); vstride = 32; vhalf = N >> 4; }
        };
        auto v_load = [&](int i) -> VRaw {
            VRaw r;
            r.up = ldv(vsrc + i * vstride);
            r.lo = ldv(vsrc + (i + vhalf) * vstride);
            r.nb = nib(vbits, i);
            return r;
        };
        // f-layer producing stage s from the virtual stage s+1 (mode 0: s = S from vT; mode 1: s = S-1 from vS1); operands of the next
        // step are in flight while the four CHKs of this one run
        auto f_virtual = [&](int s, int mode) {
            const int cnt4 = 1 << (s - 2);
            v_begin(mode);
            const bool tdst = in_tm(s);
            V4 *dst = tdst ? nullptr : stage_at(s) + lane;
            VRaw ra = v_load(0), rb = v_load(cnt4);
#pragma unroll 1
            for (int i4 = 0; i4 < cnt4; i4++) {
                const V4 x = g4<real>(ra.up, ra.lo, ra.nb), y = g4<real>(rb.up, rb.lo, rb.nb);
                if (i4 + 1 < cnt4) { ra = v_load(i4 + 1); rb = v_load(i4 + 1 + cnt4); }
                const V4 o = f4<real>(x, y);
                if (tdst) tm_st(tm, tmoff(s) + i4, o);
                else stv(dst + i4 * 32, o);
            }
            if (tdst) tm_wait_st();
        };
        // g-layer producing stage t from the virtual stage t+1 (mode 0: t = S at 3N/4 from vT; mode 1: t = S-1 at 3N/8 from vS1)
        auto g_virtual = [&](int t, int mode) {
            const int cnt4 = 1 << (t - 2);
            v_begin(mode);
            const uint32_t *bsrc = bits_at(t) + fbase + bfield(t);   // t >= LOGN-3 >= 6
            const bool tdst = in_tm(t);
            V4 *dst = tdst ? nullptr : stage_at(t) + lane;
#pragma unroll 1
            for (int i4 = 0; i4 < cnt4; i4 += 2) {   // eight loads in flight per lane, as in g_layer
                const VRaw u0 = v_load(i4), l0 = v_load(i4 + cnt4), u1 = v_load(i4 + 1), l1 = v_load(i4 + 1 + cnt4);
                const V4 o0 = g4<real>(g4<real>(u0.up, u0.lo, u0.nb), g4<real>(l0.up, l0.lo, l0.nb), nib(bsrc, i4));
                const V4 o1 = g4<real>(g4<real>(u1.up, u1.lo, u1.nb), g4<real>(l1.up, l1.lo, l1.nb), nib(bsrc, i4 + 1));
                if (tdst) { tm_st(tm, tmoff(t) + i4, o0); tm_st(tm, tmoff(t) + i4 + 1, o1); }
                else { stv(dst + i4 * 32, o0); stv(dst + (i4 + 1) * 32, o1); }
            }
            if (tdst) tm_wait_st();
        };

        // ---- one leaf: frozen -> PM only; information -> decide (SC) or fork/prune (list) ------------------
        // first: the even leaf of a pair (its stage-1 values are read again by the odd leaf); keep2: stage 2 is still needed
        auto leaf = [&](bool info, real lam, bool first, bool keep2) -> uint32_t {
            if (L == 1) return (info && !(lam >= (real)0)) ? 1u : 0u;  // SC_128.c:426-431
            const real ab = rabs(lam);
            const real t = tbl8<real>(ab);
            real pen = t;
            pen += ab;  // PHI: result = table; result += |l| on a mismatch (SCL_1024.c:489-500)
            if (!info) {
                pm = pm + ((lam < (real)0) ? pen : t);
                return 0u;
            }
            const real cm = pm + t, cM = pm + pen;      // the cheaper and the dearer child of this path (cm <= cM)
            const real c0 = (lam < (real)0) ? cM : cm;  // bit 0
            const real c1 = (lam > (real)0) ? cM : cm;  // bit 1
            // The reference keeps the candidates with PM < med, med = (L+1)-th smallest of the 2L (SCL_1024.c:619-633).
            // lo = L-th and hi = (L+1)-th smallest from a bitonic merge network over the 2L values laid out as
            // index = 2*slot + {x, y}: every merge starts with the mirror step (element i against element size-1-i, the
            // partner lane's OTHER register), steps at distance 1 are in-lane, and the last merge stops after its mirror step:
            // the lower L elements are then the L smallest, so lo is their maximum and hi the minimum of the upper L.
            real x = cm, y = cM;
#pragma unroll
            for (int size = 2; size <= L; size <<= 1) {
                const bool low = (k & (size >> 1)) == 0;
                const real ox = __shfl_xor_sync(0xffffffffu, x, size - 1);
                const real oy = __shfl_xor_sync(0xffffffffu, y, size - 1);
                x = low ? rmin(x, oy) : rmax(x, oy);
                y = low ? rmin(y, ox) : rmax(y, ox);
                if (size < L) {
#pragma unroll
                    for (int d = size >> 2; d > 0; d >>= 1) {
                        const bool lowd = (k & d) == 0;
                        const real px = __shfl_xor_sync(0xffffffffu, x, d);
                        const real py = __shfl_xor_sync(0xffffffffu, y, d);
                        x = lowd ? rmin(x, px) : rmax(x, px);
                        y = lowd ? rmin(y, py) : rmax(y, py);
                    }
                    const real mn = rmin(x, y);
                    y = rmax(x, y);
                    x = mn;
                }
            }
            const bool lowh = (k & (L >> 1)) == 0;
            real v = lowh ? rmax(x, y) : rmin(x, y);
#pragma unroll
            for (int d = 1; d < (L >> 1); d <<= 1) {
                const real pv = __shfl_xor_sync(0xffffffffu, v, d);
                v = lowh ? rmax(v, pv) : rmin(v, pv);
            }
            const real w = __shfl_xor_sync(0xffffffffu, v, L >> 1);
            const real lo = lowh ? v : w, hi = lowh ? w : v;
            bool k0, k1;  // candidate survives
            if (__all_sync(0xffffffffu, lo < hi)) {  // no tie across the list boundary in any frame of the warp (lo finite)
                k0 = c0 < hi;
                k1 = c1 < hi;
            } else {
                const bool f0 = c0 < INF, f1 = c1 < INF;
                const real d0 = f0 ? c0 : INF, d1 = f1 ? c1 : INF;
                int r0 = 0, r1 = 0, le0 = 0, le1 = 0;
#pragma unroll 1
                for (int i = 0; i < L; i++) {
                    const real v0 = __shfl_sync(0xffffffffu, d0, i, L);
                    const real v1 = __shfl_sync(0xffffffffu, d1, i, L);
                    // candidate order of the reference's PMcand array: bit-0 branch of path i at i, bit-1 branch at i+L
                    r0 += (int)((v0 < d0) || (v0 == d0 && i < k)) + (int)(v1 < d0);
                    r1 += (int)(v0 <= d1) + (int)((v1 < d1) || (v1 == d1 && i < k));
                    le0 += (int)(v0 <= d0) + (int)(v1 <= d0);
                    le1 += (int)(v0 <= d1) + (int)(v1 <= d1);
                }
                k0 = f0 && r0 < L;
                k1 = f1 && r1 < L;
                const bool q0 = f0 && le0 <= L, q1 = f1 && le1 <= L;  // the reference's "PM < med"
                const uint32_t amb = __ballot_sync(0xffffffffu, (q0 != k0) || (q1 != k1));
                if (amb & (LMASK << fbase)) flags |= kInfoTie;
            }
            const uint32_t m0 = __ballot_sync(0xffffffffu, k0), m1 = __ballot_sync(0xffffffffu, k1);
            const uint32_t fmask = LMASK << fbase;
            const uint32_t both = m0 & m1 & fmask, dead = ~(m0 | m1) & fmask;
            int src = lane;
            real pc1 = c1;
            if (__any_sync(0xffffffffu, !(k0 || k1))) {  // some slot of the warp is free: paths may move (otherwise every path
                                                          // keeps exactly one child in place and nothing is exchanged)
                // t-th free slot (ascending) takes the bit-1 branch of the t-th both-survivor (ascending): SCL_1024.c:636-660.
                // The both-survivors publish their lane at their rank; the free slots read the entry of their own rank.
                const uint32_t below = (1u << lane) - 1u;
                if (k0 && k1) sm_slot[fbase + __popc(both & below)] = (uint32_t)lane;
                __syncwarp();
                if (!(k0 || k1) && __popc(dead & below) < __popc(both)) src = (int)sm_slot[fbase + __popc(dead & below)];
                __syncwarp();
                pc1 = __shfl_sync(0xffffffffu, c1, src);
                if (keep2) {  // stage 2 is read again by the g step before leaf 2 only
#pragma unroll
                    for (int e = 0; e < 4; e++) s2[e] = __shfl_sync(0xffffffffu, s2[e], src);
                }
                if (first) {  // stage 1 is read by the g step of the odd leaf that follows
                    s1[0] = __shfl_sync(0xffffffffu, s1[0], src);
                    s1[1] = __shfl_sync(0xffffffffu, s1[1], src);
                }
                ptr = __shfl_sync(0xffffffffu, ptr, src);
                bptr = __shfl_sync(0xffffffffu, bptr, src);
                Blow = __shfl_sync(0xffffffffu, Blow, src);
                B5 = __shfl_sync(0xffffffffu, B5, src);
                ug = __shfl_sync(0xffffffffu, ug, src);
            }
            uint32_t u;
            if (src != lane) { pm = pc1; u = 1u; }
            else if (k0) { pm = c0; u = 0u; }
            else if (k1) { pm = c1; u = 1u; }
            else { pm = INF; u = 0u; }  // empty slot: stays a copy of path 0 (which keeps its bit-0 branch while the list fills)
            return u;
        };

        // =================================================================== the N/4 leaf groups
        V4 *const st3 = stage_at(3);
        // two V4s of stage s (3 or 4): of this lane's home array / of the array the pointer word names.  A tensor-memory row is
        // reachable by its own lane only: everybody loads its own and the pointed-to slot's values arrive by shuffle.
        auto ld_own2 = [&](int s, int ia, int ib, V4 &A, V4 &B) {
            if (in_tm(s)) { A = tm_ld<real>(tm, tmoff(s) + ia); B = tm_ld<real>(tm, tmoff(s) + ib); tm_wait_ld(); }
            else { const V4 *src = stage_at(s) + lane; A = ldv(src + ia * 32); B = ldv(src + ib * 32); }
        };
        auto ld_ptr2 = [&](int s, int ia, int ib, V4 &A, V4 &B) {
            if (in_tm(s)) {
                A = tm_ld<real>(tm, tmoff(s) + ia); B = tm_ld<real>(tm, tmoff(s) + ib);
                tm_wait_ld();
                if (L > 1) { const int sl = fbase + pfield(s); A = shfl_v4<real>(A, sl); B = shfl_v4<real>(B, sl); }
            } else { const V4 *src = stage_at(s) + fbase + pfield(s); A = ldv(src + ia * 32); B = ldv(src + ib * 32); }
        };

        // =================================================================== the frozen prefix
        // Before the first information bit a frame has ONE path and every decision is 0, so the schedule has no data dependence:
        // the subtree over the first 2^D leaves (D = bits of the first information index) is a plain butterfly with g = lower +
        // upper.  The frame's L lanes split it: f-layers down to stage D, then in-place levels D-1..2 in slot 0's stage-D array,
        // then the two in-register levels and PHI per leaf.  The path metric is the sum of the leaf penalties IN LEAF ORDER
        // (bit-exact with the reference's running sum), taken by every lane from the stored penalties.  On the way down the
        // block that holds the first information bit is copied into slot 0's arrays of every stage, which is the state the
        // loop below resumes from (all pointer fields 0, all partial sums 0).
        int j4_start = 0;
        // virtual top stages (see f_virtual): not when the prefix butterfly would work in stage T's array (first information bit
        // beyond leaf N/4), and only where the three stages involved are scratch stages with their bits in a bit array
        const bool virt = POLAR_VIRT && LOGN >= 9 && (LOGN - 3 >= GLO || HAS_TM) && LOGN - 3 >= 6 &&
                          !(L > 1 && a.coop_groups >= 2 && 4 * a.coop_groups > (N >> 2));
        if (L > 1 && a.coop_groups >= 2) {
            int P = a.coop_groups;
            if (P > N / 8) P = N / 8;               // keep the subtree inside the first half: its root is then an f-layer output
            int D = 32 - __clz(4 * P - 1);          // smallest subtree [0, 2^D) that holds the leaves 0..4P-1; 3 <= D <= LOGN-1
            if (HAS_TM && D < GLO) D = GLO;         // the butterfly works in ONE array shared by the frame's lanes: a global stage (a larger
                                                    // subtree is as good: every block copied out below starts at or before leaf 4P)
            const bool inside = 4 * P < (1 << D);   // it also holds leaf 4P: the loop resumes inside it
            for (int s = LOGN - 1; s >= D; s--) f_layer(s, true);
            V4 *const buf = stage_at(D) + fbase;  // V4 i of the subtree at buf[i * 32]
            const int n4 = 1 << (D - 2);
#pragma unroll 1
            for (int s = D - 1; s >= 2; s--) {    // level s: blocks of 2^(s+1) values -> f half | g half
                const int h4 = 1 << (s - 2);      // V4s per half
#pragma unroll 1
                for (int q = k; q < (n4 >> 1); q += L) {
                    const int i4 = ((q >> (s - 2)) << (s - 1)) | (q & (h4 - 1));
                    const V4 up = ldv(buf + i4 * 32), lo = ldv(buf + (i4 + h4) * 32);
                    stv(buf + i4 * 32, f4<real>(up, lo));
                    stv(buf + (i4 + h4) * 32, g4<real>(up, lo, 0u));
                }
                __syncwarp();
                if (inside && s >= 3) {  // the stage-s block that contains leaf 4P
                    const int b4 = ((4 * P) >> s) << (s - 2);
                    if (in_tm(s)) {  // tensor memory: every lane keeps a copy in its own row (all pointer fields say slot 0)
#pragma unroll 1
                        for (int q = 0; q < h4; q++) tm_st(tm, tmoff(s) + q, ldv(buf + (b4 + q) * 32));
                        tm_wait_st();
                    } else {
                        V4 *dst = stage_at(s) + fbase;
#pragma unroll 1
                        for (int q = k; q < h4; q += L) stv(dst + q * 32, ldv(buf + (b4 + q) * 32));
                        __syncwarp();
                    }
                }
            }
#pragma unroll 1
            for (int q = k; q < n4; q += L) {     // levels 1 and 0 of one 4-block, then the penalties of its four leaves
                const V4 v = ldv(buf + q * 32);
                const real f0 = chk_lean<real>(v.v[0], v.v[2]), f1 = chk_lean<real>(v.v[1], v.v[3]);
                const real g0 = v.v[2] + v.v[0], g1 = v.v[3] + v.v[1];
                real lam[4] = {chk_lean<real>(f0, f1), f1 + f0, chk_lean<real>(g0, g1), g1 + g0};
                V4 pen;
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const real ab = rabs(lam[e]);
                    const real t = tbl8<real>(ab);
                    real tp = t;
                    tp += ab;                      // PHI of a frozen leaf: table, plus |l| when l < 0 (SCL_1024.c:489-500)
                    pen.v[e] = (lam[e] < (real)0) ? tp : t;
                }
                stv(buf + q * 32, pen);
            }
            __syncwarp();
            real acc = (real)0;
#pragma unroll 1
            for (int q = 0; q < P; q++) {
                const V4 pen = ldv(buf + q * 32);
                acc = acc + pen.v[0];
                acc = acc + pen.v[1];
                acc = acc + pen.v[2];
                acc = acc + pen.v[3];
            }
            if (k == 0) pm = acc;
            // partial sums of the prefix are zero: slot 0's bit arrays (every pointer field says slot 0)
#pragma unroll 1
            for (int sb = 6; sb <= LOGN; sb++) {
                uint32_t *z = bits_at(sb) + fbase;
#pragma unroll 1
                for (int w = k; w < (1 << (sb - 5)); w += L) z[w * 32] = 0u;
            }
            __syncwarp();
            j4_start = P;
        }

#pragma unroll 1
        for (int j4 = j4_start; j4 < N / 4; j4++) {
            // ---- descend to stage 2 of this 4-block.  Stage 2 lives in registers only (its four values are consumed by
            // the four leaves below and cloned by shuffle); stage 3 goes through registers to the f step that follows it.
            if (j4 & 1) {  // g at stage 2 from stage 3 (via the pointer word)
                V4 u, l;
                ld_ptr2(3, 0, 1, u, l);
                const V4 v = g4<real>(u, l, Blow & 0xFu);
#pragma unroll
                for (int e = 0; e < 4; e++) s2[e] = v.v[e];
            } else {
                V4 a3, b3;
                int top = 3;
                bool stored3 = false;
                if (j4 & 2) {  // g at stage 3 from stage 4 (via the pointer word)
                    const uint32_t bw = Blow >> 4;
                    V4 u, l;
                    ld_ptr2(4, 0, 2, u, l);
                    a3 = g4<real>(u, l, bw & 0xFu);
                    ld_ptr2(4, 1, 3, u, l);
                    b3 = g4<real>(u, l, (bw >> 4) & 0xFu);
                } else {       // a longer chain: g at the stage the finished block opens, f-layers down to stage 4, f at stage 3
                    int s = LOGN - 1;
                    top = s;
                    if (POLAR_CHUNK > 0 && !POLAR_VIRT && !HAS_TM && FOLD_F3) {
                        // Chunked chain: the g-layer at stage t is produced POLAR_CHUNK rows of each half at a time and the f-layer below
                        // consumes those rows at once, while they are still in the L2 (a whole top g-layer takes a warp ~200 us, the L2
                        // keeps a line ~30 us: the unchunked f-layer re-reads everything from HBM).  One call site per loop body.
                        int t = 0;                                  // stage of the g-layer that opens the chain (0: none, leaf 0)
                        if (j4 != 0) { t = __ffs(j4) - 1 + 2; top = t; s = t - 1; }
                        const int h = (t >= 5) ? (1 << (t - 3)) : 0;   // rows per half of stage t = rows of the f-layer at t-1
                        const int c = (t >= POLAR_CHUNK_MIN && h > POLAR_CHUNK) ? POLAR_CHUNK : 0;   // 0: whole layers
                        int r0 = 0, ph = (t == 0) ? 2 : 0;
                        while (s >= 3) {
                            if (ph < 2) {                            // g rows: c of each half, or the whole layer
                                g_layer(t, c ? r0 + ph * h : 0, c ? c : -1);
                                ph = c ? ph + 1 : 2;
                                continue;
                            }
                            const bool part = c && s == t - 1;
                            f_layer(s, false, part ? r0 : 0, part ? c : -1);
                            if (part && r0 + c < h) { r0 += c; ph = 0; continue; }
                            s--;
                        }
                        ld_own2(3, 0, 1, a3, b3);
                        stored3 = true;
                    } else {
                    if (j4 != 0) {
                        s = __ffs(j4) - 1 + 2;
                        top = s;
                        if (virt && s == LOGN - 1) {          // leaf N/2: f-layer at S straight from the channel
                            f_virtual(LOGN - 2, 0);
                            s = LOGN - 3;
                        } else if (virt && s == LOGN - 2 && j4 == (N >> 4)) {   // leaf N/4: f-layer at S-1 from the virtual block 1 of S
                            f_virtual(LOGN - 3, 1);
                            s = LOGN - 4;
                        } else if (virt && s == LOGN - 2) {   // leaf 3N/4: g-layer at S from the virtual half of T (stored)
                            g_virtual(s, 0);
                            s--;
                        } else if (virt && s == LOGN - 3 && j4 == 3 * (N >> 5)) {  // leaf 3N/8: g-layer at S-1 from the virtual block 1 of S
                            g_virtual(s, 1);
                            s--;
                        } else {
                            g_layer(s);
                            s--;
                        }
                    }
                    if (FOLD_F3) {
#pragma unroll 1
                        for (; s >= 3; s--) f_layer(s, false);   // stage 3 too: one copy of the four-CHK block less (code size: see f4)
                        ld_own2(3, 0, 1, a3, b3);
                        stored3 = true;
                    } else {
#pragma unroll 1
                        for (; s >= 4; s--) f_layer(s, false);
                        V4 u, l;                 // own home
                        ld_own2(4, 0, 2, u, l);
                        a3 = f4<real>(u, l);
                        ld_own2(4, 1, 3, u, l);
                        b3 = f4<real>(u, l);
                    }
                    }
                }
                if (stored3) {
                } else if (in_tm(3)) {
                    tm_st(tm, tmoff(3), a3);
                    tm_st(tm, tmoff(3) + 1, b3);
                    tm_wait_st();
                } else {
                    stv(st3 + lane, a3);
                    stv(st3 + 32 + lane, b3);
                }
                const V4 v = f4<real>(a3, b3);
#pragma unroll
                for (int e = 0; e < 4; e++) s2[e] = v.v[e];
                set_pfields(top);
                __syncwarp();
            }
            if (POLAR_PF_MAX && (j4 & 3) == 3 && j4 + 1 < N / 4) {
                // the next leaf group opens with a g-layer whose source (stage t+1, whichever slot the pointer word will name after
                // the four leaves below) is one of the frame's home arrays: every lane prefetches its own into the L1
                const int t = __ffs(j4 + 1) + 1;
                if (t <= POLAR_PF_MAX && t + 1 < LOGN && t + 1 >= GLO) {
                    const V4 *pf = stage_at(t + 1) + lane;
#pragma unroll 1
                    for (int i4 = 0; i4 < (1 << (t - 1)); i4++) prefetch_l1(pf + i4 * 32);
                }
            }
            ug = 0;
            const uint32_t inib = a.m.info[j4 >> 3] >> ((j4 & 7) * 4);  // which of the four leaves carry information
            if (POLAR_FRZ4 && L > 1 && (inib & 0xFu) == 0u) {
                // four frozen leaves (every decision 0): the two in-register levels are a fixed butterfly, so the four leaf LLRs and
                // their table look-ups are independent; only the four path-metric additions keep the leaf order (SCL_1024.c:489-500)
                const real f0 = chk_lean<real>(s2[0], s2[2]), f1 = chk_lean<real>(s2[1], s2[3]);
                const real g0 = s2[2] + s2[0], g1 = s2[3] + s2[1];
                const real lam[4] = {chk_lean<real>(f0, f1), f1 + f0, chk_lean<real>(g0, g1), g1 + g0};
                real pen[4];
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const real ab = rabs(lam[e]);
                    const real t = tbl8<real>(ab);
                    real tp = t;
                    tp += ab;
                    pen[e] = (lam[e] < (real)0) ? tp : t;
                }
#pragma unroll
                for (int e = 0; e < 4; e++) pm = pm + pen[e];
            } else
#pragma unroll 1
            for (int p = 0; p < 2; p++) {
                if (p == 0) {  // f at stage 1
                    s1[0] = chk_lean<real>(s2[0], s2[2]);
                    s1[1] = chk_lean<real>(s2[1], s2[3]);
                } else {       // g at stage 1, partial sums (u0^u1, u1)
                    s1[0] = s2[2] + RT::flip(s2[0], (ug ^ (ug >> 1)) & 1u);
                    s1[1] = s2[3] + RT::flip(s2[1], (ug >> 1) & 1u);
                }
                if (FOLD_LEAF) {
                    // the two leaves of the pair go through ONE copy of the leaf body (code size: see f4)
                    uint32_t ua = 0;
#pragma unroll 1
                    for (int q = 0; q < 2; q++) {
                        real lam;
                        if (q == 0) lam = chk_lean<real>(s1[0], s1[1]);   // f at stage 0
                        else lam = s1[1] + RT::flip(s1[0], ua);           // g at stage 0
                        ua = leaf((inib >> (2 * p + q)) & 1u, lam, q == 0, p == 0);
                        ug |= ua << (2 * p + q);  // ug travels with the path when the next leaf clones it
                    }
                } else {
                    const uint32_t ua = leaf((inib >> (2 * p)) & 1u, chk_lean<real>(s1[0], s1[1]), true, p == 0);   // f at stage 0
                    ug |= ua << (2 * p);  // ug travels with the path when the next leaf clones it
                    const uint32_t ub = leaf((inib >> (2 * p + 1)) & 1u, s1[1] + RT::flip(s1[0], ua), false, p == 0);  // g at stage 0
                    ug |= ub << (2 * p + 1);
                }
            }

            // ---- partial sums of the finished 4-block, pushed up while the block closes larger blocks -------
            const uint32_t u0 = ug & 1u, u1 = (ug >> 1) & 1u, u2b = (ug >> 2) & 1u, u3 = (ug >> 3) & 1u;
            uint32_t t32 = (u0 ^ u1 ^ u2b ^ u3) | ((u1 ^ u3) << 1) | ((u2b ^ u3) << 2) | (u3 << 3);
            const int z = __ffs(~j4) - 1;  // trailing ones of j4
            int T = 2 + z;
            if (T > LOGN) T = LOGN;
            if (T > 2) t32 = ((Blow & 0xF) ^ t32) | (t32 << 4);
            if (T > 3) t32 = (((Blow >> 4) & 0xFF) ^ t32) | (t32 << 8);
            if (T > 4) t32 = (((Blow >> 12) & 0xFFFF) ^ t32) | (t32 << 16);
            if (T == 2) Blow = (Blow & ~0xFu) | t32;
            else if (T == 3) Blow = (Blow & ~0xFF0u) | (t32 << 4);
            else if (T == 4) Blow = (Blow & ~0xFFFF000u) | (t32 << 12);
            else if (T == 5) B5 = t32;
            else {
                uint32_t *dst = bits_at(T) + lane;
                dst[0] = B5 ^ t32;
                dst[32] = t32;
#pragma unroll 1
                for (int sb = 6; sb < T; sb++) {
                    const int len = 1 << (sb - 5);
                    const uint32_t *srcb = bits_at(sb) + fbase + bfield(sb);
#pragma unroll 1
                    for (int i = 0; i < len; i++) {
                        const uint32_t d = dst[i * 32];
                        dst[(i + len) * 32] = d;
                        dst[i * 32] = d ^ srcb[i * 32];
                    }
                }
                set_bfield(T);
                __syncwarp();
            }
        }

        // =================================================================== termination
        // x_hat = B[LOGN] of this path; u_hat = x_hat F^{(x)n}
        uint32_t xw[W];
        if (LOGN == 5) {
            xw[0] = B5;
        } else {
            const uint32_t *srcb = bits_at(LOGN) + fbase + bfield(LOGN);
#pragma unroll
            for (int w = 0; w < W; w++) xw[w] = srcb[w * 32];
        }
#pragma unroll
        for (int w = 0; w < W; w++) xw[w] = polar_word_stages(xw[w]);
#pragma unroll
        for (int d = 1; d < W; d <<= 1)
#pragma unroll
            for (int w = 0; w < W; w++)
                if (!(w & d)) xw[w] ^= xw[w + d];

        int best = 0;
        if (L > 1) {
            bool pass = false;
            if (a.use_crc) {  // CRcheck: remainder of the I[]-ordered word modulo g(D), as r parity masks
                uint32_t syn = 0;
#pragma unroll 1
                for (int b = 0; b < a.r; b++) {
                    uint32_t acc = 0;
#pragma unroll
                    for (int w = 0; w < W; w++) acc ^= xw[w] & __ldg(a.crc_masks + b * W + w);
                    syn |= (uint32_t)(__popc(acc) & 1);
                }
                pass = (syn == 0) && (pm < INF);
            }
            // CASCL_1024_L8.c:725-755 / SCL_1024.c:667-674: first index wins ties.  Both scans run on every lane
            // (frames of one warp may disagree on whether a CRC-passing path exists; shuffles must stay convergent).
            int bi = -1, mi = 0;
            real bpm = INF, mpm = __shfl_sync(0xffffffffu, pm, 0, L);
#pragma unroll
            for (int i = 0; i < L; i++) {
                const real v = __shfl_sync(0xffffffffu, pm, i, L);
                const bool ps = __shfl_sync(0xffffffffu, (int)pass, i, L) != 0;
                if (ps && (bi < 0 || v < bpm)) { bi = i; bpm = v; }
                if (i > 0 && v < mpm) { mi = i; mpm = v; }
            }
            if (bi < 0) {
                if (a.use_crc) flags |= kInfoCrcFail;
                bi = mi;
            }
            best = bi;
        }
        uint32_t fr_flags = flags;
        if (L > 1) {  // OR over the frame's lanes
#pragma unroll
            for (int d = 1; d < L; d <<= 1) fr_flags |= __shfl_xor_sync(0xffffffffu, fr_flags, d);
        }
        if (valid && k == best) {
            uint32_t nerr = 0;
            if (a.truth) {
                const uint32_t *tw = a.truth + frame * (size_t)W;
#pragma unroll
                for (int w = 0; w < W; w++) nerr += __popc((xw[w] ^ __ldg(tw + w)) & a.m.cnt[w]);
            }
            if (a.u_hat) {
                uint32_t *ow = a.u_hat + frame * (size_t)W;
#pragma unroll
                for (int w = 0; w < W; w++) ow[w] = xw[w] & a.m.info[w];
            }
            if (a.frame_info) a.frame_info[frame] = (nerr > 0xFFFFu ? 0xFFFFu : nerr) | fr_flags;
            if (a.counters) {
                atomicAdd(a.counters + CNT_FRAMES, 1ull);
                if (nerr) { atomicAdd(a.counters + CNT_ERR_BLOCKS, 1ull); atomicAdd(a.counters + CNT_ERR_BITS, (unsigned long long)nerr); }
                if (fr_flags & kInfoTie) atomicAdd(a.counters + CNT_TIE, 1ull);
                if (fr_flags & kInfoCrcFail) atomicAdd(a.counters + CNT_CRCFAIL, 1ull);
            }
        }
        __syncwarp();
    }
    if (HAS_TM) {  // the allocation goes back once every warp of the CTA has left its loop
        __syncthreads();
        if (wi == 0) tm_free(tm_base, C::TM_COLS);
    }
}

// ---------------------------------------------------------------- dispatch
#ifndef POLAR_SMEM_TOP
#define POLAR_SMEM_TOP 5
#endif
#ifndef POLAR_BITS_TOP
#define POLAR_BITS_TOP 7
#endif
// Tensor-memory layouts are compiled in only on request (-DPOLAR_TML=a -DPOLAR_TMH=b: fp32 LLR stages a..b-1 in TMEM).  Measured
// on B200 (CA-SCL 1024 L=8, fp32; profiles/r2_tmem_experiments.md): stage 6 in TMEM + stages 3..5 in shared memory (7 CTAs of
// four warps) cuts the HBM traffic by 29 % but decodes 11.2 M frames/s against 13.4 M for the layout below -- the larger code
// (stall_no_instruction 22 % of the samples), 28 instead of 32 warps and the lost L1 outweigh the shorter memory stalls; stages
// 3..5 in TMEM and no shared-memory stage (32 warps, 200 KB of L1): 10.2 M frames/s, 12 % more instructions (own-row loads + shuffles
// for every pointer read, tcgen05.wait).  The code stays because the CPU emulator tests pin it and the option costs nothing.
#ifndef POLAR_TML
#define POLAR_TML 3
#endif
#ifndef POLAR_TMH
#define POLAR_TMH 0   // TMH <= TML: no tensor memory (the product configuration)
#endif
// which kernel configuration serves (real, N, L); shared by the launcher below and by the CPU emulator tests
template <typename real, int LOGN, int L, bool ALLOW_TM = true>
struct ListDispatchCfg {
    static constexpr bool TM = ALLOW_TM && POLAR_TMH > POLAR_TML && sizeof(real) == 4 && LOGN >= POLAR_TMH + 1;
    static constexpr int TML = TM ? POLAR_TML : 0, TMH = TM ? POLAR_TMH : 0;
    static constexpr int SMEM_TOP = TM ? POLAR_TML : POLAR_SMEM_TOP, BITS_TOP = POLAR_BITS_TOP;
    using C = ListCfg<real, LOGN, L, SMEM_TOP, BITS_TOP, TML, TMH>;
    static constexpr int THREADS = C::THREADS;
};

#ifndef POLAR_EMU
template <typename real, int LOGN, int L>
struct ListDispatch {
    using D = ListDispatchCfg<real, LOGN, L>;
    static constexpr int SMEM_TOP = D::SMEM_TOP, BITS_TOP = D::BITS_TOP, TML = D::TML, TMH = D::TMH;
    using C = typename D::C;
    static cudaError_t plan(ListPlan *p)
    {
        auto kern = list_decode_kernel<real, LOGN, L, SMEM_TOP, BITS_TOP, TML, TMH>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_CTA);
        if (e != cudaSuccess) return e;
        int nb = 0;
        if (C::TM_COLS) {
            // The occupancy API answers 1 for every kernel that contains tcgen05.alloc, although CTAs do share an SM as long as
            // their allocations fit its 512 columns (tools/ubench/tmem_occ.cu: seven 64-column CTAs run side by side): count by hand.
            cudaFuncAttributes fa;
            if ((e = cudaFuncGetAttributes(&fa, kern)) != cudaSuccess) return e;
            int dev = 0;
            cudaDeviceProp pr;
            if ((e = cudaGetDevice(&dev)) != cudaSuccess || (e = cudaGetDeviceProperties(&pr, dev)) != cudaSuccess) return e;
            const int regs_per_warp = ((fa.numRegs * 32 + 255) / 256) * 256;
            const int by_regs = pr.regsPerMultiprocessor / (regs_per_warp * C::WARPS);
            const int by_smem = (int)(pr.sharedMemPerMultiprocessor / (C::SMEM_CTA + fa.sharedSizeBytes + pr.reservedSharedMemPerBlock));
            const int by_threads = pr.maxThreadsPerMultiProcessor / C::THREADS;
            nb = std::min(std::min(by_regs, by_smem), std::min(by_threads, 512 / (C::TM_COLS ? C::TM_COLS : 512)));
        } else {
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, C::THREADS, C::SMEM_CTA);
            if (e != cudaSuccess) return e;
        }
        p->scratch_per_cta = C::GS_BYTES * C::WARPS;
        p->smem = C::SMEM_CTA;
        p->ctas_per_sm = nb;
        p->frames_per_cta = C::FPW * C::WARPS;
        return cudaSuccess;
    }
    static cudaError_t launch(const ListArgs &a, int grid, cudaStream_t st)
    {
        list_decode_kernel<real, LOGN, L, SMEM_TOP, BITS_TOP, TML, TMH><<<grid, C::THREADS, C::SMEM_CTA, st>>>(a);
        return cudaGetLastError();
    }
};

#ifdef POLAR_DEV_CASES  // development builds: only the bench / parity configurations (compile time)
#define POLAR_LIST_CASES(X) X(7, 8) X(10, 1) X(10, 8)
#else
#define POLAR_LIST_CASES(X) \
    X(5, 1) X(5, 2) X(5, 4) X(5, 8) \
    X(6, 1) X(6, 2) X(6, 4) X(6, 8) \
    X(7, 1) X(7, 2) X(7, 4) X(7, 8) X(7, 16) X(7, 32) \
    X(8, 1) X(8, 2) X(8, 4) X(8, 8) X(8, 16) X(8, 32) \
    X(9, 1) X(9, 2) X(9, 4) X(9, 8) X(9, 16) X(9, 32) \
    X(10, 1) X(10, 2) X(10, 4) X(10, 8) X(10, 16) X(10, 32)
#endif

cudaError_t list_plan(int n, int L, bool f64, ListPlan *plan)
{
#define X(NN, LL) \
    if (n == NN && L == LL) return f64 ? ListDispatch<double, NN, LL>::plan(plan) : ListDispatch<float, NN, LL>::plan(plan);
    POLAR_LIST_CASES(X)
#undef X
    return cudaErrorInvalidValue;
}

cudaError_t launch_list(const ListArgs &a, int n, int L, bool f64, int grid, cudaStream_t st)
{
#define X(NN, LL) \
    if (n == NN && L == LL) return f64 ? ListDispatch<double, NN, LL>::launch(a, grid, st) : ListDispatch<float, NN, LL>::launch(a, grid, st);
    POLAR_LIST_CASES(X)
#undef X
    return cudaErrorInvalidValue;
}
#endif  // !POLAR_EMU

}  // namespace polar
