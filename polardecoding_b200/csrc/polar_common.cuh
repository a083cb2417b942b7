// Device-side primitives shared by every kernel of the polar decoding engine (sm_100a).
//
// Arithmetic contract (parity): every floating-point operation below is an IEEE add, subtract,
// compare, min or sign manipulation performed in the SAME order as the reference C code, so the
// `double` instantiation reproduces the reference's fp64 results bit for bit; the `float`
// instantiation is the throughput mode (same formulas, fp32 rounding).
//   chk<real>  : CHK(),  /root/reference/SC_128.c:284-315
//   phi_tbl    : table part of PHI(), /root/reference/SCL_1024.c:481-502 (same 8-level table)
#pragma once
#ifdef POLAR_EMU  // CPU warp emulator, test infrastructure only (tests/emu)
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>

namespace polar {

template <typename real> struct real_traits;
template <> struct real_traits<float> {
    static __device__ __forceinline__ float inf() { return __int_as_float(0x7f800000); }
    static __device__ __forceinline__ float flip(float v, uint32_t bit) {  // bit ? -v : v
        return __int_as_float(__float_as_int(v) ^ (int)(bit << 31));
    }
    static __device__ __forceinline__ float xsign(float m, float a, float b) {  // m with sign(a)*sign(b) applied
        return __int_as_float(__float_as_int(m) ^ ((__float_as_int(a) ^ __float_as_int(b)) & 0x80000000));
    }
    static __device__ __forceinline__ bool same_bits(float a, float b) { return __float_as_int(a) == __float_as_int(b); }
};
template <> struct real_traits<double> {
    static __device__ __forceinline__ double inf() { return __longlong_as_double(0x7ff0000000000000ll); }
    static __device__ __forceinline__ double flip(double v, uint32_t bit) {
        return __hiloint2double(__double2hiint(v) ^ (int)(bit << 31), __double2loint(v));
    }
    static __device__ __forceinline__ double xsign(double m, double a, double b) {
        return __hiloint2double(__double2hiint(m) ^ ((__double2hiint(a) ^ __double2hiint(b)) & 0x80000000), __double2loint(m));
    }
    static __device__ __forceinline__ bool same_bits(double a, double b) {
        return __double_as_longlong(a) == __double_as_longlong(b);
    }
};

__device__ __forceinline__ float rabs(float v) { return fabsf(v); }
__device__ __forceinline__ double rabs(double v) { return fabs(v); }
__device__ __forceinline__ float rmin(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ double rmin(double a, double b) { return fmin(a, b); }

// 8-level table for ln(1+e^-x), x >= 0 (SC_128.c:293-300).  Generic form: a depth-3 select tree
// (7 compares + 7 selects; used for double).
template <typename real>
__device__ __forceinline__ real tbl8(real a)
{
    const real lo = (a < (real)0.433) ? ((a < (real)0.196) ? (real)0.65 : (real)0.55) : ((a < (real)0.71) ? (real)0.45 : (real)0.35);
    const real hi = (a < (real)2.252) ? ((a < (real)1.508) ? (real)0.25 : (real)0.15) : ((a < (real)4.5) ? (real)0.05 : (real)0);
    return (a < (real)1.05) ? lo : hi;
}

// fp32 form on the FMA pipe (measured 1.8x the compare/select form on B200, tools/ubench/chk_variants.cu):
//   [a < t] = sat((t - a) * 2^60)        exact: t*2^60 is representable, the FMA rounds once, |t-a| >= 1 ulp >> 2^-60
//   T(a)    = sum of increments d_k over the thresholds above a, accumulated from the largest threshold down.
// The increments are the fp32 differences of consecutive table literals; each partial sum fl(c + d_k) lands exactly on
// the next literal (0.05f, 0.15f, ... 0.65f; checked in tests/test_cabi.py), so the result is bit-identical to the
// compare/select form -- the same values the reference's table holds, rounded to fp32.
template <>
__device__ __forceinline__ float tbl8<float>(float a)
{
    const float NB = -1.152921504606846976e18f;  // -2^60
    float t = __saturatef(fmaf(a, NB, 4.5f * 1.152921504606846976e18f)) * __int_as_float(0x3d4ccccd);
    t = fmaf(__saturatef(fmaf(a, NB, 2.252f * 1.152921504606846976e18f)), __int_as_float(0x3dccccce), t);
    t = fmaf(__saturatef(fmaf(a, NB, 1.508f * 1.152921504606846976e18f)), __int_as_float(0x3dcccccc), t);
    t = fmaf(__saturatef(fmaf(a, NB, 1.05f * 1.152921504606846976e18f)), __int_as_float(0x3dcccccc), t);
    t = fmaf(__saturatef(fmaf(a, NB, 0.71f * 1.152921504606846976e18f)), __int_as_float(0x3dcccccc), t);
    t = fmaf(__saturatef(fmaf(a, NB, 0.433f * 1.152921504606846976e18f)), __int_as_float(0x3dccccd0), t);
    t = fmaf(__saturatef(fmaf(a, NB, 0.196f * 1.152921504606846976e18f)), __int_as_float(0x3dccccc8), t);
    return t;
}

// reference-order select chain in fp32, kept to check the two forms against each other (tools/ubench, tests)
__device__ __forceinline__ float tbl8_select_f32(float a)
{
    float t = 0.f;
    t = (a < 4.5f) ? 0.05f : t;
    t = (a < 2.252f) ? 0.15f : t;
    t = (a < 1.508f) ? 0.25f : t;
    t = (a < 1.05f) ? 0.35f : t;
    t = (a < 0.71f) ? 0.45f : t;
    t = (a < 0.433f) ? 0.55f : t;
    t = (a < 0.196f) ? 0.65f : t;
    return t;
}

// CHK(a,b) = sign(a)sign(b) min(|a|,|b|) + (T(|a+b|) - T(|a-b|)).
// The reference multiplies the int sign product into the double magnitude (exact) and uses sign(0)=+1;
// a sign-bit XOR differs only for a -0.0 operand, where the magnitude is 0 and (+-0) + delta is the
// same value for every delta the table difference can produce (delta is never -0.0).
template <typename real>
__device__ __forceinline__ real chk(real a, real b)
{
    const real sa = rabs(a + b);
    const real da = rabs(a - b);
    const real delta = tbl8<real>(sa) - tbl8<real>(da);
    const real m = rmin(rabs(a), rabs(b));
    return real_traits<real>::xsign(m, a, b) + delta;
}

// fp32 CHK with the two table sums accumulated together by packed fp32x2 FMAs (Blackwell FFMA2): 27 instead of 34 issue
// slots, bit-identical results.  FFMA2 costs ~2.5 FFMA pipe slots (tools/ubench/chk_variants.cu, V7), so this only pays
// where the issue rate, not the FMA pipe, is the limit: the list decoder gains 5 %, BP (FMA-pipe heavy) loses 2 % -- so
// list_decode.cu uses chk_lean, bp_decode.cu uses chk_mix_f32 with two packed steps (below).
#ifndef POLAR_EMU
__device__ __forceinline__ unsigned long long pk2(float lo, float hi)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpk2(unsigned long long v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b)
{
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ float step_alu(float x, float t)
{
    float r;
    asm("set.lt.f32.f32 %0, %1, %2;" : "=f"(r) : "f"(x), "f"(t));
    return r;
}
#else  // the same operations, element by element, for the CPU warp emulator (IEEE single, one rounding per fma)
__device__ __forceinline__ unsigned long long pk2(float lo, float hi) { return (unsigned long long)__float_as_uint(lo) | ((unsigned long long)__float_as_uint(hi) << 32); }
__device__ __forceinline__ void unpk2(unsigned long long v, float &lo, float &hi) { lo = __uint_as_float((unsigned)v); hi = __uint_as_float((unsigned)(v >> 32)); }
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b)
{
    float al, ah, bl, bh;
    unpk2(a, al, ah); unpk2(b, bl, bh);
    volatile float l = al * bl, h = ah * bh;
    return pk2(l, h);
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c)
{
    float al, ah, bl, bh, cl, ch;
    unpk2(a, al, ah); unpk2(b, bl, bh); unpk2(c, cl, ch);
    return pk2(fmaf(al, bl, cl), fmaf(ah, bh, ch));
}
__device__ __forceinline__ float step_alu(float x, float t) { return (x < t) ? 1.0f : 0.0f; }
#endif
template <typename real>
__device__ __forceinline__ real chk_lean(real a, real b) { return chk<real>(a, b); }
template <>
__device__ __forceinline__ float chk_lean<float>(float a, float b)
{
    const float NB = -1.152921504606846976e18f, B = 1.152921504606846976e18f;
    const float s = fabsf(a + b), d = fabsf(a - b);
    unsigned long long acc, t;
#define POLAR_ST(x, T) __saturatef(fmaf(x, NB, T * B))
#define POLAR_STEP(T, H) acc = fma2(pk2(POLAR_ST(s, T), POLAR_ST(d, T)), pk2(__int_as_float(H), __int_as_float(H)), acc);
    t = pk2(POLAR_ST(s, 4.5f), POLAR_ST(d, 4.5f));
    acc = mul2(t, pk2(__int_as_float(0x3d4ccccd), __int_as_float(0x3d4ccccd)));
    POLAR_STEP(2.252f, 0x3dccccce) POLAR_STEP(1.508f, 0x3dcccccc) POLAR_STEP(1.05f, 0x3dcccccc)
    POLAR_STEP(0.71f, 0x3dcccccc) POLAR_STEP(0.433f, 0x3dccccd0) POLAR_STEP(0.196f, 0x3dccccc8)
#undef POLAR_STEP
#undef POLAR_ST
    float ts, td;
    unpk2(acc, ts, td);
    const float m = fminf(fabsf(a), fabsf(b));
    return real_traits<float>::xsign(m, a, b) + (ts - td);
}

// Balanced form for kernels that sit near several limits at once (BP: issue slots, FMA pipe, ALU pipe).  Per table step
// the indicator [x < t] comes either from the FMA pipe (FFMA.SAT, as above) or from the ALU pipe (FSET.BF, the first KM steps),
// and the two sums accumulate packed (one FFMA2 for both, the first KP steps) or scalar.  Every choice gives exactly 1.0f/0.0f
// indicators and the same accumulation order, so the results are bit-identical for all (KP, KM).
template <int KP, int KM>
__device__ __forceinline__ float chk_mix_f32(float a, float b)
{
    const float NB = -1.152921504606846976e18f, B = 1.152921504606846976e18f;
    const float s = fabsf(a + b), d = fabsf(a - b);
    const float T[7] = {4.5f, 2.252f, 1.508f, 1.05f, 0.71f, 0.433f, 0.196f};
    const int H[7] = {0x3d4ccccd, 0x3dccccce, 0x3dcccccc, 0x3dcccccc, 0x3dcccccc, 0x3dccccd0, 0x3dccccc8};
    float hs[7], hd[7];
#pragma unroll
    for (int k = 0; k < 7; k++) {
        hs[k] = (k < KM) ? step_alu(s, T[k]) : __saturatef(fmaf(s, NB, T[k] * B));
        hd[k] = (k < KM) ? step_alu(d, T[k]) : __saturatef(fmaf(d, NB, T[k] * B));
    }
    float ts, td;
    if (KP == 0) {
        ts = hs[0] * __int_as_float(H[0]);
        td = hd[0] * __int_as_float(H[0]);
    } else {
        unsigned long long acc = mul2(pk2(hs[0], hd[0]), pk2(__int_as_float(H[0]), __int_as_float(H[0])));
#pragma unroll
        for (int k = 1; k < 7; k++)
            if (k < KP) acc = fma2(pk2(hs[k], hd[k]), pk2(__int_as_float(H[k]), __int_as_float(H[k])), acc);
        unpk2(acc, ts, td);
    }
#pragma unroll
    for (int k = 1; k < 7; k++)
        if (k >= KP) {
            ts = fmaf(hs[k], __int_as_float(H[k]), ts);
            td = fmaf(hd[k], __int_as_float(H[k]), td);
        }
    const float m = fminf(fabsf(a), fabsf(b));
    return real_traits<float>::xsign(m, a, b) + (ts - td);
}

// ---------------------------------------------------------------- Philox4x32-10 (counter based)
struct philox4 { uint32_t x, y, z, w; };

__device__ __forceinline__ philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1)
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    philox4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

// ---------------------------------------------------------------- packed-bit polar transform
// In-register butterfly stages 0..4 of x = u F^{(x)n} on one 32-bit word (bit p of the word = position p).
__device__ __forceinline__ uint32_t polar_word_stages(uint32_t w)
{
    w ^= (w >> 1) & 0x55555555u;
    w ^= (w >> 2) & 0x33333333u;
    w ^= (w >> 4) & 0x0F0F0F0Fu;
    w ^= (w >> 8) & 0x00FF00FFu;
    w ^= (w >> 16) & 0x0000FFFFu;
    return w;
}

}  // namespace polar
