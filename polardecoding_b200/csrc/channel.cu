// Fused frame source: payload -> CRC -> placement on the information set -> polar encoding -> BPSK ->
// AWGN -> channel LLR, one kernel, counter-based random numbers (sm_100a).
//
// What it replaces (reference): the per-frame block of main(), /root/reference/SC_128.c:171-202 and
// /root/reference/CASCL_1024_L8.c:245-292 (PN-63 payload :126-138,180; CRC by multiplication with g(D)
// :251-266 or systematic, CASCL_1024_sys.c:776-789; O(N^2) row-XOR encoder :183-191; y = +-1 + noise;
// LLR = 2y/std/std, SC_128.c:418) and the Ranq1 / Marsaglia-polar generator (SC_128.c:236-267).
// The reference's generator is sequential (rejection sampling), so it cannot be partitioned; here every
// group of four code bits of every frame draws from Philox4x32-10 at counter (frame, position/4) under the
// key `seed`, which makes a frame's noise independent of batch size, launch geometry and rank.
//
// Layout: a warp owns 1024 consecutive code bits = 1024/N frames; lane l holds packed word l of that span.
// Encoding is the n-stage butterfly on packed words (5 in-register stages + shuffles).  LLRs are written as
// 16-byte vectors, 512 contiguous bytes per warp instruction; the truth vector u is written packed.
#include "engine.h"
#include <cuda_fp16.h>
#include <algorithm>
#include "polar_common.cuh"

namespace polar {

template <typename real>
__global__ void __launch_bounds__(128) channel_kernel(const ChannelArgs a, const unsigned long long pn63)
{
    extern __shared__ uint32_t sm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int N = a.N, W = N >> 5, FPW = 32 / W;  // W words per frame, FPW frames per warp
    // per warp: 32 words payload v, 32 words w (payload+CRC), 32 words u
    uint32_t *vw = sm + wib * 96, *ww = vw + 32, *uw = vw + 64;
    const unsigned long long spans = (a.B + FPW - 1) / FPW;
    const int fw = lane / W, wl = lane - fw * W;  // my frame within the span, my word within the frame
    const int KW = (a.K + 31) >> 5;               // payload words per frame (<= W)

    for (unsigned long long sp = (unsigned long long)blockIdx.x * wpb + wib; sp < spans; sp += (unsigned long long)gridDim.x * wpb) {
        const unsigned long long f_local = sp * FPW + fw;           // frame index within this call
        const unsigned long long F = a.first_frame + f_local;       // global frame index
        const bool valid = f_local < a.B;

        // ---- payload word wl of my frame
        uint32_t v = 0;
        if (wl < KW) {
            if (a.data_mode == 0) {
                const uint32_t m = (uint32_t)((F % 63ull) * (unsigned long long)(a.K % 63) % 63ull);  // phase after F frames (SC_128.c:214-215)
                uint32_t idx = (m + 32u * (uint32_t)wl) % 63u;
#pragma unroll 4
                for (int b = 0; b < 32; b++) {
                    v |= (uint32_t)((pn63 >> idx) & 1ull) << b;
                    idx = (idx == 62u) ? 0u : idx + 1u;
                }
            } else {
                const philox4 p = philox4x32_10((uint32_t)F, (uint32_t)(F >> 32), (uint32_t)(wl >> 2), 0xDA7Au, (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
                v = (wl & 3) == 0 ? p.x : (wl & 3) == 1 ? p.y : (wl & 3) == 2 ? p.z : p.w;
            }
            const int rem = a.K - 32 * wl;
            if (rem < 32) v &= (1u << rem) - 1u;
        }
        vw[lane] = v;
        uw[lane] = 0;
        __syncwarp();
        const uint32_t *vf = vw + fw * W;  // my frame's payload words
        auto vbit = [&](int i) -> uint32_t { return (i >= 0 && i < a.K) ? ((vf[i >> 5] >> (i & 31)) & 1u) : 0u; };

        // ---- w = payload (+CRC), word wl of my frame
        uint32_t wv = 0;
        if (a.r == 0) {
            wv = v;
        } else if (!a.crc_systematic) {
            // w(D) = v(D) g(D): w = XOR over the taps e of (v << e), on packed words  (CASCL_1024_L8.c:251-266)
            for (int e = 0; e <= a.r; e++) {
                if (!((a.crc_poly >> e) & 1ull)) continue;
                const int ws = e >> 5, bs = e & 31;          // word and bit part of the shift
                const int src = wl - ws;
                uint32_t lo = (src >= 0 && src < W) ? vf[src] : 0u;
                uint32_t hi = (src - 1 >= 0 && src - 1 < W) ? vf[src - 1] : 0u;
                wv ^= bs ? ((lo << bs) | (hi >> (32 - bs))) : lo;
            }
            const int rem = a.nI - 32 * wl;
            if (rem < 32) wv &= (rem > 0) ? ((1u << rem) - 1u) : 0u;
        } else {
            // parity p(D) = v(D) D^r mod g(D) in w[0..r), payload in w[r..r+K)  (CASCL_1024_sys.c:781-789)
            uint32_t par = 0;
            for (int i = wl; i < a.K; i += W)
                if (vbit(i)) par ^= __ldg(a.crc_sys + i);
            for (int d = 1; d < W; d <<= 1) par ^= __shfl_xor_sync(0xffffffffu, par, d);
            {   // payload shifted up by r bits, parity in bits 0..r-1 (r <= 32)
                const int ws = a.r >> 5, bs = a.r & 31, src = wl - ws;
                const uint32_t lo = (src >= 0 && src < W) ? vf[src] : 0u;
                const uint32_t hi = (src - 1 >= 0 && src - 1 < W) ? vf[src - 1] : 0u;
                wv = bs ? ((lo << bs) | (hi >> (32 - bs))) : lo;
                if (wl == 0) wv |= (a.r == 32) ? par : (par & ((1u << a.r) - 1u));
            }
        }
        ww[lane] = wv;
        __syncwarp();
        // ---- u[I[i]] = w[i]  (SC_128.c:179-181): scatter in reliability order
        {
            const uint32_t *wf = ww + fw * W;
            uint32_t *uf = uw + fw * W;
            for (int i = wl; i < a.nI; i += W)
                if ((wf[i >> 5] >> (i & 31)) & 1u) {
                    const int p = __ldg(a.I + i);
                    atomicOr(&uf[p >> 5], 1u << (p & 31));
                }
        }
        __syncwarp();
        const uint32_t u = uw[lane];
        // ---- x = u F^{(x)n}
        uint32_t x = polar_word_stages(u);
        for (int d = 1; d < W; d <<= 1) {
            const uint32_t o = __shfl_down_sync(0xffffffffu, x, d);
            if (!(wl & d)) x ^= o;
        }
        if (valid && a.u_packed) a.u_packed[f_local * (size_t)W + wl] = u;

        // ---- BPSK + AWGN + LLR for the span's 1024 positions, 4 per lane per step
        if (a.llr) {
            real *out = reinterpret_cast<real *>(a.llr) + sp * 1024ull;
#pragma unroll 2
            for (int it = 0; it < 8; it++) {
                const int q = it * 32 + lane;          // quad index in the span: positions 4q..4q+3
                const int word = q >> 3;               // span word holding them (= lane that owns it)
                const uint32_t xb = (__shfl_sync(0xffffffffu, x, word) >> ((q & 7) * 4)) & 0xFu;
                const int fq = word / W;               // frame of the quad within the span
                const unsigned long long Fq = a.first_frame + sp * FPW + fq;
                const uint32_t pq = (uint32_t)(q - fq * (N >> 2));  // quad index within the frame
                const philox4 p = philox4x32_10((uint32_t)Fq, (uint32_t)(Fq >> 32), pq, 0u, (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
                // Box-Muller on (0,1) uniforms
                const float u0 = ((float)p.x + 0.5f) * 2.3283064365386963e-10f, u1 = ((float)p.z + 0.5f) * 2.3283064365386963e-10f;
                const float r0 = sqrtf(-2.0f * logf(u0)), r1 = sqrtf(-2.0f * logf(u1));
                float s0, c0, s1, c1;
                sincospif(2.0f * ((float)p.y * 2.3283064365386963e-10f), &s0, &c0);
                sincospif(2.0f * ((float)p.w * 2.3283064365386963e-10f), &s1, &c1);
                const float z[4] = {r0 * c0, r0 * s0, r1 * c1, r1 * s1};
                const bool span_valid = (sp * FPW + fq) < a.B;
                real o[4];
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const real sgn = ((xb >> e) & 1u) ? (real)-1 : (real)1;
                    if (sizeof(real) == 8) {
                        const double y = (double)sgn + a.sigma_d * (double)z[e];
                        o[e] = (real)(2 * y / a.sigma_d / a.sigma_d);  // SC_128.c:418
                    } else {
                        const float y = (float)sgn + a.sigma_f * z[e];
                        o[e] = (real)(2 * y / a.sigma_f / a.sigma_f);
                    }
                }
                if (span_valid) {
                    if (sizeof(real) == 4) {
                        *reinterpret_cast<float4 *>(out + 4 * q) = make_float4((float)o[0], (float)o[1], (float)o[2], (float)o[3]);
                    } else {
                        double2 *o2 = reinterpret_cast<double2 *>(out + 4 * q);
                        o2[0] = make_double2((double)o[0], (double)o[1]);
                        o2[1] = make_double2((double)o[2], (double)o[3]);
                    }
                }
            }
        }
        __syncwarp();
    }
}

cudaError_t launch_channel(const ChannelArgs &a, bool f64, int sm_count, cudaStream_t st)
{
    // PN-63 period, bit i = PN[i]: b_0=1, b_1..5=0, b_i = b_{i-5} ^ b_{i-6}
    unsigned long long pn = 0;
    {
        int reg[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 63; i++) {
            const int b = (i == 0) ? 1 : (i < 6) ? 0 : (reg[4] ^ reg[5]);
            pn |= (unsigned long long)b << i;
            reg[5] = reg[4]; reg[4] = reg[3]; reg[3] = reg[2]; reg[2] = reg[1]; reg[1] = reg[0]; reg[0] = b;
        }
    }
    const int W = a.N >> 5, FPW = 32 / W;
    const unsigned long long spans = (a.B + FPW - 1) / FPW;
    const int threads = 128, wpb = threads / 32;
    unsigned long long blocks = (spans + wpb - 1) / wpb;
    const unsigned long long cap = (unsigned long long)sm_count * 16;
    if (blocks > cap) blocks = cap;
    if (blocks == 0) return cudaSuccess;
    const size_t smem = (size_t)wpb * 96 * sizeof(uint32_t);
    if (f64) channel_kernel<double><<<(unsigned)blocks, threads, smem, st>>>(a, pn);
    else channel_kernel<float><<<(unsigned)blocks, threads, smem, st>>>(a, pn);
    return cudaGetLastError();
}

// ---------------------------------------------------------------- small helpers
// LLR ingest: element type conversion (float / double / half -> the decoder's arithmetic type) and the optional clipping
// (pg_params.llr_clip > 0) in one pass.  Source and destination may be the same buffer when the types are equal.
template <typename D> __device__ __forceinline__ D clip_to(D v, D c) { return (c > (D)0) ? ((v > c) ? c : ((v < -c) ? -c : v)) : v; }
template <typename S, typename D>
__global__ void convert_kernel(const S *s, D *d, size_t n, D clip)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) d[i] = clip_to<D>((D)s[i], clip);
}

// packed-half LLRs (PG_LLR_F16: half the host-to-device bytes of a streamed frame), 8 values per thread step
template <typename D>
__global__ void convert_h_kernel(const uint4 *__restrict__ s, D *__restrict__ d, size_t n8, const __half *__restrict__ tail_s, size_t n, D clip)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = s[i];
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&w[e]));
            d[i * 8 + 2 * e] = clip_to<D>((D)f.x, clip);
            d[i * 8 + 2 * e + 1] = clip_to<D>((D)f.y, clip);
        }
    }
    if (blockIdx.x == 0)
        for (size_t i = n8 * 8 + threadIdx.x; i < n; i += blockDim.x) d[i] = clip_to<D>((D)__half2float(tail_s[i]), clip);
}

// src_fmt: 0 float, 1 double, 2 half (include/polargpu.h PG_LLR_*); clip <= 0: no clipping (then src_fmt must differ from the destination type)
cudaError_t launch_convert_llr(const void *src, int src_fmt, void *dst, bool dst_f64, size_t count, cudaStream_t st, double clip)
{
    if (count == 0) return cudaSuccess;
    const int threads = 256;
    size_t blocks = (count + threads - 1) / threads;
    if (blocks > 148 * 32) blocks = 148 * 32;
    const unsigned g = (unsigned)blocks;
    if (src_fmt == 2) {
        const size_t n8 = count / 8;
        const unsigned gh = (unsigned)std::max<size_t>(1, std::min<size_t>((n8 + threads - 1) / threads, 148 * 16));
        if (dst_f64) convert_h_kernel<double><<<gh, threads, 0, st>>>((const uint4 *)src, (double *)dst, n8, (const __half *)src, count, clip);
        else convert_h_kernel<float><<<gh, threads, 0, st>>>((const uint4 *)src, (float *)dst, n8, (const __half *)src, count, (float)clip);
    } else if (src_fmt == 1 && !dst_f64) convert_kernel<double, float><<<g, threads, 0, st>>>((const double *)src, (float *)dst, count, (float)clip);
    else if (src_fmt == 0 && dst_f64) convert_kernel<float, double><<<g, threads, 0, st>>>((const float *)src, (double *)dst, count, clip);
    else if (clip > 0 && src_fmt == 0) convert_kernel<float, float><<<g, threads, 0, st>>>((const float *)src, (float *)dst, count, (float)clip);
    else if (clip > 0 && src_fmt == 1) convert_kernel<double, double><<<g, threads, 0, st>>>((const double *)src, (double *)dst, count, clip);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

__global__ void unpack_kernel(const uint32_t *__restrict__ packed, uint8_t *__restrict__ bytes, size_t nbits)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbits; i += (size_t)gridDim.x * blockDim.x)
        bytes[i] = (uint8_t)((packed[i >> 5] >> (i & 31)) & 1u);
}

cudaError_t launch_unpack_bits(const uint32_t *packed, uint8_t *bytes, size_t frames, int N, cudaStream_t st)
{
    const size_t nbits = frames * (size_t)N;
    if (nbits == 0) return cudaSuccess;
    const int threads = 256;
    size_t blocks = (nbits + threads - 1) / threads;
    if (blocks > 148 * 32) blocks = 148 * 32;
    unpack_kernel<<<(unsigned)blocks, threads, 0, st>>>(packed, bytes, nbits);
    return cudaGetLastError();
}

}  // namespace polar
