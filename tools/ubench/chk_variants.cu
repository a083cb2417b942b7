// Micro-benchmark: throughput of candidate CHK implementations on B200 (development aid).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "../../polardecoding_b200/csrc/polar_common.cuh"
using namespace polar;

__device__ __forceinline__ float xsgn(float m, float a, float b) { return __int_as_float(__float_as_int(m) ^ ((__float_as_int(a) ^ __float_as_int(b)) & 0x80000000)); }

// V0: reference-order select chain
__device__ __forceinline__ float chk_v0(float a, float b)
{
    const float delta = tbl8_select_f32(fabsf(a + b)) - tbl8_select_f32(fabsf(a - b));
    return xsgn(fminf(fabsf(a), fabsf(b)), a, b) + delta;
}
// V6: shipped form (FMA-pipe table, bit-identical to V0)
__device__ __forceinline__ float chk_v6(float a, float b) { return chk<float>(a, b); }

// V1: select tree (depth 3)
__device__ __forceinline__ float tbl_tree(float x)
{
    const float lo = (x < 0.433f) ? ((x < 0.196f) ? 0.65f : 0.55f) : ((x < 0.71f) ? 0.45f : 0.35f);
    const float hi = (x < 2.252f) ? ((x < 1.508f) ? 0.25f : 0.15f) : ((x < 4.5f) ? 0.05f : 0.0f);
    return (x < 1.05f) ? lo : hi;
}
__device__ __forceinline__ float chk_v1(float a, float b)
{
    const float delta = tbl_tree(fabsf(a + b)) - tbl_tree(fabsf(a - b));
    return xsgn(fminf(fabsf(a), fabsf(b)), a, b) + delta;
}

// V2: steps on the FMA pipe: step_i(x) = sat((x - pred(t_i)) * 2^60) = [x >= t_i]; T = 0.05*(13 - sum w_i step_i)
__device__ __forceinline__ float stepsum(float x)
{
    const float BIG = 1.152921504606846976e18f;  // 2^60
    // pred(t) * 2^60 as constants: computed at compile time from nextafterf(t, 0)
    float acc;
    acc = __saturatef(fmaf(x, BIG, -(__int_as_float(__float_as_int(4.5f) - 1) * BIG)));
    acc = fmaf(__saturatef(fmaf(x, BIG, -(__int_as_float(__float_as_int(2.252f) - 1) * BIG))), 2.0f, acc);
    acc = fmaf(__saturatef(fmaf(x, BIG, -(__int_as_float(__float_as_int(1.508f) - 1) * BIG))), 2.0f, acc);
    acc = fmaf(__saturatef(fmaf(x, BIG, -(__int_as_float(__float_as_int(1.05f) - 1) * BIG))), 2.0f, acc);
    acc = fmaf(__saturatef(fmaf(x, BIG, -(__int_as_float(__float_as_int(0.71f) - 1) * BIG))), 2.0f, acc);
    acc = fmaf(__saturatef(fmaf(x, BIG, -(__int_as_float(__float_as_int(0.433f) - 1) * BIG))), 2.0f, acc);
    acc = fmaf(__saturatef(fmaf(x, BIG, -(__int_as_float(__float_as_int(0.196f) - 1) * BIG))), 2.0f, acc);
    return acc;  // = 13 - T/0.05
}
__device__ __forceinline__ float chk_v2(float a, float b)
{
    const float m = stepsum(fabsf(a - b)) - stepsum(fabsf(a + b));  // (T(s)-T(d))/0.05
    return fmaf(m, 0.05f, xsgn(fminf(fabsf(a), fabsf(b)), a, b));
}

// V3: s-table on the FMA pipe, d-table as a select tree on the ALU pipe
__device__ __forceinline__ float chk_v3(float a, float b)
{
    const float ts = fmaf(stepsum(fabsf(a + b)), -0.05f, 0.65f);
    const float delta = ts - tbl_tree(fabsf(a - b));
    return xsgn(fminf(fabsf(a), fabsf(b)), a, b) + delta;
}

// V4: FMA steps with FADD accumulation in a tree (no weights: 13 unit steps would be needed) -- instead both tables share one accumulator
__device__ __forceinline__ float chk_v4(float a, float b)
{
    const float BIG = 1.152921504606846976e18f;
    const float s = fabsf(a + b), d = fabsf(a - b);
    float acc = 0.f;
#define STEP2(T, Wt) { const float c = -(__int_as_float(__float_as_int(T) - 1) * BIG); \
        acc = fmaf(__saturatef(fmaf(d, BIG, c)), Wt, acc); acc = fmaf(__saturatef(fmaf(s, BIG, c)), -Wt, acc); }
    STEP2(4.5f, 0.05f) STEP2(2.252f, 0.1f) STEP2(1.508f, 0.1f) STEP2(1.05f, 0.1f) STEP2(0.71f, 0.1f) STEP2(0.433f, 0.1f) STEP2(0.196f, 0.1f)
#undef STEP2
    return xsgn(fminf(fabsf(a), fabsf(b)), a, b) + acc;
}

// V5: half2 compares (NOT exact near thresholds) -- speed reference only
__device__ __forceinline__ float chk_v5(float a, float b)
{
    const __half2 x = __floats2half2_rn(fabsf(a + b), fabsf(a - b));
    __half2 acc = __hlt2(x, __float2half2_rn(4.5f));
    acc = __hfma2(__hlt2(x, __float2half2_rn(2.252f)), __float2half2_rn(2.f), acc);
    acc = __hfma2(__hlt2(x, __float2half2_rn(1.508f)), __float2half2_rn(2.f), acc);
    acc = __hfma2(__hlt2(x, __float2half2_rn(1.05f)), __float2half2_rn(2.f), acc);
    acc = __hfma2(__hlt2(x, __float2half2_rn(0.71f)), __float2half2_rn(2.f), acc);
    acc = __hfma2(__hlt2(x, __float2half2_rn(0.433f)), __float2half2_rn(2.f), acc);
    acc = __hfma2(__hlt2(x, __float2half2_rn(0.196f)), __float2half2_rn(2.f), acc);
    const float m = __low2float(acc) - __high2float(acc);
    return fmaf(m, 0.05f, xsgn(fminf(fabsf(a), fabsf(b)), a, b));
}

// V7: steps as scalar FFMA.SAT, the two table sums accumulated together with packed fp32x2 FMAs (Blackwell FFMA2)
__device__ __forceinline__ unsigned long long pk(float lo, float hi)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b)
{
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float chk_v7(float a, float b)
{
    const float NB = -1.152921504606846976e18f, B = 1.152921504606846976e18f;
    const float s = fabsf(a + b), d = fabsf(a - b);
#define ST(x, T) __saturatef(fmaf(x, NB, T * B))
#define INC(h) pk(__int_as_float(h), __int_as_float(h))
    unsigned long long acc = mul2(pk(ST(s, 4.5f), ST(d, 4.5f)), INC(0x3d4ccccd));
    acc = fma2(pk(ST(s, 2.252f), ST(d, 2.252f)), INC(0x3dccccce), acc);
    acc = fma2(pk(ST(s, 1.508f), ST(d, 1.508f)), INC(0x3dcccccc), acc);
    acc = fma2(pk(ST(s, 1.05f), ST(d, 1.05f)), INC(0x3dcccccc), acc);
    acc = fma2(pk(ST(s, 0.71f), ST(d, 0.71f)), INC(0x3dcccccc), acc);
    acc = fma2(pk(ST(s, 0.433f), ST(d, 0.433f)), INC(0x3dccccd0), acc);
    acc = fma2(pk(ST(s, 0.196f), ST(d, 0.196f)), INC(0x3dccccc8), acc);
#undef ST
#undef INC
    float ts, td;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(ts), "=f"(td) : "l"(acc));
    return xsgn(fminf(fabsf(a), fabsf(b)), a, b) + (ts - td);
}

template <int V> __device__ __forceinline__ float chkv(float a, float b)
{
    if (V == 0) return chk_v0(a, b);
    if (V == 1) return chk_v1(a, b);
    if (V == 2) return chk_v2(a, b);
    if (V == 3) return chk_v3(a, b);
    if (V == 4) return chk_v4(a, b);
    if (V == 6) return chk_v6(a, b);
    if (V == 7) return chk_v7(a, b);
    return chk_v5(a, b);
}

template <int V, int ILP>
__global__ void __launch_bounds__(256) bench(const float *in, float *out, int iters)
{
    float a[ILP], c[ILP];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int i = 0; i < ILP; i++) { a[i] = in[(t * ILP + i) & 65535]; c[i] = in[(t * ILP + i + 7) & 65535]; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) { a[i] = chkv<V>(a[i], c[i]) + c[i]; c[i] = -c[i] * 1.0001f; }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += a[i];
    out[t] = s;
}

template <int V>
__global__ void check(const float *in, int n, unsigned long long *bad, float *maxerr)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const float a = in[t], b = in[(t * 7 + 3) % n];
    const float r0 = chk_v0(a, b), r = chkv<V>(a, b);
    if (r0 != r) { atomicAdd(bad, 1ull); const float e = fabsf(r0 - r); atomicMax((int *)maxerr, __float_as_int(e)); }
}

template <int V, int ILP> void run(const float *din, float *dout, const char *name)
{
    const int blocks = 148 * 8, threads = 256, iters = 2000;
    bench<V, ILP><<<blocks, threads>>>(din, dout, 10);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<V, ILP><<<blocks, threads>>>(din, dout, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double n = (double)blocks * threads * ILP * iters;
    unsigned long long *dbad; float *dmax; cudaMalloc(&dbad, 8); cudaMalloc(&dmax, 4); cudaMemset(dbad, 0, 8); cudaMemset(dmax, 0, 4);
    check<V><<<65536 / 256, 256>>>(din, 65536, dbad, dmax);
    unsigned long long bad; float mx; cudaMemcpy(&bad, dbad, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&mx, dmax, 4, cudaMemcpyDeviceToHost);
    printf("%-28s ILP %d: %.3f ms  %.2f G CHK/s  -> %.1f %% of 37.2T lane-op roofline at 27 ops/CHK | differs from V0 in %llu/65536 (max |d| %.3g)\n",
           name, ILP, ms, n / ms / 1e6, n * 27 / (ms * 1e-3) / 37.22e12 * 100, bad, mx);
}

int main()
{
    float *h = (float *)malloc(65536 * 4);
    srand(1);
    for (int i = 0; i < 65536; i++) { float u = (rand() / (float)RAND_MAX - 0.5f) * 12.f; h[i] = u; }
    h[0] = 0.196f; h[1] = 0.f; h[2] = 4.5f; h[3] = 0.f;
    {   // values within a few ulps of every threshold, as |a+b| with b = 0
        const float th[7] = {0.196f, 0.433f, 0.71f, 1.05f, 1.508f, 2.252f, 4.5f};
        int k = 8;
        for (int t = 0; t < 7; t++)
            for (int d = -3; d <= 3; d++) { int bits; memcpy(&bits, &th[t], 4); bits += d; float v; memcpy(&v, &bits, 4); h[k++] = v; h[k++] = -v; }
    }
    float *din, *dout; cudaMalloc(&din, 65536 * 4); cudaMalloc(&dout, 148 * 8 * 256 * 4);
    cudaMemcpy(din, h, 65536 * 4, cudaMemcpyHostToDevice);
    run<0, 1>(din, dout, "V0 select chain"); run<0, 4>(din, dout, "V0 select chain");
    run<1, 1>(din, dout, "V1 select tree"); run<1, 4>(din, dout, "V1 select tree");
    run<2, 1>(din, dout, "V2 FMA.SAT steps"); run<2, 4>(din, dout, "V2 FMA.SAT steps");
    run<3, 1>(din, dout, "V3 s:FMA d:tree"); run<3, 4>(din, dout, "V3 s:FMA d:tree");
    run<4, 1>(din, dout, "V4 FMA.SAT shared acc"); run<4, 4>(din, dout, "V4 FMA.SAT shared acc");
    run<5, 4>(din, dout, "V5 half2 (inexact)");
    run<6, 1>(din, dout, "V6 shipped (FMA, exact)"); run<6, 2>(din, dout, "V6 shipped (FMA, exact)"); run<6, 4>(din, dout, "V6 shipped (FMA, exact)");
    run<7, 1>(din, dout, "V7 FFMA2 accumulate (exact)"); run<7, 2>(din, dout, "V7 FFMA2 accumulate (exact)"); run<7, 4>(din, dout, "V7 FFMA2 accumulate (exact)");
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return 0;
}
