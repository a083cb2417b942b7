// TEST INFRASTRUCTURE: a single-threaded warp emulator that lets g++ compile the CUDA kernel sources of
// polardecoding_b200/csrc unchanged (with -DPOLAR_EMU) and run them on the CPU, so that kernel LOGIC (schedules,
// pointer bookkeeping, warp collectives, shared-memory / tensor-memory layouts) can be checked against the oracle in the
// build container, which has no GPU.  It is never part of the product: nothing under polardecoding_b200/ is built with
// POLAR_EMU except by tests/emu/Makefile.
//
// Model: the 32 lanes of a warp are ucontext coroutines scheduled round-robin.  Every warp collective (shuffle, vote,
// __syncwarp) deposits the lane's operand, yields, and reads the other lanes' operands after the whole warp has arrived;
// the emulator aborts if the lanes of a warp reach different collectives (divergence the hardware would not forgive).
// Warps of a CTA run one after the other (the kernels here only use __syncthreads around the tensor-memory allocation).
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <algorithm>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __restrict__

typedef int cudaError_t;
typedef void *cudaStream_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1 };

struct float4 { float x, y, z, w; } __attribute__((aligned(16)));
struct float2 { float x, y; } __attribute__((aligned(8)));
struct double2 { double x, y; } __attribute__((aligned(16)));
struct uint4 { uint32_t x, y, z, w; } __attribute__((aligned(16)));
struct uint2 { uint32_t x, y; } __attribute__((aligned(8)));
static inline float4 make_float4(float a, float b, float c, float d) { float4 r; r.x = a; r.y = b; r.z = c; r.w = d; return r; }
static inline float2 make_float2(float a, float b) { float2 r; r.x = a; r.y = b; return r; }
static inline double2 make_double2(double a, double b) { double2 r; r.x = a; r.y = b; return r; }
static inline uint4 make_uint4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { uint4 r; r.x = a; r.y = b; r.z = c; r.w = d; return r; }

namespace emu {

struct dim3_t { unsigned x, y, z; };

struct Warp {
    ucontext_t main_ctx, lane_ctx[32];
    std::vector<char> stacks;
    bool done[32];
    int cur = 0;                 // lane being run
    unsigned long long seq[32];  // collectives executed per lane
    int site[2][32];             // call-site id of the pending collective (parity buffers)
    uint64_t val[2][32];
    uint64_t aux[2][32];
    int warp_in_cta = 0;
};

struct State {
    Warp *w = nullptr;
    dim3_t block_idx{0, 0, 0}, grid_dim{1, 1, 1}, block_dim{32, 1, 1};
    unsigned char *smem = nullptr;
    uint32_t *tmem = nullptr;      // [128 lanes][512 columns]
    uint32_t tmem_next_col = 0;
    unsigned long long collectives = 0;
};
extern State g;

struct tid_proxy { operator dim3_t() const; unsigned get_x() const; };

inline int lane_id() { return g.w->cur; }
inline unsigned thread_x() { return (unsigned)(g.w->warp_in_cta * 32 + g.w->cur); }

inline void yield_lane()
{
    Warp *w = g.w;
    swapcontext(&w->lane_ctx[w->cur], &w->main_ctx);
}

// deposit (v, a) under call-site id `site`, wait for the whole warp, leave the parity index in *par
inline int arrive(int site, uint64_t v, uint64_t a = 0)
{
    Warp *w = g.w;
    const int l = w->cur;
    const int par = (int)(w->seq[l] & 1ull);
    w->site[par][l] = site;
    w->val[par][l] = v;
    w->aux[par][l] = a;
    w->seq[l]++;
    if (l == 0) g.collectives++;
    yield_lane();
    for (int i = 0; i < 32; i++) {
        if (w->seq[i] > w->seq[l]) continue;  // lane i already passed this collective (checked it itself) and moved on
        if (w->done[i] && w->seq[i] < w->seq[l]) { fprintf(stderr, "emu: lane %d exited while lane %d waits in a collective (site %d)\n", i, l, site); abort(); }
        if (w->site[par][i] != site) { fprintf(stderr, "emu: divergent collectives: lane %d at site %d, lane %d at site %d\n", l, site, i, w->site[par][i]); abort(); }
    }
    return par;
}

template <typename T> inline uint64_t to_bits(T v) { uint64_t b = 0; static_assert(sizeof(T) <= 8, "size"); memcpy(&b, &v, sizeof(T)); return b; }
template <typename T> inline T from_bits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }

template <typename T> inline T shfl_idx(int site, T v, int src, int width)
{
    const int l = lane_id();
    const int par = arrive(site, to_bits(v));
    const int base = l & ~(width - 1);
    const int s = base + (src & (width - 1));
    return from_bits<T>(g.w->val[par][s]);
}
template <typename T> inline T shfl_xor(int site, T v, int m, int width)
{
    const int l = lane_id();
    const int par = arrive(site, to_bits(v));
    int s = l ^ m;
    if ((s & ~(width - 1)) != (l & ~(width - 1))) s = l;  // out of the segment: own value
    return from_bits<T>(g.w->val[par][s]);
}
inline unsigned ballot(int site, int pred)
{
    const int par = arrive(site, (uint64_t)(pred != 0));
    unsigned m = 0;
    for (int i = 0; i < 32; i++) m |= (unsigned)(g.w->val[par][i] & 1ull) << i;
    return m;
}

void run_grid(void (*lane_fn)(void *), void *arg, unsigned grid, unsigned block_threads, size_t smem_bytes);

}  // namespace emu

struct emu_idx_x { operator unsigned() const { return emu::thread_x(); } };
struct emu_tidx { emu_idx_x x; };
static const emu_tidx threadIdx{};
#define blockIdx (emu::g.block_idx)
#define gridDim (emu::g.grid_dim)
#define blockDim (emu::g.block_dim)

#define EMU_SITE (__LINE__ * 8 + 1)
#define __shfl_sync(mask, v, ...) emu_shfl_sync(EMU_SITE, v, __VA_ARGS__)
#define __shfl_xor_sync(mask, v, ...) emu_shfl_xor_sync(EMU_SITE + 1, v, __VA_ARGS__)
#define __ballot_sync(mask, p) emu::ballot(EMU_SITE + 2, (p))
#define __all_sync(mask, p) (emu::ballot(EMU_SITE + 3, (p)) == 0xffffffffu)
#define __any_sync(mask, p) (emu::ballot(EMU_SITE + 4, (p)) != 0u)
#define __syncwarp(...) ((void)emu::arrive(EMU_SITE + 5, 0))
#define __syncthreads() ((void)0)

template <typename T> inline T emu_shfl_sync(int site, T v, int src, int width = 32) { return emu::shfl_idx<T>(site, v, src, width); }
template <typename T> inline T emu_shfl_xor_sync(int site, T v, int m, int width = 32) { return emu::shfl_xor<T>(site, v, m, width); }

// ---- bit / conversion intrinsics
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
static inline long long __double_as_longlong(double d) { long long i; memcpy(&i, &d, 8); return i; }
static inline double __longlong_as_double(long long i) { double d; memcpy(&d, &i, 8); return d; }
static inline int __double2hiint(double d) { return (int)(__double_as_longlong(d) >> 32); }
static inline int __double2loint(double d) { return (int)(__double_as_longlong(d) & 0xffffffffll); }
static inline double __hiloint2double(int hi, int lo) { return __longlong_as_double((long long)(((unsigned long long)(unsigned)hi << 32) | (unsigned)lo)); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
static inline float __saturatef(float x) { return (x != x) ? 0.f : (x < 0.f ? 0.f : (x > 1.f ? 1.f : x)); }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { unsigned o = *p; *p = o + v; return o; }

// ---- tensor memory (per CTA: 128 lanes x 512 columns of 32 bits); a warp reaches the lane quarter (warp % 4) only
namespace emu {
inline uint32_t *tmem_cell(uint32_t taddr, int e)
{
    const uint32_t lane_base = taddr >> 16, col = (taddr & 0xffffu) + (uint32_t)e;
    if (lane_base != (uint32_t)(g.w->warp_in_cta % 4) * 32u) { fprintf(stderr, "emu: warp %d touches tensor-memory lanes %u..\n", g.w->warp_in_cta, lane_base); abort(); }
    if (col >= g.tmem_next_col) { fprintf(stderr, "emu: tensor-memory column %u outside the allocation (%u)\n", col, g.tmem_next_col); abort(); }
    return g.tmem + (size_t)(lane_base + (uint32_t)lane_id()) * 512u + col;
}
inline uint32_t tmem_alloc(uint32_t cols)
{
    const uint32_t base = g.tmem_next_col;
    if (base + cols > 512u) { fprintf(stderr, "emu: tensor memory exhausted\n"); abort(); }
    g.tmem_next_col += cols;
    return base;
}
}  // namespace emu
