// Internal (C++) interface between the C-ABI layer (api.cu) and the kernel translation units.
#pragma once
#ifdef POLAR_EMU  // CPU warp emulator, test infrastructure only (tests/emu)
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>

namespace polar {

constexpr int kMaxWords = 32;  // N <= 1024 -> at most 32 packed words per frame

// per-frame result word written by the decode kernels
//   bits 0..15  wrong counted bits (only when a truth vector was supplied)
//   bit 16      list decoders: exact PM tie at the list boundary in this frame
//   bit 17      CA-SCL: no path passed the CRC
//   bits 24..31 BP: sweeps executed
constexpr uint32_t kInfoTie = 1u << 16;
constexpr uint32_t kInfoCrcFail = 1u << 17;

// indices into the device counter vector (mirrors pg_counters)
enum { CNT_FRAMES = 0, CNT_ERR_BLOCKS = 1, CNT_ERR_BITS = 2, CNT_TIE = 3, CNT_CRCFAIL = 4, CNT_SWEEPS = 5, CNT_N = 8 };

struct CodeMasks {
    uint32_t info[kMaxWords];  // bit p set: position p carries payload or CRC (non-frozen)
    uint32_t cnt[kMaxWords];   // bit p set: position p enters the error count
};

// ---------------------------------------------------------------- channel kernel
struct ChannelArgs {
    void *llr;                 // [B][N] real (float or double)
    uint32_t *u_packed;        // [B][N/32] truth
    const uint16_t *I;         // [nI] reliability-ordered non-frozen positions
    const uint32_t *crc_sys;   // [K] D^(r+i) mod g (systematic CRC) or null
    unsigned long long first_frame, B;
    uint64_t crc_poly, seed;
    int N, n, K, r, nI, crc_systematic, data_mode;
    float sigma_f;
    double sigma_d;
};
cudaError_t launch_channel(const ChannelArgs &a, bool f64, int sm_count, cudaStream_t st);

// ---------------------------------------------------------------- list decoders (SC = L 1)
struct ListArgs {
    const void *llr;                 // [B][N] real
    const uint32_t *truth;           // [B][N/32] or null
    uint32_t *u_hat;                 // [B][N/32] or null
    uint32_t *frame_info;            // [B] or null
    unsigned long long *counters;    // CNT_N u64 or null
    void *gscratch;                  // per-CTA stage scratch (reals)
    const uint32_t *crc_masks;       // [r][N/32] u-domain syndrome masks
    unsigned long long B;
    int r, use_crc;
    int coop_groups;                 // 4-bit groups before the first non-frozen bit: all paths of a frame are still identical
    CodeMasks m;
};
// returns scratch bytes one CTA needs / smem bytes / max resident CTAs per SM for a configuration, or <0 if not compiled
struct ListPlan { size_t scratch_per_cta; size_t smem; int ctas_per_sm; int frames_per_cta; };
cudaError_t list_plan(int n, int L, bool f64, ListPlan *plan);
cudaError_t launch_list(const ListArgs &a, int n, int L, bool f64, int grid, cudaStream_t st);

// ---------------------------------------------------------------- BP
struct BpArgs {
    const void *llr;
    const uint32_t *truth;
    uint32_t *u_hat;
    uint32_t *frame_info;
    unsigned long long *counters;
    unsigned long long *queue;       // dynamic frame queue (one u64, zeroed before launch)
    unsigned long long *bpr_E;       // [ns][n+1] or null
    unsigned long long B;
    int iters, early_stop;
    int bpr_ns;
    int bpr_samples[8];
    CodeMasks m;
};
struct BpPlan { size_t smem; int ctas_per_sm; int threads; };
cudaError_t bp_plan(int n, bool f64, BpPlan *plan);
cudaError_t launch_bp(const BpArgs &a, int n, bool f64, int grid, cudaStream_t st);
// packed-half variant (two frames per CTA, float LLRs in; no BPR statistic)
cudaError_t bp_h2_plan(int n, BpPlan *plan);
cudaError_t launch_bp_h2(const BpArgs &a, int n, int grid, cudaStream_t st);

// ---------------------------------------------------------------- helpers
cudaError_t launch_convert_llr(const void *src, int src_fmt /* 0 float, 1 double, 2 half */, void *dst, bool dst_f64, size_t count, cudaStream_t st,
                               double clip = 0.0 /* > 0: clip to [-clip, clip]; then src and dst may be the same type / buffer */);
cudaError_t launch_unpack_bits(const uint32_t *packed, uint8_t *bytes, size_t frames, int N, cudaStream_t st);

}  // namespace polar
