// C ABI of libpolargpu.so (see include/polargpu.h): context, buffers, batching, the Monte-Carlo loop with
// the reference's stopping rule, and the counter exchange between ranks.  Host logic only -- every
// arithmetic step of the hot path is in channel.cu / list_decode.cu / bp_decode.cu.  No CPU fallback.
#include "../../include/polargpu.h"
#include "../../include/polar_q_table.h"
#include "engine.h"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace polar;

namespace {

thread_local std::string g_create_error;

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool load(std::string &err)
    {
        if (handle) return true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {
            handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) { err = std::string("dlopen libnccl failed: ") + dlerror(); return false; }
        GetUniqueId = (decltype(GetUniqueId))dlsym(handle, "ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))dlsym(handle, "ncclCommInitRank");
        AllReduce = (decltype(AllReduce))dlsym(handle, "ncclAllReduce");
        CommDestroy = (decltype(CommDestroy))dlsym(handle, "ncclCommDestroy");
        GetErrorString = (decltype(GetErrorString))dlsym(handle, "ncclGetErrorString");
        if (!GetUniqueId || !CommInitRank || !AllReduce || !CommDestroy) { err = "libnccl lacks required symbols"; return false; }
        return true;
    }
};
NcclApi g_nccl;

int ilog2(int v) { int k = 0; while ((1 << k) < v) k++; return k; }
// LLR element formats of the streaming calls (PG_LLR_*)
inline bool llr_fmt_ok(int f) { return f == PG_LLR_F32 || f == PG_LLR_F64 || f == PG_LLR_F16; }
inline size_t llr_esz(int f) { return f == PG_LLR_F64 ? 8 : (f == PG_LLR_F16 ? 2 : 4); }

// the public PG_INFO_* layout (include/polargpu.h) is the kernels' frame-info word (engine.h)
static_assert(PG_INFO_TIE == polar::kInfoTie && PG_INFO_CRC_FAIL == polar::kInfoCrcFail, "frame-info layout");

}  // namespace

constexpr int kPipeLanes = 4, kPipeBufs = 2 * kPipeLanes;

struct pg_ctx {
    pg_params p;
    int n = 0, W = 0, nI = 0;
    bool f64 = false;
    bool h2 = false;           // BP with packed-half messages (PG_REAL_H2); device LLRs are float
    int sm_count = 0;
    std::vector<int> I;
    std::vector<uint8_t> inI;
    CodeMasks masks;
    std::string err;

    cudaStream_t st = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};  // channel start/stop, decode start/stop
    bool ev_ch = false, ev_dec = false, ev_suppress = false;
    uint64_t launches = 0;

    // device tables
    uint16_t *d_I = nullptr;
    uint32_t *d_crc_masks = nullptr, *d_crc_sys = nullptr;
    unsigned long long *d_counters = nullptr, *d_queue = nullptr, *d_bpr = nullptr;
    unsigned long long *d_cnt2 = nullptr;   // counters of the calls that return their own sums (pg_simulate*, pg_decode_llr_counted);
                                            // d_counters belongs to pg_decode_count_device / pg_counters_read alone
    int bpr_ns = 0, bpr_samples[8] = {0};

    // work buffers for `cap` frames
    size_t cap = 0;
    void *d_llr = nullptr;     // real
    void *d_in = nullptr;      // staging for caller LLRs of the other type
    uint32_t *d_truth = nullptr, *d_uhat = nullptr, *d_info = nullptr;
    uint8_t *d_bytes = nullptr;
    uint32_t *h_info = nullptr;             // pinned
    unsigned long long *h_counters = nullptr;  // pinned, CNT_N
    void *d_scratch = nullptr;
    int grid = 0;
    size_t scratch_per_cta = 0;
    size_t frames_per_cta = 1;  // frames one CTA of the decode kernel holds at a time
    // pipelined host path (pg_decode_llr*): a copy stream feeds two buffer sets per lane; the list decoders run kPipeLanes
    // quarter-grid kernels side by side on as many compute streams (their phases interleave instead of marching in lockstep),
    // BP runs one
    cudaStream_t st_copy = nullptr, st_lane[kPipeLanes] = {};  // st_lane[0] is st
    cudaEvent_t ev_h2d[kPipeBufs] = {}, ev_free[kPipeBufs] = {};
    cudaEvent_t ev_fork = nullptr, ev_join[kPipeLanes] = {};
    size_t pipe_cap = 0;
    bool pipe_ready = false;
    void *p_llr[kPipeBufs] = {}, *p_in[kPipeBufs] = {};
    uint32_t *p_uhat[kPipeBufs] = {}, *p_info[kPipeBufs] = {};
    uint8_t *p_bytes[kPipeBufs] = {};
    size_t chunk_max = 0;

    ncclComm_t comm = nullptr;
    unsigned long long *d_xchg = nullptr;  // nranks*CNT_N
    // pipelined Monte-Carlo loop (pg_simulate): a ring of round slots, the counter exchange on its own stream
    static constexpr int kRing = 4;
    cudaStream_t st_comm = nullptr;
    unsigned long long *d_rc[kRing] = {};      // this rank's counters of the round in the slot (CNT_N)
    unsigned long long *d_xm[kRing] = {};      // [nranks][CNT_N] exchange matrix of the round
    unsigned long long *h_xm[kRing] = {};      // pinned copy of the reduced matrix
    uint32_t *d_rinfo[kRing] = {};             // per-frame words of the round (exact stop), ring_cap frames each
    size_t ring_cap = 0;
    cudaEvent_t ev_round[kRing] = {}, ev_xdone[kRing] = {};
    uint64_t sim_rounds = 0, sim_allreduces = 0;   // statistics of the last pg_simulate call
};

#define CU(call)                                                                                 \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess) {                                                                 \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                       \
            return PG_ERR_CUDA;                                                                  \
        }                                                                                        \
    } while (0)

extern "C" const char *pg_version(void) { return "polargpu 0.2 (sm_100a)"; }

extern "C" int pg_device_count(void)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess) return 0;
    int ok = 0;
    for (int d = 0; d < ndev; d++) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, d) == cudaSuccess && prop.major == 10) ok++;
        else break;  // usable devices must be the ordinals 0..ok-1 (rank r runs on device r)
    }
    return ok;
}

extern "C" int pg_params_preset(pg_params *p, const char *prog)
{
    if (!p || !prog) return PG_ERR_ARG;
    std::memset(p, 0, sizeof(*p));
    p->list_size = 1;
    p->real = PG_REAL_F32;
    p->data_mode = PG_DATA_PN63;
    p->seed = 1024;
    p->nranks = 1;
    const std::string s(prog);
    auto code = [&](int N, int K, int r, uint64_t poly, int sys) { p->N = N; p->K = K; p->crc_bits = r; p->crc_poly = poly; p->crc_systematic = sys; };
    if (s == "SC_128" || s == "SC_128_fag") { code(128, 64, 0, 0, 0); p->decoder = PG_DEC_SC; }
    else if (s == "SC_1024") { code(1024, 512, 0, 0, 0); p->decoder = PG_DEC_SC; }
    else if (s == "SCL_128" || s == "SCL_128_fag") { code(128, 64, 0, 0, 0); p->decoder = PG_DEC_SCL; p->list_size = 8; }
    else if (s == "SCL_1024") { code(1024, 512, 0, 0, 0); p->decoder = PG_DEC_SCL; p->list_size = 8; }
    else if (s == "CASCL_128") { code(128, 64, 6, PG_CRC6_POLY, 0); p->decoder = PG_DEC_CASCL; p->list_size = 8; }
    else if (s == "CASCL_128_sys") { code(128, 64, 6, PG_CRC6_POLY, 1); p->decoder = PG_DEC_CASCL; p->list_size = 8; p->count_from = 6; }
    else if (s == "CASCL_1024_L8") { code(1024, 512, 24, PG_CRC24_POLY, 0); p->decoder = PG_DEC_CASCL; p->list_size = 8; }
    else if (s == "CASCL_1024_sys") { code(1024, 512, 24, PG_CRC24_POLY, 1); p->decoder = PG_DEC_CASCL; p->list_size = 8; p->count_from = 24; }
    else if (s == "BP_128" || s == "BP_128_fag") { code(128, 64, 0, 0, 0); p->decoder = PG_DEC_BP; p->iter_max = 100; }
    else if (s == "BP_1024") { code(1024, 512, 0, 0, 0); p->decoder = PG_DEC_BP; p->iter_max = 100; }
    else if (s == "BPr_128") { code(128, 64, 0, 0, 0); p->decoder = PG_DEC_BP; p->iter_max = 90; }
    else return PG_ERR_ARG;
    return PG_OK;
}

static int build_code(pg_ctx *ctx)
{
    const pg_params &p = ctx->p;
    ctx->n = ilog2(p.N);
    ctx->W = p.N / 32;
    ctx->nI = p.K + p.crc_bits;
    std::vector<int> q;
    for (int i = 0; i < POLAR_Q_TABLE_LEN; i++)
        if (polar_q_table_1024[i] < p.N) q.push_back(polar_q_table_1024[i]);
    ctx->I.resize(ctx->nI);
    ctx->inI.assign(p.N, 0);
    for (int i = 0; i < ctx->nI; i++) { ctx->I[i] = q[p.N - ctx->nI + i]; ctx->inI[ctx->I[i]] = 1; }  // SC_128.c:143-147
    std::memset(&ctx->masks, 0, sizeof(ctx->masks));
    for (int i = 0; i < ctx->nI; i++) {
        const int pos = ctx->I[i];
        ctx->masks.info[pos >> 5] |= 1u << (pos & 31);
        if (i >= p.count_from) ctx->masks.cnt[pos >> 5] |= 1u << (pos & 31);
    }
    return PG_OK;
}

static void free_buffers(pg_ctx *ctx)
{
    cudaFree(ctx->d_llr); cudaFree(ctx->d_in); cudaFree(ctx->d_truth); cudaFree(ctx->d_uhat); cudaFree(ctx->d_info);
    cudaFree(ctx->d_bytes);
    if (ctx->h_info) cudaFreeHost(ctx->h_info);
    ctx->d_llr = ctx->d_in = nullptr; ctx->d_truth = ctx->d_uhat = ctx->d_info = nullptr; ctx->d_bytes = nullptr; ctx->h_info = nullptr;
    ctx->cap = 0;
}

static int ensure_lanes(pg_ctx *ctx)
{
    if (!ctx->pipe_ready) {  // streams and events: all or nothing (a partial set is completed on the next call)
        if (!ctx->st_copy) CU(cudaStreamCreateWithFlags(&ctx->st_copy, cudaStreamNonBlocking));
        ctx->st_lane[0] = ctx->st;
        for (int l = 1; l < kPipeLanes; l++)
            if (!ctx->st_lane[l]) CU(cudaStreamCreateWithFlags(&ctx->st_lane[l], cudaStreamNonBlocking));
        for (int s = 0; s < kPipeBufs; s++) {
            if (!ctx->ev_h2d[s]) CU(cudaEventCreateWithFlags(&ctx->ev_h2d[s], cudaEventDisableTiming));
            if (!ctx->ev_free[s]) CU(cudaEventCreateWithFlags(&ctx->ev_free[s], cudaEventDisableTiming));
        }
        for (int l = 0; l < kPipeLanes; l++)
            if (!ctx->ev_join[l]) CU(cudaEventCreateWithFlags(&ctx->ev_join[l], cudaEventDisableTiming));
        if (!ctx->ev_fork) CU(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
        ctx->pipe_ready = true;
    }
    return PG_OK;
}

static int ensure_pipe(pg_ctx *ctx, size_t frames, bool want_bytes)
{
    int rcl = ensure_lanes(ctx);
    if (rcl) return rcl;
    if (frames <= ctx->pipe_cap && (!want_bytes || ctx->p_bytes[0])) return PG_OK;
    ctx->pipe_cap = 0;  // nothing is usable until every buffer set below exists at the new size
    const size_t N = ctx->p.N, W = ctx->W;
    for (int s = 0; s < kPipeBufs; s++) {
        cudaFree(ctx->p_llr[s]); cudaFree(ctx->p_in[s]); cudaFree(ctx->p_uhat[s]); cudaFree(ctx->p_info[s]); cudaFree(ctx->p_bytes[s]);
        ctx->p_llr[s] = ctx->p_in[s] = nullptr; ctx->p_uhat[s] = ctx->p_info[s] = nullptr; ctx->p_bytes[s] = nullptr;
        CU(cudaMalloc(&ctx->p_llr[s], frames * N * (ctx->f64 ? 8 : 4)));
        CU(cudaMalloc(&ctx->p_in[s], frames * N * 8));
        CU(cudaMalloc(&ctx->p_uhat[s], frames * W * 4));
        CU(cudaMalloc(&ctx->p_info[s], frames * 4));
        if (want_bytes) CU(cudaMalloc(&ctx->p_bytes[s], frames * N));
    }
    ctx->pipe_cap = frames;
    return PG_OK;
}

static int ensure_capacity(pg_ctx *ctx, size_t frames)
{
    if (frames <= ctx->cap) return PG_OK;
    free_buffers(ctx);
    const size_t N = ctx->p.N, W = ctx->W;
    CU(cudaMalloc(&ctx->d_llr, frames * N * (ctx->f64 ? 8 : 4)));
    CU(cudaMalloc(&ctx->d_in, frames * N * 8));
    CU(cudaMalloc(&ctx->d_truth, frames * W * 4));
    CU(cudaMalloc(&ctx->d_uhat, frames * W * 4));
    CU(cudaMalloc(&ctx->d_info, frames * 4));
    CU(cudaMalloc(&ctx->d_bytes, frames * N));
    CU(cudaMallocHost(&ctx->h_info, frames * 4));
    ctx->cap = frames;
    return PG_OK;
}

extern "C" int pg_create(const pg_params *p, pg_ctx **out)
{
    if (!p || !out) return PG_ERR_ARG;
    *out = nullptr;
    auto fail = [&](int code, const std::string &msg) { g_create_error = msg; return code; };
    if (p->N < 32 || p->N > 1024 || (p->N & (p->N - 1))) return fail(PG_ERR_ARG, "N must be a power of two in 32..1024");
    if (p->K < 1 || p->crc_bits < 0 || p->crc_bits > 32 || p->K + p->crc_bits > p->N) return fail(PG_ERR_ARG, "bad K / crc_bits");
    if (p->crc_bits > 0 && (((p->crc_poly >> p->crc_bits) & 1ull) == 0 || (p->crc_poly & 1ull) == 0)) return fail(PG_ERR_ARG, "crc_poly must contain D^r and 1");
    if (p->decoder < PG_DEC_SC || p->decoder > PG_DEC_BP) return fail(PG_ERR_ARG, "bad decoder");
    if (p->real != PG_REAL_F64 && p->real != PG_REAL_F32 && p->real != PG_REAL_H2) return fail(PG_ERR_ARG, "bad real");
    if (p->real == PG_REAL_H2 && p->decoder != PG_DEC_BP) return fail(PG_ERR_UNSUPPORTED, "PG_REAL_H2 is a BP-only mode");
    if (p->nranks < 1 || p->rank < 0 || p->rank >= p->nranks) return fail(PG_ERR_ARG, "bad rank/nranks");
    if (!(p->llr_clip >= 0.0f)) return fail(PG_ERR_ARG, "llr_clip must be >= 0");
    const int L = (p->decoder == PG_DEC_SC) ? 1 : p->list_size;
    if (p->decoder != PG_DEC_BP && (L < 1 || L > 32 || (L & (L - 1)))) return fail(PG_ERR_ARG, "list_size must be 1,2,4,8,16,32");
    if ((p->decoder == PG_DEC_SCL || p->decoder == PG_DEC_CASCL) && L < 2) return fail(PG_ERR_ARG, "list decoders need list_size >= 2");
    if (p->decoder == PG_DEC_CASCL && p->crc_bits == 0) return fail(PG_ERR_ARG, "CA-SCL needs a CRC");
    if (p->decoder == PG_DEC_BP && (p->iter_max < 1 || p->iter_max > 255 || p->N < 64)) return fail(PG_ERR_ARG, "BP needs 1 <= iter_max <= 255 and N >= 64");

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return fail(PG_ERR_NO_DEVICE, std::string("no CUDA device: ") + cudaGetErrorString(e));
    if (p->device < 0 || p->device >= ndev) return fail(PG_ERR_ARG, "bad device ordinal");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, p->device)) != cudaSuccess) return fail(PG_ERR_CUDA, cudaGetErrorString(e));
    if (prop.major != 10) return fail(PG_ERR_NO_DEVICE, "kernels are built for sm_100a only; device is sm_" + std::to_string(prop.major * 10 + prop.minor));
    if ((e = cudaSetDevice(p->device)) != cudaSuccess) return fail(PG_ERR_CUDA, cudaGetErrorString(e));

    pg_ctx *ctx = new pg_ctx();
    ctx->p = *p;
    ctx->p.list_size = L;
    ctx->f64 = (p->real == PG_REAL_F64);
    ctx->h2 = (p->real == PG_REAL_H2);
    ctx->sm_count = prop.multiProcessorCount;
    build_code(ctx);
    auto bail = [&](int code) { g_create_error = ctx->err; pg_destroy(ctx); return code; };
#define CUC(call)                                                                                \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess) { ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); return bail(PG_ERR_CUDA); } \
    } while (0)
    CUC(cudaStreamCreateWithFlags(&ctx->st, cudaStreamNonBlocking));
    for (auto &ev : ctx->ev) CUC(cudaEventCreate(&ev));
    // tables
    {
        std::vector<uint16_t> I16(ctx->I.begin(), ctx->I.end());
        CUC(cudaMalloc(&ctx->d_I, sizeof(uint16_t) * std::max<size_t>(1, I16.size())));
        CUC(cudaMemcpy(ctx->d_I, I16.data(), sizeof(uint16_t) * I16.size(), cudaMemcpyHostToDevice));
        const int r = p->crc_bits;
        if (r > 0) {
            // rem[i] = D^i mod g(D) as an r-bit word
            std::vector<uint32_t> rem(ctx->nI);
            const uint64_t low = p->crc_poly & ((r == 64 ? 0 : (1ull << r)) - 1ull);
            uint64_t cur = 1;
            for (int i = 0; i < ctx->nI; i++) {
                rem[i] = (uint32_t)cur;
                cur <<= 1;
                if ((cur >> r) & 1ull) cur = (cur & ((1ull << r) - 1ull)) ^ low;
            }
            // syndrome bit b = parity over positions p=I[i] with bit b of rem[i] set  (CRcheck, CASCL_1024_L8.c:569-598)
            std::vector<uint32_t> masks((size_t)r * ctx->W, 0u);
            for (int i = 0; i < ctx->nI; i++)
                for (int b = 0; b < r; b++)
                    if ((rem[i] >> b) & 1u) masks[(size_t)b * ctx->W + (ctx->I[i] >> 5)] |= 1u << (ctx->I[i] & 31);
            CUC(cudaMalloc(&ctx->d_crc_masks, masks.size() * 4));
            CUC(cudaMemcpy(ctx->d_crc_masks, masks.data(), masks.size() * 4, cudaMemcpyHostToDevice));
            std::vector<uint32_t> sys(p->K);
            for (int i = 0; i < p->K; i++) sys[i] = rem[r + i];  // D^(r+i) mod g: row i of Gc (CASCL_1024_sys.c:49-561)
            CUC(cudaMalloc(&ctx->d_crc_sys, sys.size() * 4));
            CUC(cudaMemcpy(ctx->d_crc_sys, sys.data(), sys.size() * 4, cudaMemcpyHostToDevice));
        }
    }
    CUC(cudaMalloc(&ctx->d_counters, CNT_N * 8));
    CUC(cudaMemset(ctx->d_counters, 0, CNT_N * 8));
    CUC(cudaMalloc(&ctx->d_cnt2, CNT_N * 8));
    CUC(cudaMemset(ctx->d_cnt2, 0, CNT_N * 8));
    CUC(cudaMalloc(&ctx->d_queue, 8));
    CUC(cudaMalloc(&ctx->d_bpr, 8 * 16 * 8));
    CUC(cudaMemset(ctx->d_bpr, 0, 8 * 16 * 8));
    CUC(cudaMallocHost(&ctx->h_counters, CNT_N * 8));
    CUC(cudaMalloc(&ctx->d_xchg, (size_t)p->nranks * CNT_N * 8));
    // launch geometry
    if (p->decoder == PG_DEC_BP) {
        BpPlan bp;
        cudaError_t pe = ctx->h2 ? bp_h2_plan(ctx->n, &bp) : bp_plan(ctx->n, ctx->f64, &bp);
        if (pe != cudaSuccess) { ctx->err = std::string("no BP kernel for this N: ") + cudaGetErrorString(pe); return bail(PG_ERR_UNSUPPORTED); }
        if (bp.ctas_per_sm < 1) { ctx->err = "BP kernel does not fit on an SM"; return bail(PG_ERR_UNSUPPORTED); }
        ctx->grid = ctx->sm_count * bp.ctas_per_sm;
        ctx->frames_per_cta = ctx->h2 ? 2 : 1;
    } else {
        ListPlan lp;
        cudaError_t pe = list_plan(ctx->n, L, ctx->f64, &lp);
        if (pe != cudaSuccess) { ctx->err = std::string("no list kernel for this (N, L): ") + cudaGetErrorString(pe); return bail(PG_ERR_UNSUPPORTED); }
        if (lp.ctas_per_sm < 1) { ctx->err = "list kernel does not fit on an SM"; return bail(PG_ERR_UNSUPPORTED); }
        int per_sm = lp.ctas_per_sm;
        if (getenv("POLARGPU_DEBUG_PLAN")) fprintf(stderr, "polargpu: list kernel plan: %d CTAs/SM, %d frames/CTA, %zu B smem/CTA, %zu B scratch/CTA\n", lp.ctas_per_sm, lp.frames_per_cta, lp.smem, lp.scratch_per_cta);
        if (const char *cap = getenv("POLARGPU_LIST_CTAS_PER_SM")) per_sm = std::max(1, std::min(per_sm, atoi(cap)));  // tuning aid
        ctx->grid = ctx->sm_count * per_sm;
        ctx->scratch_per_cta = lp.scratch_per_cta;
        ctx->frames_per_cta = (size_t)lp.frames_per_cta;
        if (lp.scratch_per_cta) CUC(cudaMalloc(&ctx->d_scratch, lp.scratch_per_cta * (size_t)ctx->grid));
    }
    {
        // frames per launch: about 2^27 LLRs, rounded to whole waves of the decode kernel so that no SM idles at the end of a launch
        const char *env = getenv("POLARGPU_CHUNK");
        size_t c = env ? (size_t)atoll(env) : ((size_t)1 << 27) / (size_t)p->N;
        const size_t wave = (size_t)ctx->grid * ctx->frames_per_cta;
        if (!env && wave > 0) c = std::max<size_t>(1, c / wave) * wave;
        ctx->chunk_max = std::max<size_t>(c, 32);
    }
#undef CUC
    *out = ctx;
    return PG_OK;
}

extern "C" void pg_destroy(pg_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->p.device);
    if (ctx->st) cudaStreamSynchronize(ctx->st);
    if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm);
    free_buffers(ctx);
    for (int l = 1; l < kPipeLanes; l++)
        if (ctx->st_lane[l]) cudaStreamSynchronize(ctx->st_lane[l]);
    for (int s = 0; s < kPipeBufs; s++) {
        cudaFree(ctx->p_llr[s]); cudaFree(ctx->p_in[s]); cudaFree(ctx->p_uhat[s]); cudaFree(ctx->p_info[s]); cudaFree(ctx->p_bytes[s]);
        if (ctx->ev_h2d[s]) cudaEventDestroy(ctx->ev_h2d[s]);
        if (ctx->ev_free[s]) cudaEventDestroy(ctx->ev_free[s]);
    }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    for (int l = 0; l < kPipeLanes; l++)
        if (ctx->ev_join[l]) cudaEventDestroy(ctx->ev_join[l]);
    if (ctx->st_comm) cudaStreamSynchronize(ctx->st_comm);
    for (int r = 0; r < pg_ctx::kRing; r++) {
        cudaFree(ctx->d_rc[r]); cudaFree(ctx->d_xm[r]); cudaFree(ctx->d_rinfo[r]);
        if (ctx->h_xm[r]) cudaFreeHost(ctx->h_xm[r]);
        if (ctx->ev_round[r]) cudaEventDestroy(ctx->ev_round[r]);
        if (ctx->ev_xdone[r]) cudaEventDestroy(ctx->ev_xdone[r]);
    }
    if (ctx->st_comm) cudaStreamDestroy(ctx->st_comm);
    if (ctx->st_copy) cudaStreamDestroy(ctx->st_copy);
    for (int l = 1; l < kPipeLanes; l++)
        if (ctx->st_lane[l]) cudaStreamDestroy(ctx->st_lane[l]);
    cudaFree(ctx->d_I); cudaFree(ctx->d_crc_masks); cudaFree(ctx->d_crc_sys); cudaFree(ctx->d_counters); cudaFree(ctx->d_cnt2); cudaFree(ctx->d_queue);
    cudaFree(ctx->d_bpr); cudaFree(ctx->d_scratch); cudaFree(ctx->d_xchg);
    if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
    for (auto &ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    if (ctx->st) cudaStreamDestroy(ctx->st);
    delete ctx;
}

extern "C" const char *pg_last_error(const pg_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" int pg_info_set(const pg_ctx *ctx, int *I_out, uint8_t *inI_out)
{
    if (!ctx) return PG_ERR_ARG;
    if (I_out) std::copy(ctx->I.begin(), ctx->I.end(), I_out);
    if (inI_out) std::copy(ctx->inI.begin(), ctx->inI.end(), inI_out);
    return PG_OK;
}

static void add_counters(pg_counters *acc, const unsigned long long *c);

// ---------------------------------------------------------------- launches
// POLARGPU_DEBUG=1: synchronise after every kernel so that a fault is attributed to the launch that caused it
static int debug_sync(pg_ctx *ctx, const char *what)
{
    static const bool on = getenv("POLARGPU_DEBUG") != nullptr;
    if (!on) return PG_OK;
    cudaError_t e = cudaStreamSynchronize(ctx->st);
    if (e != cudaSuccess) { ctx->err = std::string(what) + " kernel: " + cudaGetErrorString(e); return PG_ERR_CUDA; }
    return PG_OK;
}

static int run_channel_to(pg_ctx *ctx, double ebn0_db, uint64_t first, size_t B, void *d_llr, uint32_t *d_truth)
{
    ChannelArgs a;
    std::memset(&a, 0, sizeof(a));
    a.llr = d_llr;
    a.u_packed = d_truth;
    a.I = ctx->d_I;
    a.crc_sys = ctx->d_crc_sys;
    a.first_frame = first; a.B = B;
    a.crc_poly = ctx->p.crc_poly; a.seed = ctx->p.seed;
    a.N = ctx->p.N; a.n = ctx->n; a.K = ctx->p.K; a.r = ctx->p.crc_bits; a.nI = ctx->nI;
    a.crc_systematic = ctx->p.crc_systematic; a.data_mode = ctx->p.data_mode;
    a.sigma_d = std::pow(10.0, ebn0_db / -20.0);  // SC_128.c:167
    a.sigma_f = (float)a.sigma_d;
    CU(cudaEventRecord(ctx->ev[0], ctx->st));
    CU(launch_channel(a, ctx->f64, ctx->sm_count, ctx->st));
    CU(cudaEventRecord(ctx->ev[1], ctx->st));
    ctx->ev_ch = true;
    ctx->launches++;
    return debug_sync(ctx, "channel");
}

static int run_channel(pg_ctx *ctx, double ebn0_db, uint64_t first, size_t B, bool want_llr)
{
    int rc = run_channel_to(ctx, ebn0_db, first, B, want_llr ? ctx->d_llr : nullptr, ctx->d_truth);
    if (!rc && want_llr && ctx->p.llr_clip > 0) {  // receiver model: clip what the channel delivered, in place
        CU(launch_convert_llr(ctx->d_llr, ctx->f64 ? PG_LLR_F64 : PG_LLR_F32, ctx->d_llr, ctx->f64, B * (size_t)ctx->p.N, ctx->st, ctx->p.llr_clip));
        ctx->launches++;
    }
    return rc;
}

// decode B frames on stream `st` with at most `grid_cap` CTAs that use the scratch slots [cta_off, cta_off + grid_cap)
static int run_decode_on(pg_ctx *ctx, cudaStream_t st, int grid_cap, int cta_off, const void *d_llr, size_t B, const uint32_t *d_truth,
                         uint32_t *d_uhat, uint32_t *d_info, unsigned long long *d_cnt)
{
    const pg_params &p = ctx->p;
    const bool timed = (st == ctx->st) && !ctx->ev_suppress;
    if (timed) CU(cudaEventRecord(ctx->ev[2], st));
    if (p.decoder == PG_DEC_BP) {
        BpArgs a;
        std::memset(&a, 0, sizeof(a));
        a.llr = d_llr; a.truth = d_truth; a.u_hat = d_uhat; a.frame_info = d_info;
        a.counters = d_cnt;
        a.queue = ctx->d_queue;
        a.bpr_E = ctx->bpr_ns ? ctx->d_bpr : nullptr;
        a.bpr_ns = ctx->bpr_ns;
        std::memcpy(a.bpr_samples, ctx->bpr_samples, sizeof(a.bpr_samples));
        a.B = B; a.iters = p.iter_max; a.early_stop = ctx->h2 ? (p.bp_early_stop & 1) : p.bp_early_stop;
        a.m = ctx->masks;
        CU(cudaMemsetAsync(ctx->d_queue, 0, 8, st));
        if (ctx->h2) {
            if (a.bpr_E) { ctx->err = "the BPR statistic is not available in PG_REAL_H2 mode"; return PG_ERR_UNSUPPORTED; }
            const int grid = (int)std::min<size_t>((size_t)grid_cap, (B + 1) / 2);
            CU(launch_bp_h2(a, ctx->n, std::max(grid, 1), st));
        } else {
            const int grid = (int)std::min<size_t>((size_t)grid_cap, B);
            CU(launch_bp(a, ctx->n, ctx->f64, std::max(grid, 1), st));
        }
    } else {
        ListArgs a;
        std::memset(&a, 0, sizeof(a));
        a.llr = d_llr; a.truth = d_truth; a.u_hat = d_uhat; a.frame_info = d_info;
        a.counters = d_cnt;
        a.gscratch = ctx->d_scratch ? (char *)ctx->d_scratch + (size_t)cta_off * ctx->scratch_per_cta : nullptr;
        a.crc_masks = ctx->d_crc_masks;
        a.B = B; a.r = p.crc_bits; a.use_crc = (p.decoder == PG_DEC_CASCL);
        {
            int first = p.N;
            for (int j = 0; j < p.N; j++) if (ctx->inI[j]) { first = j; break; }
            a.coop_groups = getenv("POLARGPU_NO_COOP") ? 0 : first / 4;
        }
        a.m = ctx->masks;
        const size_t ctas = (B + ctx->frames_per_cta - 1) / ctx->frames_per_cta;
        const int grid = (int)std::min<size_t>((size_t)grid_cap, ctas);
        CU(launch_list(a, ctx->n, p.list_size, ctx->f64, std::max(grid, 1), st));
    }
    if (timed) {
        CU(cudaEventRecord(ctx->ev[3], st));
        ctx->ev_dec = true;
    }
    ctx->launches++;
    return debug_sync(ctx, "decode");
}

// Device-resident decode of B frames, asynchronous on the ctx stream.  One full-grid launch -- except for the fp64 list decoders:
// their persistent warps slow down once they have drifted out of phase (twice the scratch bytes per frame; measured with
// tools/lane_probe.py: 4.03 M frames/s in one launch of 32 waves, 4.40 M in one-wave launches, 5.02 M in quarter-wave launches on
// four streams, where every warp decodes one frame group and leaves), so they run as short quarter-grid launches on four
// streams that fork from and join the ctx stream.  fp32 is indifferent to the launch shape (13.2 / 13.2 / 13.0 M).
static int run_decode_dev(pg_ctx *ctx, const void *d_llr, size_t B, const uint32_t *d_truth, uint32_t *d_uhat, uint32_t *d_info,
                          unsigned long long *d_cnt)
{
    const bool lanes_ok = ctx->p.decoder != PG_DEC_BP && ctx->f64 && ctx->grid >= 4 * kPipeLanes && !getenv("POLARGPU_NO_LANES");
    const size_t quarter = (size_t)(ctx->grid / kPipeLanes) * ctx->frames_per_cta;
    if (!lanes_ok || B <= quarter) return run_decode_on(ctx, ctx->st, ctx->grid, 0, d_llr, B, d_truth, d_uhat, d_info, d_cnt);
    int rc = ensure_lanes(ctx);
    if (rc) return rc;
    const size_t N = ctx->p.N, W = ctx->W, esz = 8;
    const int lane_grid = ctx->grid / kPipeLanes;
    CU(cudaEventRecord(ctx->ev[2], ctx->st));
    CU(cudaEventRecord(ctx->ev_fork, ctx->st));
    for (int l = 1; l < kPipeLanes; l++) CU(cudaStreamWaitEvent(ctx->st_lane[l], ctx->ev_fork, 0));
    size_t i = 0;
    ctx->ev_suppress = true;  // the events around the whole fork/join region time this decode, not one of its launches
    for (size_t off = 0; off < B && !rc; off += quarter, i++) {
        const size_t b = std::min(quarter, B - off);
        const int lane = (int)(i % kPipeLanes);
        rc = run_decode_on(ctx, ctx->st_lane[lane], lane_grid, lane * lane_grid, (const char *)d_llr + off * N * esz, b, d_truth ? d_truth + off * W : nullptr,
                           d_uhat ? d_uhat + off * W : nullptr, d_info ? d_info + off : nullptr, d_cnt);
    }
    ctx->ev_suppress = false;
    if (rc) return rc;
    for (int l = 1; l < kPipeLanes; l++) {
        CU(cudaEventRecord(ctx->ev_join[l], ctx->st_lane[l]));
        CU(cudaStreamWaitEvent(ctx->st, ctx->ev_join[l], 0));
    }
    CU(cudaEventRecord(ctx->ev[3], ctx->st));
    ctx->ev_dec = true;
    return PG_OK;
}

static int run_decode(pg_ctx *ctx, const void *d_llr, size_t B, const uint32_t *d_truth, uint32_t *d_uhat, uint32_t *d_info, bool count)
{
    return run_decode_dev(ctx, d_llr, B, d_truth, d_uhat, d_info, count ? ctx->d_counters : nullptr);
}

extern "C" int pg_decode_llr_device(pg_ctx *ctx, const void *d_llr, int llr_is_f64, size_t B, uint32_t *d_u_hat_packed, uint32_t *d_frame_info)
{
    if (!ctx || !d_llr || !llr_fmt_ok(llr_is_f64)) return PG_ERR_ARG;
    if (B == 0) return PG_OK;
    CU(cudaSetDevice(ctx->p.device));
    const void *src = d_llr;
    if (llr_is_f64 != (ctx->f64 ? PG_LLR_F64 : PG_LLR_F32) || ctx->p.llr_clip > 0) {
        int rc = ensure_capacity(ctx, B);
        if (rc) return rc;
        CU(launch_convert_llr(d_llr, llr_is_f64, ctx->d_llr, ctx->f64, B * (size_t)ctx->p.N, ctx->st, ctx->p.llr_clip));
        ctx->launches++;
        src = ctx->d_llr;
    }
    return run_decode(ctx, src, B, nullptr, d_u_hat_packed, d_frame_info, false);
}

extern "C" int pg_decode_count_device(pg_ctx *ctx, const void *d_llr, int llr_is_f64, size_t B, const uint32_t *d_truth_packed,
                                      uint32_t *d_u_hat_packed, uint32_t *d_frame_info)
{
    if (!ctx || !d_llr) return PG_ERR_ARG;
    if (B == 0) return PG_OK;
    CU(cudaSetDevice(ctx->p.device));
    if (llr_is_f64 != (ctx->f64 ? PG_LLR_F64 : PG_LLR_F32)) { ctx->err = "pg_decode_count_device: LLR type must be the context's arithmetic type"; return PG_ERR_ARG; }
    if (ctx->p.llr_clip > 0) {  // the caller's buffer stays untouched: clipped copy in the context's work buffer
        int rc = ensure_capacity(ctx, B);
        if (rc) return rc;
        CU(launch_convert_llr(d_llr, llr_is_f64, ctx->d_llr, ctx->f64, B * (size_t)ctx->p.N, ctx->st, ctx->p.llr_clip));
        ctx->launches++;
        d_llr = ctx->d_llr;
    }
    return run_decode(ctx, d_llr, B, d_truth_packed, d_u_hat_packed, d_frame_info, true);
}

extern "C" int pg_counters_read(pg_ctx *ctx, pg_counters *out, int reset)
{
    if (!ctx || !out) return PG_ERR_ARG;
    CU(cudaSetDevice(ctx->p.device));
    CU(cudaMemcpyAsync(ctx->h_counters, ctx->d_counters, CNT_N * 8, cudaMemcpyDeviceToHost, ctx->st));
    if (reset) CU(cudaMemsetAsync(ctx->d_counters, 0, CNT_N * 8, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    std::memcpy(out, ctx->h_counters, CNT_N * 8);
    return PG_OK;
}

extern "C" int pg_channel_device(pg_ctx *ctx, double ebn0_db, uint64_t first_frame, size_t B, void *d_llr, uint32_t *d_u_packed)
{
    if (!ctx) return PG_ERR_ARG;
    if (B == 0) return PG_OK;
    CU(cudaSetDevice(ctx->p.device));
    return run_channel_to(ctx, ebn0_db, first_frame, B, d_llr, d_u_packed);
}

static uint64_t wave_frames(const pg_ctx *ctx)
{
    return (uint64_t)ctx->grid * ctx->frames_per_cta;
}

// Host-pointer decode.  Batches larger than one wave of the decode kernel are cut into wave-sized chunks and pipelined over
// two buffer sets: the H2D copy of chunk i+1 (copy stream) overlaps the decode + D2H of chunk i (compute stream).
static int decode_host(pg_ctx *ctx, const void *llr, int llr_is_f64, size_t B, uint8_t *u_hat, uint32_t *u_hat_packed, uint32_t *flags)
{
    if (!ctx || !llr || !llr_fmt_ok(llr_is_f64)) return PG_ERR_ARG;
    CU(cudaSetDevice(ctx->p.device));
    const size_t N = ctx->p.N, W = ctx->W;
    const size_t esz = llr_esz(llr_is_f64);
    const bool conv = llr_is_f64 != (ctx->f64 ? PG_LLR_F64 : PG_LLR_F32) || ctx->p.llr_clip > 0;
    // Chunk = what one launch keeps resident.  A wave-sized launch runs all its warps in lockstep through the same phases of the
    // schedule (memory-heavy top layers, then leaf-heavy stretches), which costs the list kernel ~15 %; four quarter-grid launches on
    // four streams, offset by a quarter chunk each, interleave those phases and quarter the pipeline fill
    // (measured, 8 waves of CA-SCL 1024: 10.0 / 11.35 / 11.8 Mframes/s with 1 / 2 / 4 lanes).
    int lanes = (ctx->p.decoder != PG_DEC_BP) ? kPipeLanes : 1;
    if (const char *e = getenv("POLARGPU_LANES")) lanes = std::max(1, std::min(kPipeLanes, atoi(e)));  // tuning aid
    if (ctx->p.decoder == PG_DEC_BP || ctx->grid < lanes) lanes = 1;
    const int nbuf = 2 * lanes, lane_grid = ctx->grid / lanes;
    // BP frames come from a device-side queue (no tail imbalance), and a wave is only a few hundred frames: several waves per chunk
    size_t bp_waves = 8;
    if (const char *e = getenv("POLARGPU_BP_CHUNK_WAVES")) bp_waves = (size_t)std::max(1, atoi(e));  // tuning aid
    const size_t per_launch = (ctx->p.decoder == PG_DEC_BP) ? (size_t)wave_frames(ctx) * bp_waves : (size_t)wave_frames(ctx) / lanes;
    const size_t pc = std::min<size_t>(ctx->chunk_max, std::max<size_t>(per_launch, 1024));
    if (B > pc && !getenv("POLARGPU_NO_PIPELINE")) {
        int rc = ensure_pipe(ctx, pc, u_hat != nullptr);
        if (rc) return rc;
        auto enqueue_all = [&]() -> int {
            size_t i = 0;
            for (size_t off = 0; off < B; off += pc, i++) {
                const size_t b = std::min(pc, B - off);
                const int s = (int)(i % nbuf), lane = (int)(i % lanes);
                cudaStream_t cs = ctx->st_lane[lane];
                if (i >= (size_t)nbuf) CU(cudaStreamWaitEvent(ctx->st_copy, ctx->ev_free[s], 0));
                void *dst = conv ? ctx->p_in[s] : ctx->p_llr[s];
                CU(cudaMemcpyAsync(dst, (const char *)llr + off * N * esz, b * N * esz, cudaMemcpyHostToDevice, ctx->st_copy));
                CU(cudaEventRecord(ctx->ev_h2d[s], ctx->st_copy));
                CU(cudaStreamWaitEvent(cs, ctx->ev_h2d[s], 0));
                if (conv) { CU(launch_convert_llr(ctx->p_in[s], llr_is_f64, ctx->p_llr[s], ctx->f64, b * N, cs, ctx->p.llr_clip)); ctx->launches++; }
                int rc2 = run_decode_on(ctx, cs, lane_grid, lane * lane_grid, ctx->p_llr[s], b, nullptr, ctx->p_uhat[s], ctx->p_info[s], nullptr);
                if (rc2) return rc2;
                if (u_hat) {
                    CU(launch_unpack_bits(ctx->p_uhat[s], ctx->p_bytes[s], b, (int)N, cs));
                    ctx->launches++;
                    CU(cudaMemcpyAsync(u_hat + off * N, ctx->p_bytes[s], b * N, cudaMemcpyDeviceToHost, cs));
                }
                if (u_hat_packed) CU(cudaMemcpyAsync(u_hat_packed + off * W, ctx->p_uhat[s], b * W * 4, cudaMemcpyDeviceToHost, cs));
                if (flags) CU(cudaMemcpyAsync(flags + off, ctx->p_info[s], b * 4, cudaMemcpyDeviceToHost, cs));
                CU(cudaEventRecord(ctx->ev_free[s], cs));
            }
            return PG_OK;
        };
        rc = enqueue_all();
        if (rc) {  // copies and kernels of earlier chunks are still in flight and write into the caller's buffers: wait for them
            cudaStreamSynchronize(ctx->st_copy);
            for (int l = 0; l < lanes; l++) cudaStreamSynchronize(ctx->st_lane[l]);
            return rc;
        }
        for (int l = 0; l < lanes; l++) CU(cudaStreamSynchronize(ctx->st_lane[l]));
    } else {
        for (size_t off = 0; off < B; off += ctx->chunk_max) {
            const size_t b = std::min(ctx->chunk_max, B - off);
            int rc = ensure_capacity(ctx, b);
            if (rc) return rc;
            void *dst = conv ? ctx->d_in : ctx->d_llr;
            CU(cudaMemcpyAsync(dst, (const char *)llr + off * N * esz, b * N * esz, cudaMemcpyHostToDevice, ctx->st));
            if (conv) { CU(launch_convert_llr(ctx->d_in, llr_is_f64, ctx->d_llr, ctx->f64, b * N, ctx->st, ctx->p.llr_clip)); ctx->launches++; }
            rc = run_decode(ctx, ctx->d_llr, b, nullptr, ctx->d_uhat, ctx->d_info, false);
            if (rc) return rc;
            if (u_hat) {
                CU(launch_unpack_bits(ctx->d_uhat, ctx->d_bytes, b, (int)N, ctx->st));
                ctx->launches++;
                CU(cudaMemcpyAsync(u_hat + off * N, ctx->d_bytes, b * N, cudaMemcpyDeviceToHost, ctx->st));
            }
            if (u_hat_packed) CU(cudaMemcpyAsync(u_hat_packed + off * W, ctx->d_uhat, b * W * 4, cudaMemcpyDeviceToHost, ctx->st));
            if (flags) CU(cudaMemcpyAsync(flags + off, ctx->d_info, b * 4, cudaMemcpyDeviceToHost, ctx->st));
            CU(cudaStreamSynchronize(ctx->st));
        }
    }
    if (flags)
        for (size_t i = 0; i < B; i++) flags[i] = PG_INFO_TO_FLAGS(flags[i]);
    return PG_OK;
}

extern "C" int pg_decode_llr(pg_ctx *ctx, const void *llr, int llr_is_f64, size_t B, uint8_t *u_hat, uint32_t *flags)
{
    return decode_host(ctx, llr, llr_is_f64, B, u_hat, nullptr, flags);
}

extern "C" int pg_decode_llr_packed(pg_ctx *ctx, const void *llr, int llr_is_f64, size_t B, uint32_t *u_hat_packed, uint32_t *flags)
{
    return decode_host(ctx, llr, llr_is_f64, B, nullptr, u_hat_packed, flags);
}

extern "C" int pg_decode_llr_counted(pg_ctx *ctx, const void *llr, int llr_is_f64, size_t B, const uint8_t *u_true, uint8_t *u_hat,
                                     pg_counters *acc, uint16_t *frame_err)
{
    if (!ctx || !llr || !u_true || !acc || !llr_fmt_ok(llr_is_f64)) return PG_ERR_ARG;
    CU(cudaSetDevice(ctx->p.device));
    const size_t N = ctx->p.N, W = ctx->W;
    const size_t esz = llr_esz(llr_is_f64);
    const bool conv = llr_is_f64 != (ctx->f64 ? PG_LLR_F64 : PG_LLR_F32) || ctx->p.llr_clip > 0;
    std::vector<uint32_t> packed;
    for (size_t off = 0; off < B; off += ctx->chunk_max) {
        const size_t b = std::min(ctx->chunk_max, B - off);
        int rc = ensure_capacity(ctx, b);
        if (rc) return rc;
        packed.assign(b * W, 0u);
        for (size_t f = 0; f < b; f++)
            for (size_t j = 0; j < N; j++)
                if (u_true[(off + f) * N + j]) packed[f * W + (j >> 5)] |= 1u << (j & 31);
        CU(cudaMemcpyAsync(ctx->d_truth, packed.data(), b * W * 4, cudaMemcpyHostToDevice, ctx->st));
        void *dst = conv ? ctx->d_in : ctx->d_llr;
        CU(cudaMemcpyAsync(dst, (const char *)llr + off * N * esz, b * N * esz, cudaMemcpyHostToDevice, ctx->st));
        if (conv) { CU(launch_convert_llr(ctx->d_in, llr_is_f64, ctx->d_llr, ctx->f64, b * N, ctx->st, ctx->p.llr_clip)); ctx->launches++; }
        CU(cudaMemsetAsync(ctx->d_cnt2, 0, CNT_N * 8, ctx->st));
        rc = run_decode_dev(ctx, ctx->d_llr, b, ctx->d_truth, ctx->d_uhat, ctx->d_info, ctx->d_cnt2);
        if (rc) return rc;
        if (u_hat) {
            CU(launch_unpack_bits(ctx->d_uhat, ctx->d_bytes, b, (int)N, ctx->st));
            ctx->launches++;
            CU(cudaMemcpyAsync(u_hat + off * N, ctx->d_bytes, b * N, cudaMemcpyDeviceToHost, ctx->st));
        }
        CU(cudaMemcpyAsync(ctx->h_counters, ctx->d_cnt2, CNT_N * 8, cudaMemcpyDeviceToHost, ctx->st));
        CU(cudaMemcpyAsync(ctx->h_info, ctx->d_info, b * 4, cudaMemcpyDeviceToHost, ctx->st));
        CU(cudaStreamSynchronize(ctx->st));
        add_counters(acc, ctx->h_counters);
        if (frame_err)
            for (size_t i = 0; i < b; i++) frame_err[off + i] = (uint16_t)(ctx->h_info[i] & 0xFFFFu);
    }
    return PG_OK;
}

extern "C" int pg_channel(pg_ctx *ctx, double ebn0_db, uint64_t first_frame, size_t B, void *llr_out, uint8_t *u_out)
{
    if (!ctx) return PG_ERR_ARG;
    CU(cudaSetDevice(ctx->p.device));
    const size_t N = ctx->p.N, esz = ctx->f64 ? 8 : 4;
    for (size_t off = 0; off < B; off += ctx->chunk_max) {
        const size_t b = std::min(ctx->chunk_max, B - off);
        int rc = ensure_capacity(ctx, b);
        if (rc) return rc;
        rc = run_channel(ctx, ebn0_db, first_frame + off, b, true);
        if (rc) return rc;
        if (llr_out) CU(cudaMemcpyAsync((char *)llr_out + off * N * esz, ctx->d_llr, b * N * esz, cudaMemcpyDeviceToHost, ctx->st));
        if (u_out) {
            CU(launch_unpack_bits(ctx->d_truth, ctx->d_bytes, b, (int)N, ctx->st));
            ctx->launches++;
            CU(cudaMemcpyAsync(u_out + off * N, ctx->d_bytes, b * N, cudaMemcpyDeviceToHost, ctx->st));
        }
        CU(cudaStreamSynchronize(ctx->st));
    }
    return PG_OK;
}

// one chunk: channel + decode + count; leaves the chunk's counters (delta) in h_counters and frame_info in h_info (if wanted)
static int simulate_chunk(pg_ctx *ctx, double ebn0_db, uint64_t first, size_t b, bool want_info)
{
    int rc = ensure_capacity(ctx, b);
    if (rc) return rc;
    CU(cudaMemsetAsync(ctx->d_cnt2, 0, CNT_N * 8, ctx->st));
    rc = run_channel(ctx, ebn0_db, first, b, true);
    if (rc) return rc;
    rc = run_decode_dev(ctx, ctx->d_llr, b, ctx->d_truth, nullptr, ctx->d_info, ctx->d_cnt2);
    if (rc) return rc;
    CU(cudaMemcpyAsync(ctx->h_counters, ctx->d_cnt2, CNT_N * 8, cudaMemcpyDeviceToHost, ctx->st));
    if (want_info) CU(cudaMemcpyAsync(ctx->h_info, ctx->d_info, b * 4, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    return PG_OK;
}

static void add_counters(pg_counters *acc, const unsigned long long *c)
{
    acc->frames += c[CNT_FRAMES]; acc->err_blocks += c[CNT_ERR_BLOCKS]; acc->err_bits += c[CNT_ERR_BITS];
    acc->tie_frames += c[CNT_TIE]; acc->crc_fail += c[CNT_CRCFAIL]; acc->bp_sweeps += c[CNT_SWEEPS];
}

extern "C" int pg_simulate_batch(pg_ctx *ctx, double ebn0_db, uint64_t first_frame, size_t B, pg_counters *acc, uint16_t *frame_err)
{
    if (!ctx || !acc) return PG_ERR_ARG;
    CU(cudaSetDevice(ctx->p.device));
    for (size_t off = 0; off < B; off += ctx->chunk_max) {
        const size_t b = std::min(ctx->chunk_max, B - off);
        int rc = simulate_chunk(ctx, ebn0_db, first_frame + off, b, frame_err != nullptr);
        if (rc) return rc;
        add_counters(acc, ctx->h_counters);
        if (frame_err)
            for (size_t i = 0; i < b; i++) frame_err[off + i] = (uint16_t)(ctx->h_info[i] & 0xFFFFu);
    }
    return PG_OK;
}

// sum-exchange of a small u64 vector over the ranks (identity for nranks == 1)
static int exchange(pg_ctx *ctx, unsigned long long *host_vec, size_t count)
{
    if (ctx->p.nranks == 1) return PG_OK;
    if (!ctx->comm) { ctx->err = "nranks > 1 but pg_comm_init was not called"; return PG_ERR_NCCL; }
    CU(cudaMemcpyAsync(ctx->d_xchg, host_vec, count * 8, cudaMemcpyHostToDevice, ctx->st));
    ncclResult_t r = g_nccl.AllReduce(ctx->d_xchg, ctx->d_xchg, count, ncclUint64, ncclSum, ctx->comm, ctx->st);
    if (r != ncclSuccess) { ctx->err = std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error"); return PG_ERR_NCCL; }
    CU(cudaMemcpyAsync(host_vec, ctx->d_xchg, count * 8, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    return PG_OK;
}

// ring of round slots for pg_simulate: per-round counters, exchange matrix (device + pinned host), events; per-frame words on demand
static int ensure_ring(pg_ctx *ctx, size_t frames, bool want_info)
{
    const size_t R = (size_t)ctx->p.nranks;
    if (!ctx->st_comm) {
        {   // the exchange stream outranks the decode streams: its tiny kernels take the first SM slot that frees up
            int lo = 0, hi = 0;
            CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            CU(cudaStreamCreateWithPriority(&ctx->st_comm, cudaStreamNonBlocking, hi));
        }
        for (int r = 0; r < pg_ctx::kRing; r++) {
            CU(cudaMalloc(&ctx->d_rc[r], CNT_N * 8));
            CU(cudaMalloc(&ctx->d_xm[r], R * CNT_N * 8));
            CU(cudaMallocHost(&ctx->h_xm[r], R * CNT_N * 8));
            CU(cudaEventCreateWithFlags(&ctx->ev_round[r], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&ctx->ev_xdone[r], cudaEventDisableTiming));
        }
    }
    if (want_info && frames > ctx->ring_cap) {
        ctx->ring_cap = 0;
        for (int r = 0; r < pg_ctx::kRing; r++) {
            cudaFree(ctx->d_rinfo[r]);
            ctx->d_rinfo[r] = nullptr;
            CU(cudaMalloc(&ctx->d_rinfo[r], frames * 4));
        }
        ctx->ring_cap = frames;
    }
    return PG_OK;
}

// Enqueue one round into ring slot `slot`: channel + decode + count on the compute stream, then -- on the exchange stream, so that the
// next round's kernels start right behind this round's -- the round's [nranks][CNT_N] matrix (own row = own counters), its NCCL
// all-reduce when there is more than one rank, and the copy of the result into pinned host memory.  Nothing here waits for the GPU.
static int enqueue_round(pg_ctx *ctx, int slot, double ebn0_db, uint64_t start, size_t mine, bool want_info)
{
    const int R = ctx->p.nranks, me = ctx->p.rank;
    CU(cudaStreamWaitEvent(ctx->st, ctx->ev_xdone[slot], 0));  // the slot's previous round has left the device buffers
    CU(cudaMemsetAsync(ctx->d_rc[slot], 0, CNT_N * 8, ctx->st));
    if (mine) {
        int rc = run_channel(ctx, ebn0_db, start, mine, true);
        if (rc) return rc;
        rc = run_decode_dev(ctx, ctx->d_llr, mine, ctx->d_truth, nullptr, want_info ? ctx->d_rinfo[slot] : nullptr, ctx->d_rc[slot]);
        if (rc) return rc;
    }
    CU(cudaEventRecord(ctx->ev_round[slot], ctx->st));
    CU(cudaStreamWaitEvent(ctx->st_comm, ctx->ev_round[slot], 0));
    if (R > 1) {
        CU(cudaMemsetAsync(ctx->d_xm[slot], 0, (size_t)R * CNT_N * 8, ctx->st_comm));
        CU(cudaMemcpyAsync(ctx->d_xm[slot] + (size_t)me * CNT_N, ctx->d_rc[slot], CNT_N * 8, cudaMemcpyDeviceToDevice, ctx->st_comm));
        ncclResult_t r = g_nccl.AllReduce(ctx->d_xm[slot], ctx->d_xm[slot], (size_t)R * CNT_N, ncclUint64, ncclSum, ctx->comm, ctx->st_comm);
        if (r != ncclSuccess) { ctx->err = std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error"); return PG_ERR_NCCL; }
        ctx->sim_allreduces++;
        CU(cudaMemcpyAsync(ctx->h_xm[slot], ctx->d_xm[slot], (size_t)R * CNT_N * 8, cudaMemcpyDeviceToHost, ctx->st_comm));
    } else {
        CU(cudaMemcpyAsync(ctx->h_xm[slot], ctx->d_rc[slot], CNT_N * 8, cudaMemcpyDeviceToHost, ctx->st_comm));
    }
    CU(cudaEventRecord(ctx->ev_xdone[slot], ctx->st_comm));
    ctx->sim_rounds++;
    return PG_OK;
}

// The Monte-Carlo loop of the reference's main() (SC_128.c:164-222), pipelined: round i+1 is enqueued before the host looks at round
// i, so the GPU never waits for the host or for the collective.  A run that overshoots its stopping condition by the one round in
// flight discards that round: the result is defined by global frame order (SC_128.c:169), not by what was computed.
// With a frame budget only (target_err_blocks == 0) the number of rounds is known in advance: the counters accumulate on the
// device over all rounds and ONE all-reduce combines the ranks at the end.
extern "C" int pg_simulate(pg_ctx *ctx, double ebn0_db, uint64_t first_frame, uint64_t target_err_blocks, uint64_t max_frames,
                           int exact_stop, pg_counters *out)
{
    if (!ctx || !out || (target_err_blocks == 0 && max_frames == 0)) return PG_ERR_ARG;
    CU(cudaSetDevice(ctx->p.device));
    const int R = ctx->p.nranks, me = ctx->p.rank;
    static_assert(sizeof(pg_counters) == CNT_N * 8, "pg_counters layout");
    std::memset(out, 0, sizeof(*out));
    if (R > 1 && !ctx->comm) { ctx->err = "nranks > 1 but pg_comm_init was not called"; return PG_ERR_NCCL; }
    ctx->sim_rounds = ctx->sim_allreduces = 0;
    uint64_t next = first_frame;  // first global frame of the next round

    if (target_err_blocks == 0) {
        // ---- frame budget only: fixed schedule, device-side accumulation, one exchange
        int rc = ensure_capacity(ctx, std::min<uint64_t>(ctx->chunk_max, (max_frames + R - 1) / R));
        if (rc) return rc;
        CU(cudaMemsetAsync(ctx->d_cnt2, 0, CNT_N * 8, ctx->st));
        uint64_t left = max_frames;
        unsigned queued = 0;
        while (left) {
            const size_t chunk = (size_t)std::min<uint64_t>(ctx->chunk_max, (left + R - 1) / R);  // the last round splits what is left
            uint64_t start = 0, mine = 0;
            pg_partition(next, chunk, R, me, left, &start, &mine);
            if (mine) {
                rc = run_channel(ctx, ebn0_db, start, (size_t)mine, true);
                if (rc) return rc;
                rc = run_decode_dev(ctx, ctx->d_llr, (size_t)mine, ctx->d_truth, nullptr, nullptr, ctx->d_cnt2);
                if (rc) return rc;
            }
            ctx->sim_rounds++;
            const uint64_t round = std::min<uint64_t>(left, (uint64_t)R * chunk);
            left -= round;
            next += round;
            if ((++queued & 63u) == 0) CU(cudaStreamSynchronize(ctx->st));  // bound the launch queue
        }
        CU(cudaMemcpyAsync(ctx->h_counters, ctx->d_cnt2, CNT_N * 8, cudaMemcpyDeviceToHost, ctx->st));
        CU(cudaStreamSynchronize(ctx->st));
        std::memcpy(out, ctx->h_counters, CNT_N * 8);
        if (R > 1) ctx->sim_allreduces++;
        return exchange(ctx, reinterpret_cast<unsigned long long *>(out), CNT_N);
    }

    // ---- error target (with or without a frame budget): rounds of nranks x chunk frames, chunk doubling from 4096
    struct Round { uint64_t first; size_t chunk; uint64_t mine; uint64_t round_frames; };
    Round ring[pg_ctx::kRing];
    size_t chunk = std::min<size_t>(ctx->chunk_max, 4096);
    {
        int rc = ensure_capacity(ctx, ctx->chunk_max);
        if (!rc) rc = ensure_ring(ctx, ctx->chunk_max, exact_stop != 0);
        if (rc) return rc;
    }
    uint64_t planned = 0;          // frames of all rounds enqueued so far
    unsigned head = 0, tail = 0;   // rounds enqueued / rounds merged
    auto enqueue_next = [&]() -> int {
        const uint64_t budget = max_frames ? (max_frames - planned) : ~0ull;
        Round rd;
        rd.first = next; rd.chunk = chunk;
        uint64_t start = 0;
        pg_partition(next, chunk, R, me, budget, &start, &rd.mine);
        rd.round_frames = std::min<uint64_t>(budget, (uint64_t)R * chunk);
        const int slot = (int)(head % pg_ctx::kRing);
        int rc = enqueue_round(ctx, slot, ebn0_db, start, (size_t)rd.mine, exact_stop != 0);
        if (rc) return rc;
        ring[slot] = rd;
        head++;
        planned += rd.round_frames;
        next += (uint64_t)R * chunk;
        if (chunk < ctx->chunk_max) chunk = std::min(ctx->chunk_max, chunk * 2);
        return PG_OK;
    };
    auto drain = [&]() {  // rounds in flight past the stopping point: let them finish (every rank enqueued the same ones)
        cudaStreamSynchronize(ctx->st);
        cudaStreamSynchronize(ctx->st_comm);
    };
    // Rounds in flight.  The decode kernels are persistent, so the exchange of round i finds a free SM only while round i+1 drains:
    // the host learns about round i at the end of round i+1, and round i+2 must already be queued for the GPU not to idle -- three
    // in flight.  (The ranks may then also drift apart by a round or two instead of meeting at every all-reduce.)
    // The BPR statistic sums on the device over everything that was decoded: no round may run past the stopping point then.
    const unsigned depth = ctx->bpr_ns ? 1u : (unsigned)(pg_ctx::kRing - 1);
    int rc = enqueue_next();
    if (rc) return rc;
    std::vector<pg_counters> xc((size_t)R);
    while (true) {
        // keep `depth` rounds in flight while the host waits for the oldest one
        while (head - tail < depth && (!max_frames || planned < max_frames)) {
            rc = enqueue_next();
            if (rc) { drain(); return rc; }
        }
        const int slot = (int)(tail % pg_ctx::kRing);
        CU(cudaEventSynchronize(ctx->ev_xdone[slot]));
        std::memset(xc.data(), 0, sizeof(pg_counters) * xc.size());
        if (R > 1) std::memcpy(xc.data(), ctx->h_xm[slot], (size_t)R * CNT_N * 8);
        else std::memcpy(&xc[0], ctx->h_xm[slot], CNT_N * 8);
        int cut = -1;
        uint64_t need = 0;
        pg_merge_round(xc.data(), R, target_err_blocks, exact_stop, out, &cut, &need);
        bool done = false;
        if (cut >= 0) {
            // the target-th block error fell into rank `cut`'s chunk: that rank truncates, everybody learns the result
            pg_counters part;
            std::memset(&part, 0, sizeof(part));
            drain();
            if (cut == me) {
                rc = ensure_capacity(ctx, (size_t)ring[slot].mine);
                if (rc) return rc;
                CU(cudaMemcpy(ctx->h_info, ctx->d_rinfo[slot], (size_t)ring[slot].mine * 4, cudaMemcpyDeviceToHost));
                pg_truncate_info(ctx->h_info, (size_t)ring[slot].mine, need, &part);
            }
            rc = exchange(ctx, reinterpret_cast<unsigned long long *>(&part), CNT_N);
            if (rc) return rc;
            if (R > 1) ctx->sim_allreduces++;
            add_counters(out, reinterpret_cast<unsigned long long *>(&part));
            return PG_OK;
        }
        tail++;
        if (out->err_blocks >= target_err_blocks) done = true;
        if (max_frames && out->frames >= max_frames) done = true;
        if (done) { drain(); return PG_OK; }
        if (tail == head) {  // nothing in flight and not done (frame budget exhausted exactly would have ended above)
            rc = enqueue_next();
            if (rc) return rc;
        }
    }
}

extern "C" int pg_simulate_stats(const pg_ctx *ctx, uint64_t *rounds, uint64_t *allreduces)
{
    if (!ctx) return PG_ERR_ARG;
    if (rounds) *rounds = ctx->sim_rounds;
    if (allreduces) *allreduces = ctx->sim_allreduces;
    return PG_OK;
}

// ---------------------------------------------------------------- BPR statistic
extern "C" int pg_bpr_config(pg_ctx *ctx, const int *sample_sweeps, int ns)
{
    if (!ctx || ns < 0 || ns > 8 || (ns && !sample_sweeps)) return PG_ERR_ARG;
    ctx->bpr_ns = ns;
    for (int i = 0; i < ns; i++) ctx->bpr_samples[i] = sample_sweeps[i];
    return pg_bpr_reset(ctx);
}

extern "C" int pg_bpr_read(pg_ctx *ctx, uint64_t *E)
{
    if (!ctx || !E) return PG_ERR_ARG;
    CU(cudaSetDevice(ctx->p.device));
    CU(cudaStreamSynchronize(ctx->st));
    std::vector<unsigned long long> tmp(8 * 16);
    CU(cudaMemcpy(tmp.data(), ctx->d_bpr, tmp.size() * 8, cudaMemcpyDeviceToHost));
    for (int q = 0; q < ctx->bpr_ns; q++)
        for (int s = 0; s <= ctx->n; s++) E[q * (ctx->n + 1) + s] = tmp[q * 16 + s];
    return PG_OK;
}

extern "C" int pg_bpr_reset(pg_ctx *ctx)
{
    if (!ctx) return PG_ERR_ARG;
    CU(cudaSetDevice(ctx->p.device));
    CU(cudaMemsetAsync(ctx->d_bpr, 0, 8 * 16 * 8, ctx->st));
    return PG_OK;
}

// ---------------------------------------------------------------- multi-GPU
extern "C" int pg_comm_unique_id(void *id128)
{
    if (!id128) return PG_ERR_ARG;
    std::string err;
    if (!g_nccl.load(err)) { g_create_error = err; return PG_ERR_NCCL; }
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) { g_create_error = "ncclGetUniqueId failed"; return PG_ERR_NCCL; }
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    std::memcpy(id128, &id, 128);
    return PG_OK;
}

extern "C" int pg_comm_init(pg_ctx *ctx, const void *id128)
{
    if (!ctx || !id128) return PG_ERR_ARG;
    if (!g_nccl.load(ctx->err)) return PG_ERR_NCCL;
    CU(cudaSetDevice(ctx->p.device));
    ncclUniqueId id;
    std::memcpy(&id, id128, 128);
    ncclResult_t r = g_nccl.CommInitRank(&ctx->comm, ctx->p.nranks, id, ctx->p.rank);
    if (r != ncclSuccess) { ctx->err = std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error"); ctx->comm = nullptr; return PG_ERR_NCCL; }
    // NCCL connects its channels lazily, inside the first collective (about a second with eight ranks): do that here, not in the
    // first round of a Monte-Carlo loop whose pipeline would sit idle behind it
    CU(cudaMemsetAsync(ctx->d_xchg, 0, CNT_N * 8, ctx->st));
    r = g_nccl.AllReduce(ctx->d_xchg, ctx->d_xchg, CNT_N, ncclUint64, ncclSum, ctx->comm, ctx->st);
    if (r != ncclSuccess) { ctx->err = std::string("ncclAllReduce (warm-up): ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error"); return PG_ERR_NCCL; }
    CU(cudaStreamSynchronize(ctx->st));
    return PG_OK;
}

extern "C" int pg_allreduce_counters(pg_ctx *ctx, pg_counters *c)
{
    if (!ctx || !c) return PG_ERR_ARG;
    static_assert(sizeof(pg_counters) == CNT_N * 8, "pg_counters layout");
    CU(cudaSetDevice(ctx->p.device));
    return exchange(ctx, reinterpret_cast<unsigned long long *>(c), CNT_N);
}

// ---------------------------------------------------------------- introspection
extern "C" int pg_sync(pg_ctx *ctx)
{
    if (!ctx) return PG_ERR_ARG;
    CU(cudaSetDevice(ctx->p.device));
    CU(cudaStreamSynchronize(ctx->st));
    return PG_OK;
}

extern "C" uint64_t pg_wave_frames(const pg_ctx *ctx) { return ctx ? wave_frames(ctx) : 0; }

extern "C" void *pg_stream(pg_ctx *ctx) { return ctx ? (void *)ctx->st : nullptr; }
extern "C" uint64_t pg_kernel_launches(const pg_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int pg_last_kernel_ms(pg_ctx *ctx, float *decode_ms, float *channel_ms)
{
    if (!ctx) return PG_ERR_ARG;
    CU(cudaSetDevice(ctx->p.device));
    CU(cudaStreamSynchronize(ctx->st));
    if (decode_ms) { *decode_ms = 0; if (ctx->ev_dec) CU(cudaEventElapsedTime(decode_ms, ctx->ev[2], ctx->ev[3])); }
    if (channel_ms) { *channel_ms = 0; if (ctx->ev_ch) CU(cudaEventElapsedTime(channel_ms, ctx->ev[0], ctx->ev[1])); }
    return PG_OK;
}
