// TEST INFRASTRUCTURE: runs polardecoding_b200/csrc/list_decode.cu on the CPU warp emulator (cuda_emu.h).
// Exports one C function for tests/test_emu_list.py; the kernel source is compiled unchanged with -DPOLAR_EMU.
#include "../../polardecoding_b200/csrc/list_decode.cu"

namespace {
struct Job { void (*kern)(const polar::ListArgs); polar::ListArgs a; };
void lane_main(void *p)
{
    Job *j = static_cast<Job *>(p);
    j->kern(j->a);
}

// tm: 0 = the configuration without tensor memory, 1 = the dispatcher's choice (fp32: stage 6 in tensor memory when N >= 256),
// 2 = tensor-memory layout forced for any arithmetic type (fp64 too, so that its logic can be held against the oracle bit for bit)
template <typename C, typename real, int LOGN, int L, int ST, int BT, int TML, int TMH>
int run_cfg(const polar::ListArgs &a0, unsigned grid, std::vector<unsigned char> &scratch)
{
    scratch.assign((size_t)C::GS_BYTES * C::WARPS * grid + 64, 0xCD);
    Job j;
    j.a = a0;
    j.a.gscratch = scratch.data();
    j.kern = polar::list_decode_kernel<real, LOGN, L, ST, BT, TML, TMH>;
    emu::run_grid(lane_main, &j, grid, C::THREADS, C::SMEM_CTA);
    return 0;
}
template <typename real, int LOGN, int L>
int run_case(const polar::ListArgs &a0, unsigned grid, std::vector<unsigned char> &scratch, int tm)
{
    if (tm >= 2) {  // forced tensor-memory layouts, any arithmetic type: 2 = stages 3..5 (no shared-memory stage), 3 = stage 6 only
        if constexpr (LOGN >= 8) {
            if (tm == 2) {
                using C = polar::ListCfg<real, LOGN, L, 3, POLAR_BITS_TOP, 3, 6>;
                return run_cfg<C, real, LOGN, L, 3, POLAR_BITS_TOP, 3, 6>(a0, grid, scratch);
            }
            using C = polar::ListCfg<real, LOGN, L, 6, POLAR_BITS_TOP, 6, 7>;
            return run_cfg<C, real, LOGN, L, 6, POLAR_BITS_TOP, 6, 7>(a0, grid, scratch);
        }
        return -2;
    }
    if (tm == 1) {
        using D = polar::ListDispatchCfg<real, LOGN, L, true>;
        return run_cfg<typename D::C, real, LOGN, L, D::SMEM_TOP, D::BITS_TOP, D::TML, D::TMH>(a0, grid, scratch);
    }
    using D = polar::ListDispatchCfg<real, LOGN, L, false>;
    return run_cfg<typename D::C, real, LOGN, L, D::SMEM_TOP, D::BITS_TOP, D::TML, D::TMH>(a0, grid, scratch);
}
}  // namespace

// llr: [B][N] of the arithmetic type; info/cnt: [N/32]; crc_masks: [r][N/32]; u_hat: [B][N/32]; frame_info: [B]
extern "C" int emu_list_decode(int n, int L, int f64, const void *llr, unsigned long long B, const uint32_t *info, const uint32_t *cnt,
                               const uint32_t *crc_masks, int r, int use_crc, int coop_groups, unsigned grid, uint32_t *u_hat,
                               uint32_t *frame_info, unsigned long long *collectives, int tm)
{
    polar::ListArgs a;
    memset(&a, 0, sizeof(a));
    a.llr = llr; a.u_hat = u_hat; a.frame_info = frame_info; a.crc_masks = crc_masks;
    a.B = B; a.r = r; a.use_crc = use_crc; a.coop_groups = coop_groups;
    const int W = (1 << n) / 32;
    for (int w = 0; w < W; w++) { a.m.info[w] = info[w]; a.m.cnt[w] = cnt[w]; }
    std::vector<unsigned char> scratch;
    emu::g.collectives = 0;
    int rc = -1;
#define X(NN, LL) \
    if (n == NN && L == LL) rc = f64 ? run_case<double, NN, LL>(a, grid, scratch, tm) : run_case<float, NN, LL>(a, grid, scratch, tm);
#ifdef EMU_FEW   // option flavours of the build (tests/emu_lib.py FLAVORS) only need the configurations their tests use
    X(9, 8) X(10, 1) X(10, 8)
#else
    X(5, 1) X(5, 4) X(6, 8) X(7, 1) X(7, 2) X(7, 8) X(7, 32) X(8, 4) X(9, 8) X(9, 16) X(10, 1) X(10, 2) X(10, 8) X(10, 16) X(10, 32)
#endif
#undef X
    if (collectives) *collectives = emu::g.collectives;
    return rc;
}
