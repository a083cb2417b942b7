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

template <typename real, int LOGN, int L>
int run_case(const polar::ListArgs &a0, unsigned grid, std::vector<unsigned char> &scratch)
{
    using D = polar::ListDispatchCfg<real, LOGN, L>;
    using C = typename D::C;
    scratch.assign((size_t)C::GS_BYTES * grid + 64, 0xCD);
    Job j;
    j.a = a0;
    j.a.gscratch = scratch.data();
    j.kern = polar::list_decode_kernel<real, LOGN, L, D::SMEM_TOP, D::BITS_TOP>;
    emu::run_grid(lane_main, &j, grid, D::THREADS, C::SMEM * (D::THREADS / 32));
    return 0;
}
}  // namespace

// llr: [B][N] of the arithmetic type; info/cnt: [N/32]; crc_masks: [r][N/32]; u_hat: [B][N/32]; frame_info: [B]
extern "C" int emu_list_decode(int n, int L, int f64, const void *llr, unsigned long long B, const uint32_t *info, const uint32_t *cnt,
                               const uint32_t *crc_masks, int r, int use_crc, int coop_groups, unsigned grid, uint32_t *u_hat,
                               uint32_t *frame_info, unsigned long long *collectives)
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
    if (n == NN && L == LL) rc = f64 ? run_case<double, NN, LL>(a, grid, scratch) : run_case<float, NN, LL>(a, grid, scratch);
    X(5, 1) X(5, 4) X(6, 8) X(7, 1) X(7, 2) X(7, 8) X(7, 32) X(8, 4) X(9, 8) X(9, 16) X(10, 1) X(10, 2) X(10, 8) X(10, 16) X(10, 32)
#undef X
    if (collectives) *collectives = emu::g.collectives;
    return rc;
}
