// TEST INFRASTRUCTURE: scheduler of the warp emulator (see cuda_emu.h).
#include "cuda_emu.h"

namespace emu {

State g;

namespace {
struct LaneStart { void (*fn)(void *); void *arg; };
LaneStart g_start;
void lane_entry()
{
    g_start.fn(g_start.arg);
    Warp *w = g.w;
    w->done[w->cur] = true;
    swapcontext(&w->lane_ctx[w->cur], &w->main_ctx);
}
}  // namespace

void run_grid(void (*lane_fn)(void *), void *arg, unsigned grid, unsigned block_threads, size_t smem_bytes)
{
    const size_t kStack = 256 * 1024;
    const unsigned warps = (block_threads + 31) / 32;
    std::vector<unsigned char> smem(smem_bytes + 64);
    std::vector<uint32_t> tmem(128 * 512);
    Warp w;
    w.stacks.resize(32 * kStack);
    g.grid_dim = {grid, 1, 1};
    g.block_dim = {block_threads, 1, 1};
    g_start.fn = lane_fn;
    g_start.arg = arg;
    for (unsigned b = 0; b < grid; b++) {
        g.block_idx = {b, 0, 0};
        g.smem = (unsigned char *)(((uintptr_t)smem.data() + 15) & ~(uintptr_t)15);
        memset(g.smem, 0xA5, smem_bytes);          // uninitialised shared memory is garbage on the device too
        std::fill(tmem.begin(), tmem.end(), 0xDEADBEEFu);
        g.tmem = tmem.data();
        g.tmem_next_col = 0;
        for (unsigned wi = 0; wi < warps; wi++) {
            g.w = &w;
            w.warp_in_cta = (int)wi;
            for (int l = 0; l < 32; l++) {
                w.done[l] = false;
                w.seq[l] = 0;
                getcontext(&w.lane_ctx[l]);
                w.lane_ctx[l].uc_stack.ss_sp = w.stacks.data() + (size_t)l * kStack;
                w.lane_ctx[l].uc_stack.ss_size = kStack;
                w.lane_ctx[l].uc_link = &w.main_ctx;
                makecontext(&w.lane_ctx[l], (void (*)())lane_entry, 0);
            }
            bool alive = true;
            while (alive) {
                alive = false;
                for (int l = 0; l < 32; l++) {
                    if (w.done[l]) continue;
                    w.cur = l;
                    swapcontext(&w.main_ctx, &w.lane_ctx[l]);
                    alive = true;
                }
            }
        }
    }
    g.w = nullptr;
}

}  // namespace emu
