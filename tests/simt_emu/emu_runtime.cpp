// TEST INFRASTRUCTURE ONLY (see cuda_runtime.h in this directory): fiber scheduler of the SIMT emulation.
#include "cuda_runtime.h"

namespace simt {

Block* g_block = nullptr;
unsigned char* g_dynsmem = nullptr;
// 0: lanes are resumed in ascending order within every scheduling round, 1: descending, 2: a fixed pseudo-random
// permutation per round.  A kernel whose threads only communicate across collectives / barriers gives bit-identical
// results under every order; one that relies on lock-step execution does not (a poor man's racecheck).
static int g_order = 0;
static const size_t STACK = 512 * 1024;

static void fiber_exit_bookkeeping(Block* b, Fiber* f) {
    f->done = true;
    b->live--;
    WarpCtx& w = b->warps[f->warp];
    w.live--;
    if (b->live > 0 && b->bar_arrived == b->live) { b->bar_arrived = 0; b->bar_gen++; }   // the others were waiting for this thread
}

static void trampoline() {
    Block* b = g_block;
    Fiber* f = b->cur;
    b->body();
    fiber_exit_bookkeeping(b, f);
    swapcontext(&f->ctx, &b->sched);      // never resumed
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    const int nthreads = (int)(block.x * block.y * block.z);
    const int nwarps = (nthreads + 31) / 32;
    std::vector<char> stacks((size_t)nthreads * STACK);
    std::vector<unsigned char> dyn(smem + 64);
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                Block b;
                b.fibers.resize(nthreads);
                b.warps.resize(nwarps);
                b.nthreads = b.live = nthreads;
                b.bid = dim3(bx, by, bz); b.bdim = block; b.gdim = grid;
                b.body = body;
                for (int w = 0; w < nwarps; w++) b.warps[w].live = std::min(32, nthreads - 32 * w);
                g_block = &b;
                g_dynsmem = (unsigned char*)(((uintptr_t)dyn.data() + 63) & ~(uintptr_t)63);
                for (int t = 0; t < nthreads; t++) {
                    Fiber& f = b.fibers[t];
                    f.tid = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
                    f.lane = t & 31; f.warp = t >> 5;
                    getcontext(&f.ctx);
                    f.ctx.uc_stack.ss_sp = stacks.data() + (size_t)t * STACK;
                    f.ctx.uc_stack.ss_size = STACK;
                    f.ctx.uc_link = &b.sched;
                    makecontext(&f.ctx, trampoline, 0);
                }
                long rounds = 0;
                uint32_t lcg = 12345u;
                while (b.live > 0) {
                    if (g_order == 2) lcg = lcg * 1664525u + 1013904223u;
                    const int off = g_order == 2 ? (int)((lcg >> 8) % (uint32_t)nthreads) : 0;
                    const int stride = g_order == 2 ? 1 + 2 * (int)((lcg >> 20) % 8u) : 1;     // odd stride: a permutation when nthreads is a power of two
                    for (int i = 0; i < nthreads; i++) {
                        int t = g_order == 1 ? nthreads - 1 - i : (g_order == 2 && (nthreads & (nthreads - 1)) == 0 ? (off + i * stride) & (nthreads - 1) : i);
                        Fiber& f = b.fibers[t];
                        if (f.done) continue;
                        b.cur = &f;
                        swapcontext(&b.sched, &f.ctx);
                    }
                    if (++rounds > 50000000L) { fprintf(stderr, "simt_emu: block (%u,%u) does not terminate (divergent collective?)\n", bx, by); abort(); }
                }
                g_block = nullptr;
            }
}

}  // namespace simt

extern "C" void simt_emu_set_order(int order) { simt::g_order = order; }
