// TEST INFRASTRUCTURE ONLY (see cuda_runtime.h in this directory): fiber scheduler of the SIMT emulation.
#include "cuda_runtime.h"

namespace simt {

Block* g_block = nullptr;
unsigned char* g_dynsmem = nullptr;
unsigned long long g_collectives = 0;
// 0: lanes are resumed in ascending order within every scheduling round, 1: descending, 2: a fixed pseudo-random
// permutation per round.  A kernel whose threads only communicate across collectives / barriers gives bit-identical
// results under every order; one that relies on lock-step execution does not (a poor man's racecheck).
static int g_order = 0;
static const size_t STACK = 512 * 1024;

#if defined(__x86_64__)
#define SIMT_ASM_SWITCH 1
// save the callee-saved registers on the current stack, store its pointer in *from, continue on the stack *to
extern "C" void simt_swap(void** from, void** to);
asm(R"(
.text
.globl simt_swap
.type simt_swap,@function
simt_swap:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq (%rsi), %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size simt_swap,.-simt_swap
)");
#endif

void yield() {
    Block* b = g_block;
    Fiber* f = b->cur;
#ifdef SIMT_ASM_SWITCH
    simt_swap(&f->sp, &b->sched_sp);
#else
    swapcontext(&f->ctx, &b->sched);
#endif
}

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
    yield();                              // never resumed
    abort();
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    const int nthreads = (int)(block.x * block.y * block.z);
    const int nwarps = (nthreads + 31) / 32;
    // fiber stacks and the dynamic shared-memory window are kept between launches (a fresh std::vector would zero-fill
    // 512 KB per thread on every launch)
    static char* stack_pool = nullptr; static size_t stack_pool_size = 0;
    static unsigned char* dyn_pool = nullptr; static size_t dyn_pool_size = 0;
    if ((size_t)nthreads * STACK > stack_pool_size) { free(stack_pool); stack_pool_size = (size_t)nthreads * STACK; stack_pool = (char*)malloc(stack_pool_size); }
    if (smem + 64 > dyn_pool_size) { free(dyn_pool); dyn_pool_size = smem + 64; dyn_pool = (unsigned char*)malloc(dyn_pool_size); }
    struct { char* p; char* data() { return p; } } stacks{ stack_pool };
    struct { unsigned char* p; unsigned char* data() { return p; } } dyn{ dyn_pool };
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
#ifdef SIMT_ASM_SWITCH
                    {   // initial frame: six zeroed callee-saved registers, then the entry point as the return address;
                        // after the `ret` the stack pointer is 8 below a 16-byte boundary, as at any function entry
                        uintptr_t top = ((uintptr_t)(stacks.data() + (size_t)(t + 1) * STACK)) & ~(uintptr_t)15;
                        void** sp = (void**)(top - 8);
                        *sp = nullptr;
                        *--sp = (void*)trampoline;
                        for (int r = 0; r < 6; r++) *--sp = nullptr;
                        f.sp = sp;
                    }
#else
                    getcontext(&f.ctx);
                    f.ctx.uc_stack.ss_sp = stacks.data() + (size_t)t * STACK;
                    f.ctx.uc_stack.ss_size = STACK;
                    f.ctx.uc_link = &b.sched;
                    makecontext(&f.ctx, trampoline, 0);
#endif
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
#ifdef SIMT_ASM_SWITCH
                        simt_swap(&b.sched_sp, &f.sp);
#else
                        swapcontext(&b.sched, &f.ctx);
#endif
                    }
                    if (++rounds > 50000000L) { fprintf(stderr, "simt_emu: block (%u,%u) does not terminate (divergent collective?)\n", bx, by); abort(); }
                }
                g_block = nullptr;
            }
}

}  // namespace simt

extern "C" void simt_emu_set_order(int order) { simt::g_order = order; }
extern "C" unsigned long long simt_emu_collectives() { return simt::g_collectives; }
