/* emu_cuda.cpp -- fiber scheduler of the SIMT emulator (see emu_cuda.h).
 * TEST INFRASTRUCTURE ONLY. */
#include "emu_cuda.h"

emu_thread *emu_cur = nullptr;
emu_dim3 emu_blockDim = {1, 1, 1}, emu_gridDim = {1, 1, 1};

static void *sched_sp = nullptr;
static const std::function<void()> *cur_body = nullptr;

extern "C" void emu_switch(void **save_sp, void *load_sp);
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_switch,.-emu_switch
)");

void emu_yield(void)
{
    emu_thread *t = emu_cur;
    emu_switch(&t->sp, sched_sp);
}

static void fiber_main(void)
{
    (*cur_body)();
    emu_thread *t = emu_cur;
    t->done = 1;
    for (;;) emu_switch(&t->sp, sched_sp);
}

static const size_t STACK_BYTES = 96 * 1024;

void emu_launch(unsigned grid, unsigned block, size_t smem_bytes, const std::function<void()> &body)
{
    emu_gridDim = {grid, 1, 1};
    emu_blockDim = {block, 1, 1};
    cur_body = &body;
    const size_t nthreads = (size_t)grid * block;
    std::vector<emu_thread> threads(nthreads);
    std::vector<emu_block> blocks(grid);
    const int nwarps = (int)((block + 31) / 32);
    std::vector<emu_warp> warps((size_t)grid * nwarps);
    memset(warps.data(), 0, warps.size() * sizeof(emu_warp));
    for (unsigned b = 0; b < grid; b++) {
        emu_block &B = blocks[b];
        B.smem = (unsigned char *)aligned_alloc(128, ((smem_bytes + 127) / 128 + 1) * 128);
        memset(B.smem, 0xCD, smem_bytes);
        B.warps = &warps[(size_t)b * nwarps];
        B.nwarps = nwarps;
        B.bar_arrived = 0;
        B.bar_gen = 0;
        B.nthreads = (int)block;
        for (int w = 0; w < nwarps; w++) {
            const int n = (int)block - 32 * w;
            B.warps[w].nlanes = n > 32 ? 32 : n;
        }
        for (unsigned t = 0; t < block; t++) {
            emu_thread &T = threads[(size_t)b * block + t];
            T.tid = {t, 0, 0};
            T.bid = {b, 0, 0};
            T.block = &B;
            T.warp = &B.warps[t / 32];
            T.lane = (int)(t % 32);
            T.done = 0;
            T.stack = aligned_alloc(64, STACK_BYTES);
            uintptr_t top = ((uintptr_t)T.stack + STACK_BYTES) & ~(uintptr_t)15;
            void **sp = (void **)top;
            *--sp = nullptr;               /* fake return address: entry sees rsp % 16 == 8 */
            *--sp = (void *)&fiber_main;   /* popped by `ret` in emu_switch */
            for (int i = 0; i < 6; i++) *--sp = nullptr; /* rbp rbx r12 r13 r14 r15 */
            T.sp = (void *)sp;
        }
    }
    size_t remaining = nthreads;
    uint64_t idle_rounds = 0;
    while (remaining) {
        size_t progressed = 0;
        for (size_t i = 0; i < nthreads; i++) {
            emu_thread &T = threads[i];
            if (T.done == 2) continue;
            emu_cur = &T;
            emu_switch(&sched_sp, T.sp);
            if (T.done == 1) {
                T.done = 2;
                remaining--;
                progressed++;
            }
        }
        (void)progressed;
        if (++idle_rounds > (1ull << 34)) { fprintf(stderr, "emu: livelock\n"); abort(); }
    }
    emu_cur = nullptr;
    for (auto &T : threads) free(T.stack);
    for (auto &B : blocks) free(B.smem);
}
