// tests/emu/emu_runtime.cpp — TEST INFRASTRUCTURE: block / thread scheduler of the CPU stand-in (see cuda_runtime.h).
#include <sys/mman.h>
#include <vector>
#include "cuda_runtime.h"

// Context switch between the scheduler and a thread's fiber: callee-saved registers and the stack pointer
// only (glibc's swapcontext also saves the signal mask with a system call per switch, which made a
// barrier cost microseconds).  x86-64 System V.
extern "C" void emu_switch(void** save_sp, void* load_sp);
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

namespace emu {

thread_local uint3 t_threadIdx, t_blockIdx;
thread_local dim3 t_blockDim, t_gridDim;

namespace {
enum State { RUNNABLE, WAIT_BLOCK, WAIT_WARP, DONE };
struct FiberRec { void* sp = nullptr; char* stack = nullptr; State st = DONE; unsigned tid = 0; };
const size_t kStack = 256 << 10;
thread_local std::vector<FiberRec*> t_fibers;
thread_local void* t_main_sp = nullptr;
thread_local FiberRec* t_cur = nullptr;
thread_local const std::function<void()>* t_body = nullptr;
thread_local uint32_t t_warp_buf[32][32];            // [warp][lane] exchange slots of the shuffles

void trampoline() {
    (*t_body)();
    t_cur->st = DONE;
    emu_switch(&t_cur->sp, t_main_sp);
    abort();                                          // a finished fiber is never resumed
}
void yield(State s) {
    FiberRec* f = t_cur;
    f->st = s;
    emu_switch(&f->sp, t_main_sp);
    t_threadIdx = uint3{f->tid, 0, 0};               // restored by the scheduler too; kept for clarity
}
void run_block(unsigned n_threads) {
    while (t_fibers.size() < n_threads) {
        FiberRec* f = new FiberRec();
        f->stack = (char*)mmap(nullptr, kStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_STACK, -1, 0);
        if (f->stack == MAP_FAILED) { fprintf(stderr, "emu: cannot map a fiber stack\n"); abort(); }
        t_fibers.push_back(f);
    }
    for (unsigned i = 0; i < n_threads; ++i) {
        FiberRec* f = t_fibers[i];
        // initial frame: six callee-saved registers, then the address emu_switch returns to; after that `ret`
        // the stack pointer must look like just after a call (8 mod 16)
        void** top = (void**)(((uintptr_t)(f->stack + kStack) & ~(uintptr_t)15) - 8);
        *--top = (void*)&trampoline;
        for (int r = 0; r < 6; ++r) *--top = nullptr;
        f->sp = top;
        f->st = RUNNABLE; f->tid = i;
    }
    for (;;) {
        bool any_live = false, progressed = false;
        for (unsigned i = 0; i < n_threads; ++i) {
            FiberRec* f = t_fibers[i];
            if (f->st != RUNNABLE) { any_live |= f->st != DONE; continue; }
            t_cur = f; t_threadIdx = uint3{i, 0, 0};
            emu_switch(&t_main_sp, f->sp);
            progressed = true;
            any_live |= f->st != DONE;
        }
        if (!any_live) return;
        // open the barriers whose participants have all arrived (exited threads count as arrived)
        bool opened = false;
        for (unsigned w = 0; w * 32 < n_threads; ++w) {
            bool all = true, some = false;
            for (unsigned l = w * 32; l < n_threads && l < w * 32 + 32; ++l) {
                const State s = t_fibers[l]->st;
                if (s == WAIT_WARP) some = true; else if (s != DONE) all = false;
            }
            if (all && some) { for (unsigned l = w * 32; l < n_threads && l < w * 32 + 32; ++l) if (t_fibers[l]->st == WAIT_WARP) t_fibers[l]->st = RUNNABLE; opened = true; }
        }
        if (!opened) {
            bool all = true, some = false;
            for (unsigned i = 0; i < n_threads; ++i) { const State s = t_fibers[i]->st; if (s == WAIT_BLOCK) some = true; else if (s != DONE) all = false; }
            if (all && some) { for (unsigned i = 0; i < n_threads; ++i) if (t_fibers[i]->st == WAIT_BLOCK) t_fibers[i]->st = RUNNABLE; opened = true; }
        }
        if (!opened && !progressed) { fprintf(stderr, "emu: deadlock (divergent barrier) in block (%u,%u)\n", t_blockIdx.x, t_blockIdx.y); abort(); }
        if (!opened) {
            bool runnable = false;
            for (unsigned i = 0; i < n_threads; ++i) runnable |= t_fibers[i]->st == RUNNABLE;
            if (!runnable) { fprintf(stderr, "emu: deadlock (threads wait at different barriers) in block (%u,%u)\n", t_blockIdx.x, t_blockIdx.y); abort(); }
        }
    }
}
}  // namespace

void launch(dim3 grid, dim3 block, const std::function<void()>& body) {
    static const bool skip = getenv("EMU_SKIP_KERNELS") != nullptr;      // host-side timing runs: plan and submit only
    if (skip) return;
    if (block.y != 1 || block.z != 1 || block.x > 1024) { fprintf(stderr, "emu: only 1-D blocks of <= 1024 threads\n"); abort(); }
    t_body = &body; t_blockDim = block; t_gridDim = grid;
    for (unsigned z = 0; z < grid.z; ++z) for (unsigned y = 0; y < grid.y; ++y) for (unsigned x = 0; x < grid.x; ++x) {
        t_blockIdx = uint3{x, y, z};
        run_block(block.x);
    }
    t_body = nullptr;
}

void syncthreads() { yield(WAIT_BLOCK); }

uint32_t shfl(uint32_t v, int a, int mode) {
    const unsigned tid = t_cur->tid, w = tid >> 5, lane = tid & 31;
    t_warp_buf[w][lane] = v;
    yield(WAIT_WARP);
    int src = mode == 0 ? (int)lane - a : mode == 1 ? (int)(lane ^ (unsigned)a) : mode == 2 ? (a & 31) : (int)lane + a;
    const uint32_t r = (src < 0 || src > 31) ? v : t_warp_buf[w][src];
    yield(WAIT_WARP);                                  // nobody overwrites a slot before every lane has read
    return r;
}

}  // namespace emu
