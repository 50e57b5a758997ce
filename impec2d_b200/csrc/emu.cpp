// emu.cpp -- CPU emulator behind shim.h (IMPEC_EMU builds only; test infrastructure).
// Blocks run one after another.  Non-cooperative kernels run their threads as plain nested loops;
// cooperative kernels (those that call __syncthreads or warp shuffles) run every thread of a block
// as a coroutine and switch at each barrier.  The switch is a dozen instructions of x86-64 assembly
// (callee-saved registers + stack pointer; glibc's swapcontext makes a sigprocmask system call per
// switch, which dominated the run time of the CPU test suite) and the coroutine stacks are reused.
#ifdef IMPEC_EMU
#include "shim.h"
#include <vector>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <algorithm>

#if !defined(__x86_64__)
#error "the CPU emulator's context switch is written for x86-64"
#endif
extern "C" void impec_emu_switch(void **save_sp, void *load_sp);
asm(R"(
    .text
    .globl impec_emu_switch
    .hidden impec_emu_switch
    .type impec_emu_switch,@function
impec_emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    subq $8, %rsp
    stmxcsr (%rsp)
    fnstcw 4(%rsp)
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    ldmxcsr (%rsp)
    fldcw 4(%rsp)
    addq $8, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
    .size impec_emu_switch, .-impec_emu_switch
)");

namespace emu {
dim3 threadIdx_, blockIdx_, blockDim_, gridDim_;
unsigned char *dyn_smem = nullptr;

namespace {
const size_t kStack = 64 * 1024;
struct Fiber {
    void *sp = nullptr;
    bool done = false;
    dim3 tid;
};
std::vector<Fiber> fibers;
std::vector<unsigned char *> stacks;             // one per thread of the largest block seen so far; never freed
void *sched_sp = nullptr;
int current = -1;
const std::function<void()> *cur_body = nullptr;
std::vector<double> exch;
std::vector<int> shfl_count;                      // shuffles issued by every thread of the current block (all lanes of a warp must agree)
bool in_coop = false;

void trampoline() {
    (*cur_body)();
    fibers[current].done = true;
    impec_emu_switch(&fibers[current].sp, sched_sp);
    std::abort();                                 // a finished coroutine is never resumed
}
void prepare(Fiber &f, unsigned char *stack) {
    // initial frame as impec_emu_switch leaves it: [mxcsr | fpu cw][r15][r14][r13][r12][rbx][rbp][return address]
    uintptr_t top = ((uintptr_t)stack + kStack) & ~(uintptr_t)15;
    uint64_t *sp = (uint64_t *)(top - 16);        // slot of the return address: 16-byte aligned, so rsp % 16 == 8 on entry
    sp[0] = (uint64_t)(uintptr_t)&trampoline;
    sp[1] = 0;
    sp -= 7;
    for (int a = 1; a < 7; a++) sp[a] = 0;
    unsigned mx = 0; unsigned short cw = 0;
    asm volatile("stmxcsr %0" : "=m"(mx));
    asm volatile("fnstcw %0" : "=m"(cw));
    sp[0] = (uint64_t)mx | ((uint64_t)cw << 32);
    f.sp = sp;
}
}  // namespace

void syncthreads() {
    if (!in_coop) { std::fprintf(stderr, "emu: __syncthreads in a kernel launched with IMPEC_LAUNCH (use IMPEC_LAUNCH_COOP)\n"); std::abort(); }
    impec_emu_switch(&fibers[current].sp, sched_sp);
    threadIdx_ = fibers[current].tid;
}

double shfl(double v, int arg, int mode) {
    const int nthreads = (int)fibers.size();
    const int me = current, lane = me & 31, base = me & ~31;
    exch[me] = v;
    shfl_count[me]++;
    syncthreads();
    // the kernels shuffle with the full mask: every lane of the warp must still be running (a lane that returned early is undefined
    // behaviour on the GPU, where the emulator would silently hand out its last value)
    for (int l = 0; l < 32 && base + l < nthreads; l++)
        if (fibers[base + l].done || shfl_count[base + l] != shfl_count[me]) {
            std::fprintf(stderr, "emu: full-mask warp shuffle, but lane %d of warp %d %s\n", l, base / 32, fibers[base + l].done ? "has already exited" : "is not at the same shuffle (divergent control flow)");
            std::abort();
        }
    int src = lane;
    if (mode == 0) src = lane + arg;
    else if (mode == 1) src = lane - arg;
    else src = arg;
    double out = v;
    if (src >= 0 && src < 32 && base + src < nthreads) out = exch[base + src];
    syncthreads();
    return out;
}

// IMPEC_EMU_ORDER=reverse runs the blocks of a grid, the threads of a block and the coroutines between two barriers in the opposite order.
// A kernel without data races gives the same bits either way; one that reads what another thread or block writes in the same launch
// (an in-place stencil, a missing barrier) does not: tools/sanitize_emu.sh order runs the parity tests this way (a race check without a GPU).
// IMPEC_EMU_ORDER=shuffle: a pseudo-random permutation instead (a new one for every launch and for every round between two barriers).
static int order_mode() { static const int m = [] { const char *e = getenv("IMPEC_EMU_ORDER"); return !e ? 0 : (e[0] == 'r' ? 1 : (e[0] == 's' ? 2 : 0)); }(); return m; }
static unsigned long long rng_state = 0x9E3779B97F4A7C15ULL;
static unsigned rng() { rng_state = rng_state * 6364136223846793005ULL + 1442695040888963407ULL; return (unsigned)(rng_state >> 33); }
static void make_order(std::vector<int> &o, int n, int mode) {
    o.resize(n);
    for (int k = 0; k < n; k++) o[k] = mode == 1 ? n - 1 - k : k;
    if (mode == 2) for (int k = n - 1; k > 0; k--) std::swap(o[k], o[rng() % (unsigned)(k + 1)]);
}

void launch(dim3 grid, dim3 block, size_t smem, bool coop, const std::function<void()> &body) {
    gridDim_ = grid; blockDim_ = block;
    const int mode = order_mode();
    std::vector<int> border, torder;
    make_order(border, (int)(grid.x * grid.y * grid.z), mode);
    std::vector<unsigned char> smem_buf(smem + 16);
    dyn_smem = smem_buf.data();
    const int nthreads = (int)(block.x * block.y * block.z);
    if (coop) {
        while ((int)stacks.size() < nthreads) { void *p = nullptr; if (posix_memalign(&p, 64, kStack)) std::abort(); stacks.push_back((unsigned char *)p); }
        fibers.assign(nthreads, Fiber());
        exch.assign(nthreads, 0.0);
    }
    for (size_t bb = 0; bb < border.size(); bb++) {
            {
                const unsigned lin = (unsigned)border[bb];
                const unsigned bx = lin % grid.x, by = (lin / grid.x) % grid.y, bz = lin / (grid.x * grid.y);
                blockIdx_ = dim3(bx, by, bz);
                if (!coop) {
                    if (mode == 0) {
                        for (unsigned tz = 0; tz < block.z; tz++)
                            for (unsigned ty = 0; ty < block.y; ty++)
                                for (unsigned tx = 0; tx < block.x; tx++) { threadIdx_ = dim3(tx, ty, tz); body(); }
                    } else {
                        make_order(torder, nthreads, mode);
                        for (int t : torder) { threadIdx_ = dim3((unsigned)t % block.x, ((unsigned)t / block.x) % block.y, (unsigned)t / (block.x * block.y)); body(); }
                    }
                    continue;
                }
                in_coop = true; cur_body = &body;
                shfl_count.assign(nthreads, 0);
                int t = 0;
                for (unsigned tz = 0; tz < block.z; tz++)
                    for (unsigned ty = 0; ty < block.y; ty++)
                        for (unsigned tx = 0; tx < block.x; tx++, t++) {
                            Fiber &f = fibers[t];
                            f.tid = dim3(tx, ty, tz); f.done = false;
                            prepare(f, stacks[t]);
                        }
                bool any = true;
                while (any) {
                    any = false;
                    if (mode != 0) make_order(torder, nthreads, mode);
                    for (int k_ = 0; k_ < nthreads; k_++) {
                        const int k = mode == 0 ? k_ : torder[k_];
                        if (fibers[k].done) continue;
                        current = k; threadIdx_ = fibers[k].tid;
                        impec_emu_switch(&sched_sp, fibers[k].sp);
                        if (!fibers[k].done) any = true;
                    }
                }
                in_coop = false; current = -1;
            }
    }
    dyn_smem = nullptr;
}
}  // namespace emu
#endif
