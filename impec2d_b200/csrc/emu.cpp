// emu.cpp -- CPU emulator behind shim.h (IMPEC_EMU builds only; test infrastructure).
// Blocks run one after another.  Non-cooperative kernels run their threads as plain nested loops;
// cooperative kernels (those that call __syncthreads or warp shuffles) run every thread of a block
// as a ucontext coroutine and switch at each barrier.
#ifdef IMPEC_EMU
#include "shim.h"
#include <ucontext.h>
#include <vector>
#include <cstdio>

namespace emu {
dim3 threadIdx_, blockIdx_, blockDim_, gridDim_;
unsigned char *dyn_smem = nullptr;

namespace {
struct Fiber {
    ucontext_t ctx;
    std::vector<unsigned char> stack;
    bool done = false;
    dim3 tid;
};
std::vector<Fiber> fibers;
ucontext_t sched_ctx;
int current = -1;
const std::function<void()> *cur_body = nullptr;
std::vector<double> exch;
bool in_coop = false;

void trampoline() {
    (*cur_body)();
    fibers[current].done = true;
    swapcontext(&fibers[current].ctx, &sched_ctx);
}
}  // namespace

void syncthreads() {
    if (!in_coop) { std::fprintf(stderr, "emu: __syncthreads in a kernel launched with IMPEC_LAUNCH (use IMPEC_LAUNCH_COOP)\n"); std::abort(); }
    swapcontext(&fibers[current].ctx, &sched_ctx);
    threadIdx_ = fibers[current].tid;
}

double shfl(double v, int arg, int mode) {
    const int nthreads = (int)fibers.size();
    const int me = current, lane = me & 31, base = me & ~31;
    exch[me] = v;
    syncthreads();
    int src = lane;
    if (mode == 0) src = lane + arg;
    else if (mode == 1) src = lane - arg;
    else src = arg;
    double out = v;
    if (src >= 0 && src < 32 && base + src < nthreads) out = exch[base + src];
    syncthreads();
    return out;
}

void launch(dim3 grid, dim3 block, size_t smem, bool coop, const std::function<void()> &body) {
    gridDim_ = grid; blockDim_ = block;
    std::vector<unsigned char> smem_buf(smem + 16);
    dyn_smem = smem_buf.data();
    const int nthreads = (int)(block.x * block.y * block.z);
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                blockIdx_ = dim3(bx, by, bz);
                if (!coop) {
                    for (unsigned tz = 0; tz < block.z; tz++)
                        for (unsigned ty = 0; ty < block.y; ty++)
                            for (unsigned tx = 0; tx < block.x; tx++) { threadIdx_ = dim3(tx, ty, tz); body(); }
                    continue;
                }
                in_coop = true; cur_body = &body;
                fibers.assign(nthreads, Fiber());
                exch.assign(nthreads, 0.0);
                int t = 0;
                for (unsigned tz = 0; tz < block.z; tz++)
                    for (unsigned ty = 0; ty < block.y; ty++)
                        for (unsigned tx = 0; tx < block.x; tx++, t++) {
                            Fiber &f = fibers[t];
                            f.tid = dim3(tx, ty, tz);
                            f.stack.resize(64 * 1024);
                            getcontext(&f.ctx);
                            f.ctx.uc_stack.ss_sp = f.stack.data();
                            f.ctx.uc_stack.ss_size = f.stack.size();
                            f.ctx.uc_link = &sched_ctx;
                            makecontext(&f.ctx, trampoline, 0);
                        }
                bool any = true;
                while (any) {
                    any = false;
                    for (int k = 0; k < nthreads; k++) {
                        if (fibers[k].done) continue;
                        current = k; threadIdx_ = fibers[k].tid;
                        swapcontext(&sched_ctx, &fibers[k].ctx);
                        if (!fibers[k].done) any = true;
                    }
                }
                in_coop = false; current = -1;
            }
    dyn_smem = nullptr;
}
}  // namespace emu
#endif
