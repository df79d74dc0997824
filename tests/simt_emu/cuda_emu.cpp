// TEST / DEBUG INFRASTRUCTURE ONLY — runtime of the lock-step SIMT emulator (see cuda_emu.h).
#include "cuda_emu.h"

namespace emu {
thread_local Fiber *cur = nullptr;
thread_local Cta *cta = nullptr;
thread_local uint3 t_blockIdx;
thread_local dim3 t_blockDim, t_gridDim;
thread_local char *t_dyn_smem = nullptr;
static thread_local const std::function<void()> *t_body = nullptr;

static const size_t kStack = 256 * 1024;

void yield() { swapcontext(&cur->ctx, &cta->sched); }

void warp_barrier(uint32_t mask) {
    Warp *w = cur->warp;
    Bar *b = nullptr;
    for (int i = 0; i < 8; i++)
        if (w->bars[i].mask == mask && w->bars[i].arrived > 0) { b = &w->bars[i]; break; }
    if (!b)
        for (int i = 0; i < 8; i++)
            if (w->bars[i].arrived == 0) { b = &w->bars[i]; b->mask = mask; break; }
    if (!b) { fprintf(stderr, "emu: out of warp barrier slots\n"); abort(); }
    if (!((mask >> cur->lane) & 1)) { fprintf(stderr, "emu: lane %d not in sync mask %08x\n", cur->lane, mask); abort(); }
    uint32_t gen = b->gen;
    if (++b->arrived == __builtin_popcount(mask)) {
        b->arrived = 0;
        b->gen++;
        cta->progress++;
    } else {
        while (b->gen == gen) yield();
    }
}

void cta_barrier() {
    Cta *c = cta;
    uint32_t gen = c->sync_gen;
    if (++c->sync_arrived == c->live) {
        c->sync_arrived = 0;
        c->sync_gen++;
        c->progress++;
    } else {
        while (c->sync_gen == gen) yield();
    }
}

static void fiber_entry() {
    (*t_body)();
    cur->done = true;
    cta->live--;
    cta->progress++;
    // a thread exiting counts as having arrived at __syncthreads for the rest (CUDA semantics for exited threads)
    if (cta->live > 0 && cta->sync_arrived == cta->live) {
        cta->sync_arrived = 0;
        cta->sync_gen++;
    }
    swapcontext(&cur->ctx, &cta->sched);
}

static void run_cta(Cta &c, dim3 block) {
    int n = c.nthreads;
    c.sync_arrived = 0;
    c.sync_gen = 0;
    c.live = n;
    c.progress = 0;
    for (auto &w : c.warps) memset(&w, 0, sizeof(w));
    for (int t = 0; t < n; t++) {
        Fiber &f = c.fibers[t];
        f.tid.x = t % block.x;
        f.tid.y = (t / block.x) % block.y;
        f.tid.z = t / (block.x * block.y);
        f.lane = t & 31;
        f.warp = &c.warps[t >> 5];
        f.done = false;
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack;
        f.ctx.uc_stack.ss_size = kStack;
        f.ctx.uc_link = &c.sched;
        makecontext(&f.ctx, (void (*)())fiber_entry, 0);
    }
    cta = &c;
    int remaining = n;
    while (remaining > 0) {
        uint64_t before = c.progress;
        remaining = 0;
        for (int t = 0; t < n; t++) {
            Fiber &f = c.fibers[t];
            if (f.done) continue;
            cur = &f;
            swapcontext(&c.sched, &f.ctx);
            if (!f.done) remaining++;
        }
        if (remaining > 0 && c.progress == before) {
            fprintf(stderr, "emu: deadlock in block (%u,%u): %d threads blocked (mismatched barrier/shuffle masks?)\n",
                    t_blockIdx.x, t_blockIdx.y, remaining);
            abort();
        }
    }
}

void launch(dim3 grid, dim3 block, size_t dyn_smem, const std::function<void()> &body) {
    size_t nblocks = (size_t)grid.x * grid.y * grid.z;
    int nthreads = block.x * block.y * block.z;
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 1;
    const char *env = getenv("NVB_EMU_THREADS");
    if (env) hw = (unsigned)atoi(env);
    if (hw > nblocks) hw = (unsigned)nblocks;
    if (hw == 0) return;
    std::atomic<size_t> next(0);
    auto worker = [&]() {
        Cta c;
        c.nthreads = nthreads;
        c.fibers.resize(nthreads);
        c.warps.resize((nthreads + 31) / 32);
        for (auto &f : c.fibers) f.stack = (char *)malloc(kStack);
        std::vector<char> smem(dyn_smem + 16);
        t_dyn_smem = smem.data();
        t_blockDim = block;
        t_gridDim = grid;
        t_body = &body;
        for (;;) {
            size_t b = next.fetch_add(1);
            if (b >= nblocks) break;
            t_blockIdx.x = (unsigned)(b % grid.x);
            t_blockIdx.y = (unsigned)((b / grid.x) % grid.y);
            t_blockIdx.z = (unsigned)(b / ((size_t)grid.x * grid.y));
            run_cta(c, block);
        }
        for (auto &f : c.fibers) free(f.stack);
    };
    if (hw == 1) {
        worker();
    } else {
        std::vector<std::thread> th;
        for (unsigned i = 0; i < hw; i++) th.emplace_back(worker);
        for (auto &t : th) t.join();
    }
}
}  // namespace emu
