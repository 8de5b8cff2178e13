// simt_emul.h -- TEST INFRASTRUCTURE ONLY: a cooperative (fiber based) emulation of one CUDA warp on the host.
//
// The throughput kernel of the product (acados_b200/csrc/cuipm_fast_core.h) is written against a handful of
// warp primitives (lane id, warp barrier, shuffles, votes, asynchronous global->shared copies).  On the GPU they
// map to the hardware instructions (acados_b200/csrc/cuipm_fast.cu); here they map to 32 ucontext fibers that are
// switched at every barrier, so that the SAME kernel body runs on the CPU of a machine without a GPU and can be
// compared with the oracle (tests/test_fast_emul.py).  Lanes run one after the other between barriers, in an order
// that can be reversed (SIMT_EMUL_ORDER) to expose missing barriers in either direction; asynchronous copies are
// deferred until the wait, so that reading a staged block before the wait shows up as a wrong result.
// Nothing in the product includes this file.
#ifndef SIMT_EMUL_H_
#define SIMT_EMUL_H_

#include <ucontext.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

namespace simt {

struct PendingCopy { void *dst; const void *src; int bytes; };

struct Warp
{
    static constexpr int NL = 32;
    ucontext_t main_ctx;
    ucontext_t ctx[NL];
    std::vector<char> stack[NL];
    bool done[NL];
    int cur = 0;
    int gen = 0, arrived = 0;
    int order = 0;                 // 0: lanes 0..31, 1: lanes 31..0
    double xd[NL], xd2[NL];
    int xi[NL];
    std::vector<PendingCopy> pend[NL];
    std::function<void()> body;
    long barriers = 0;
    long events = 0;               // progress other than barriers (transaction barriers completing)
};

// emulation of an mbarrier with transaction count (bulk asynchronous copies): the copies are queued and performed by
// the first lane that observes the completed phase, so that a lane reading a block before its wait sees stale data
struct MBar
{
    int phase = 0;                 // parity of the phase being filled
    long tx = 0;                   // bytes expected minus bytes announced by copies
    int pending_arrivals = 1;
    bool armed = false;
    std::vector<PendingCopy> copies;
};

inline Warp *&cur_warp()
{
    static thread_local Warp *w = nullptr;
    return w;
}

inline int lane() { return cur_warp()->cur; }

inline void yield_()
{
    Warp *w = cur_warp();
    swapcontext(&w->ctx[w->cur], &w->main_ctx);
}

// warp barrier: every lane must call it the same number of times
inline void sync()
{
    Warp *w = cur_warp();
    const int g = w->gen;
    if (++w->arrived == Warp::NL)
    {
        w->arrived = 0;
        w->gen++;
        w->barriers++;
    }
    while (w->gen == g) yield_();
}

inline double shfl(double v, int src)
{
    Warp *w = cur_warp();
    w->xd[w->cur] = v;
    sync();
    const double r = w->xd[src & 31];
    sync();
    return r;
}
inline int shfl_i(int v, int src)
{
    Warp *w = cur_warp();
    w->xi[w->cur] = v;
    sync();
    const int r = w->xi[src & 31];
    sync();
    return r;
}
// D = A B + C of the m8n8k4 tensor-core product: lane l holds A[l / 4][l % 4], B[l % 4][l / 4], C[l / 4][2 (l % 4) .. + 1]
inline void dmma(double &c0, double &c1, double a, double b)
{
    Warp *w = cur_warp();
    w->xd[w->cur] = a;
    w->xd2[w->cur] = b;
    sync();
    const int row = w->cur >> 2, col = 2 * (w->cur & 3);
    for (int k = 0; k < 4; k++)
    {
        c0 += w->xd[row * 4 + k] * w->xd2[col * 4 + k];
        c1 += w->xd[row * 4 + k] * w->xd2[(col + 1) * 4 + k];
    }
    sync();
}
inline unsigned ballot(bool p)
{
    Warp *w = cur_warp();
    w->xi[w->cur] = p ? 1 : 0;
    sync();
    unsigned m = 0;
    for (int i = 0; i < Warp::NL; i++) m |= (unsigned) w->xi[i] << i;
    sync();
    return m;
}
inline void cp_async(void *dst, const void *src, int bytes)
{
    Warp *w = cur_warp();
    static const bool eager = std::getenv("SIMT_EMUL_EAGER") != nullptr;       // see bulk_copy
    if (eager) { std::memcpy(dst, src, (size_t) bytes); return; }
    w->pend[w->cur].push_back({dst, src, bytes});
}
inline void cp_wait()
{
    Warp *w = cur_warp();
    for (const PendingCopy &c : w->pend[w->cur]) std::memcpy(c.dst, c.src, (size_t) c.bytes);
    w->pend[w->cur].clear();
}

inline void mbar_init(MBar *b, int count) { *b = MBar(); b->pending_arrivals = count; }
inline void bulk_copy(MBar *b, void *dst, const void *src, int bytes)
{
    if (((size_t) dst & 15) || ((size_t) src & 15) || (bytes & 15)) { std::fprintf(stderr, "simt_emul: misaligned bulk copy\n"); std::abort(); }
    // a bulk copy may land at any time between its issue and the wait: SIMT_EMUL_EAGER=1 delivers it at once (a buffer that is
    // overwritten while still in use shows up), the default delivers it at the wait (a buffer that is read too early shows up)
    static const bool eager = std::getenv("SIMT_EMUL_EAGER") != nullptr;
    if (eager) std::memcpy(dst, src, (size_t) bytes);
    else b->copies.push_back({dst, src, bytes});
    b->tx -= bytes;
}
// arrive (one of `count` arrivals) and expect `bytes` more bytes of copies
inline void mbar_arrive_tx(MBar *b, int bytes)
{
    b->tx += bytes;
    b->pending_arrivals--;
    cur_warp()->events++;
}
// waits for the phase of parity `parity` to complete (all arrivals in, all announced bytes delivered)
inline void mbar_wait(MBar *b, int parity)
{
    while (b->phase == parity && !(b->pending_arrivals == 0 && b->tx == 0)) yield_();
    if (b->phase == parity)
    {   // first lane to observe completion: deliver the data, open the next phase
        for (const PendingCopy &c : b->copies) std::memcpy(c.dst, c.src, (size_t) c.bytes);
        b->copies.clear();
        b->phase ^= 1;
        b->pending_arrivals = 1;
        cur_warp()->events++;
    }
}

inline void trampoline_()
{
    Warp *w = cur_warp();
    w->body();
    w->done[w->cur] = true;
    swapcontext(&w->ctx[w->cur], &w->main_ctx);
}

// runs body() once per lane as one warp; returns the number of barriers executed
inline long run_warp(const std::function<void()> &body, int order = 0, size_t stack_bytes = 1 << 20)
{
    Warp w;
    w.body = body;
    w.order = order;
    Warp *prev = cur_warp();
    cur_warp() = &w;
    for (int i = 0; i < Warp::NL; i++)
    {
        w.stack[i].resize(stack_bytes);
        w.done[i] = false;
        getcontext(&w.ctx[i]);
        w.ctx[i].uc_stack.ss_sp = w.stack[i].data();
        w.ctx[i].uc_stack.ss_size = stack_bytes;
        w.ctx[i].uc_link = &w.main_ctx;
        makecontext(&w.ctx[i], (void (*)()) trampoline_, 0);
    }
    for (;;)
    {
        bool all = true;
        const int g0 = w.gen, a0 = w.arrived;
        const long e0 = w.events;
        int ndone0 = 0;
        for (int i = 0; i < Warp::NL; i++) ndone0 += w.done[i];
        for (int s = 0; s < Warp::NL; s++)
        {
            const int i = w.order ? Warp::NL - 1 - s : s;
            if (w.done[i]) continue;
            all = false;
            w.cur = i;
            swapcontext(&w.main_ctx, &w.ctx[i]);
        }
        if (all) break;
        int ndone1 = 0;
        for (int i = 0; i < Warp::NL; i++) ndone1 += w.done[i];
        if (w.gen == g0 && ndone1 == ndone0 && w.arrived == a0 && w.events == e0)
        {
            std::fprintf(stderr, "simt_emul: no progress (divergent barrier: %d lanes wait, %d finished)\n", w.arrived, ndone1);
            std::abort();
        }
    }
    cur_warp() = prev;
    return w.barriers;
}

}  // namespace simt
#endif
