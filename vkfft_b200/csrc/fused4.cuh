// Fused Four-Step: both passes of N = n1 * n2 in ONE persistent launch, the intermediate kept in L2.
//
// The reference (and this engine's two-launch plan) writes the whole intermediate to DRAM in the first dispatch and
// reads it back in the second (vkFFT_Scheduler.h:2582-2893: "numAxisUploads = 2"), i.e. two HBM round trips, which caps
// those sizes at 0.5 of the copy roofline.  Here:
//   * sequences are grouped into UNITS; the scratch is a ring of R units (a few MB ... tens of MB, far below the 126 MB
//     L2), rewritten over and over, so it lives in L2: pass A's stores and pass B's loads never reach HBM;
//   * persistent CTAs take TILES from two ordered queues.  A tile of pass A (Q_A neighbouring columns of one sequence:
//     strided n1-point transforms + the Four-Step phase) may be taken when its unit's ring slot is free; a tile of pass B
//     (Q_B rows: contiguous n2-point transforms, transposed store to the final place) when ALL A tiles of its unit are
//     done.  All tiles form ONE ordered sequence (pass A of unit u interleaved with pass B of unit u-L) that the CTAs walk
//     with fetch-add tickets, so pass B trails pass A by L units and reads and writes overlap; all scheduler traffic (the
//     next ticket, the two readiness limits, the release of the previous tile) is issued while the current tile's first
//     loads are in flight (Engine::run_at hook) and only looked at after its last store, so no memory round trip sits on
//     the tile path;
//   * pass B reads the scratch with ld.global.cg (it was written by other SMs in this launch) and, once the legs are in
//     registers, drops the lines from L2 (discard.global.L2) so that dead scratch is never written back to HBM.
// Both passes run the very same stage code as the stand-alone kernels (Engine<C>::run_at), so results are identical
// to the two-launch plan bit for bit.
#pragma once
#include "stockham.cuh"
#include "pipe.cuh"

namespace b200fft {

// Memory ordering of the tile protocol.  Producers: the CTA's stores, bar.sync, then thread 0 bumps the unit's done-counter
// with RELEASE semantics (one MEMBAR.GPU per tile: the tile's stores / discards are performed at L2 before the count
// moves).  Consumers: every claim word is read and updated with RELAXED operations at L2 (no fence, no L1 flush --
// an acquire here would put a MEMBAR + CCTL.IVALL on the per-tile path and throw the twiddle tables out of L1), and
// the data itself is then read with ld.global.cg, i.e. from L2, where the producer's fence has already put it; the
// addresses depend on the claimed ticket, so the loads cannot be issued early.  Only the once-per-unit prefix advance
// uses acquire loads.
#if defined(__CUDA_ARCH__)
B2_D uint32_t fz_ld_acquire(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
B2_D uint32_t fz_ld_relaxed(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
B2_D uint32_t fz_add(uint32_t* p, uint32_t v) {       // relaxed fetch-add
    uint32_t o;
    asm volatile("atom.relaxed.gpu.global.add.u32 %0, [%1], %2;" : "=r"(o) : "l"(p), "r"(v) : "memory");
    return o;
}
B2_D uint32_t fz_add_release(uint32_t* p, uint32_t v) {
    uint32_t o;
    asm volatile("atom.release.gpu.global.add.u32 %0, [%1], %2;" : "=r"(o) : "l"(p), "r"(v) : "memory");
    return o;
}
B2_D uint32_t fz_cas(uint32_t* p, uint32_t cmp, uint32_t val) {
    uint32_t o;
    asm volatile("atom.relaxed.gpu.global.cas.b32 %0, [%1], %2, %3;" : "=r"(o) : "l"(p), "r"(cmp), "r"(val) : "memory");
    return o;
}
B2_D uint32_t fz_cas_release(uint32_t* p, uint32_t cmp, uint32_t val) {
    uint32_t o;
    asm volatile("atom.release.gpu.global.cas.b32 %0, [%1], %2, %3;" : "=r"(o) : "l"(p), "r"(cmp), "r"(val) : "memory");
    return o;
}
B2_D void fz_sleep(unsigned ns) { __nanosleep(ns); }
#else
// CPU emulation: one CTA at a time, only thread 0 touches the control words
B2_D uint32_t fz_ld_acquire(const uint32_t* p) { return *p; }
B2_D uint32_t fz_ld_relaxed(const uint32_t* p) { return *p; }
B2_D uint32_t fz_add(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o + v; return o; }
B2_D uint32_t fz_add_release(uint32_t* p, uint32_t v) { return fz_add(p, v); }
B2_D uint32_t fz_cas(uint32_t* p, uint32_t cmp, uint32_t val) { uint32_t o = *p; if (o == cmp) *p = val; return o; }
B2_D uint32_t fz_cas_release(uint32_t* p, uint32_t cmp, uint32_t val) { return fz_cas(p, cmp, val); }
B2_D void fz_sleep(unsigned) {}
#endif

// NBUF: tile buffers per CTA.  2: the NEXT tile is copied into the other buffer by TMA while this one is transformed;
// 1 (tiles too large to double): the next tile is copied into the same buffer as soon as the last stage has its legs in
// registers, i.e. it overlaps the last butterflies and the global stores.
template <class CA, class CB, int NBUF = 2>
struct Fused4 {
    // strides known at compile time: pass A walks columns of an n1 x n2 matrix (element stride n2 on both sides), pass B
    // stores transposed (element stride n1)
    using EA = Engine<CA, 0, CB::N, CB::N>;
    using EB = Engine<CB, XF_LDCG | XF_DISCARD, 0, CA::N>;
    using T = typename CA::T;
    using X = cpx<T>;
    static_assert(CA::THREADS == CB::THREADS, "both passes run in the same CTA shape");
    static_assert(CA::LAYOUT == LAY_ELEM && CB::LAYOUT == LAY_LINE, "pass A: interleaved columns, pass B: contiguous rows");
    static_assert(CA::V == 1 && CB::V == 1, "one butterfly per thread and step");
    static_assert(NBUF == 1 || NBUF == 2, "one or two tile buffers");
    static constexpr int THREADS = CA::THREADS;
    static constexpr int MINB = CA::MINB < CB::MINB ? CA::MINB : CB::MINB;
    static constexpr int TILE_BYTES = ((CA::SMEM_BYTES > CB::SMEM_BYTES ? CA::SMEM_BYTES : CB::SMEM_BYTES) + 127) / 128 * 128;
    static constexpr int SMEM_BYTES = NBUF * TILE_BYTES + 64;     // + two mbarriers + mailbox: two slots of (kind, unit, tile)
    static constexpr uint32_t BYTES_A = (uint32_t)(CA::N * CA::Q * sizeof(X)), BYTES_B = (uint32_t)(CB::N * CB::Q * sizeof(X));
    enum { NONE = 0, TILE_A = 1, TILE_B = 2 };

    // ---- scheduler: ONE ordered queue, state in the registers of thread 0 --------------------------------------------------
    // All tiles of the launch form one sequence that every CTA walks with a fetch-add ticket:
    //      block b = 0 .. NU+L-1 :   the pass-A tiles of unit b  interleaved with  the pass-B tiles of unit b-L
    // so pass B runs L units behind pass A -- far enough that a unit's pass-A tiles have normally all finished by the time
    // its first pass-B tile is handed out (L is sized from the number of resident CTAs by the planner), and HBM reads
    // (pass A) and HBM writes (pass B) are always in flight together.  A tile may START when
    //      pass B:  unit <  READY_UNITS            (in-order prefix of units whose pass A is complete)
    //      pass A:  unit <  FREED_UNITS + R        (the ring slot it overwrites has been drained; R > L)
    // otherwise thread 0 polls the two words (both monotonic).  A tile only ever waits for tiles EARLIER in the sequence and
    // a CTA holds one started tile plus one ticket it has not started, so the lowest unfinished tile can always run: no
    // deadlock, whatever the number of resident CTAs.
    // Nothing here waits for a memory round trip on the tile path: the next ticket, the refresh of the two limits and the
    // release of the PREVIOUS tile (one MEMBAR.GPU + one atomic) are issued while the tile's first loads are in flight
    // (Engine hook) and only looked at after its last store.
    struct Sched {
        uint32_t ready, freed;         // last values read of READY_UNITS / FREED_UNITS
        uint32_t pend_kind, pend_unit; // finished tile whose completion has not been published yet
        uint32_t newT, dcount;         // in flight: next-next ticket, done-counter before this CTA's increment
        uint32_t done_kind, done_unit;
    };

    // ticket -> (pass, unit, tile within the unit's pass)
    B2_D static uint32_t decode(const b2_fused_params& F, uint32_t item, uint32_t& unit, uint32_t& tile) {
        const uint32_t TA = F.TA, TB = F.TB, NU = F.NU, L = F.reserved;       // reserved = lead L (units)
        const uint32_t headA = L * TA, mid = (NU - L) * (TA + TB);
        if (item < headA) { unit = item / TA; tile = item % TA; return TILE_A; }
        if (item < headA + mid) {
            const uint32_t r = item - headA, blk = L + r / (TA + TB), pos = r % (TA + TB);
            const uint32_t a0 = (uint32_t)(((uint64_t)pos * TA) / (TA + TB)), a1 = (uint32_t)(((uint64_t)(pos + 1) * TA) / (TA + TB));
            if (a1 > a0) { unit = blk; tile = a0; return TILE_A; }
            unit = blk - L; tile = pos - a1; return TILE_B;
        }
        const uint32_t r = item - headA - mid;
        if (r >= L * TB) { unit = 0; tile = 0; return NONE; }
        unit = NU - L + r / TB; tile = r % TB;
        return TILE_B;
    }
    B2_D static bool runnable(const b2_fused_params& F, const Sched& S, uint32_t kind, uint32_t unit) {
        return kind == TILE_B ? unit < S.ready : unit < S.freed + F.R;
    }

    B2_D static void publish(const b2_fused_params& F, Sched& S) {      // release the pending tile (if any); async result in S.dcount
        S.done_kind = S.pend_kind; S.done_unit = S.pend_unit;
        if (S.pend_kind != NONE) {
            uint32_t* done = F.ctl + B2_FCTL_WORDS + (S.pend_kind == TILE_B ? F.NU : 0u);
            S.dcount = fz_add_release(done + S.pend_unit, 1u);
            S.pend_kind = NONE;
        }
    }
    // the unit just published was completed by this CTA's increment: advance the in-order prefix (once per unit and pass)
    B2_D static void after_publish(const b2_fused_params& F, Sched& S) {
        if (S.done_kind == NONE) return;
        const uint32_t per_unit = S.done_kind == TILE_A ? F.TA : F.TB;
        if (S.dcount + 1 != per_unit) { S.done_kind = NONE; return; }
        const uint32_t word = S.done_kind == TILE_A ? B2_FCTL_READY_UNITS : B2_FCTL_FREED_UNITS;
        const uint32_t* done = F.ctl + B2_FCTL_WORDS + (S.done_kind == TILE_B ? F.NU : 0u);
        S.done_kind = NONE;
        for (;;) {
            const uint32_t p = fz_ld_acquire(F.ctl + word);
            if (p >= F.NU) return;
            if (fz_ld_acquire(done + p) != per_unit) return;
            fz_cas_release(F.ctl + word, p, p + 1);
        }
    }

    // scheduler traffic of one tile, issued by thread 0 while the tile's first-stage legs are being read
    B2_D static void sched_issue(const b2_fused_params& F, Sched& S) {
        S.newT = fz_add(F.ctl + B2_FCTL_NEXT_A, 1u);
        S.ready = fz_ld_relaxed(F.ctl + B2_FCTL_READY_UNITS);
        S.freed = fz_ld_relaxed(F.ctl + B2_FCTL_FREED_UNITS);
        publish(F, S);
    }

    B2_D static void seq_coords(const b2_pass_params& P, uint32_t seq, uint32_t& o0, uint32_t& o1, uint32_t& o2) {
        o0 = seq % P.nb[0]; seq /= P.nb[0];
        o1 = seq % P.nb[1]; seq /= P.nb[1];
        o2 = seq;
    }
    struct Where { uint32_t grp, o0, o1, o2; int64_t obase_in, obase_out; };
    B2_D static Where locate(const b2_fused_params& F, uint32_t kind, uint32_t unit, uint32_t tile) {
        constexpr uint64_t NN = (uint64_t)CA::N * (uint64_t)CB::N;              // points per sequence
        Where w;
        if (kind == TILE_A) {
            const uint32_t ga = F.A.G / CA::Q, sq = tile / ga;
            w.grp = tile % ga;
            seq_coords(F.A, unit * F.U + sq, w.o0, w.o1, w.o2);
            w.obase_in = (int64_t)w.o0 * F.A.in_bs[0] + (int64_t)w.o1 * F.A.in_bs[1] + (int64_t)w.o2 * F.A.in_bs[2];
            w.obase_out = (int64_t)(((uint64_t)(unit % F.R) * F.U + sq) * NN);
        } else {
            const uint32_t gb = F.B.G / CB::Q, sq = tile / gb;
            w.grp = tile % gb;
            seq_coords(F.B, unit * F.U + sq, w.o0, w.o1, w.o2);
            w.obase_in = (int64_t)(((uint64_t)(unit % F.R) * F.U + sq) * NN);
            w.obase_out = (int64_t)w.o0 * F.B.out_bs[0] + (int64_t)w.o1 * F.B.out_bs[1] + (int64_t)w.o2 * F.B.out_bs[2];
        }
        return w;
    }

    // ---- TMA: copy a tile into a shared-memory buffer; completion on the buffer's mbarrier ------------------------------------
    //   pass A: Q_A neighbouring columns = n1 row segments of Q_A*8 bytes, n2*8 bytes apart -> buf[p*Q_A + q]  (the
    //           interleaved-line layout the stages use, so they run in place); the 32 lanes of warp 0 share the row copies
    //   pass B: Q_B contiguous rows of the scratch = ONE contiguous block                   -> buf[q*n2 + p]  (dense; the
    //           first scatter moves it to the padded layout)
    // Called by every lane of warp 0 (emulation: by thread 0 alone); the arguments are taken from lane 0.
    B2_D static void warp_issue(const b2_fused_params& F, bool go, uint32_t kind, uint32_t unit, uint32_t tile, X* buf, uint64_t* bar) {
#if defined(__CUDA_ARCH__)
        const int lane = threadIdx.x & 31, nlanes = 32;
        go = __shfl_sync(0xffffffffu, go ? 1 : 0, 0) != 0;
        kind = __shfl_sync(0xffffffffu, kind, 0); unit = __shfl_sync(0xffffffffu, unit, 0); tile = __shfl_sync(0xffffffffu, tile, 0);
#else
        const int lane = 0, nlanes = 1;
#endif
        if (!go) return;
        const Where w = locate(F, kind, unit, tile);
        if (kind == TILE_A) {
            if (lane == 0) mbar_expect_tx(bar, BYTES_A);
#if defined(__CUDA_ARCH__)
            __syncwarp();
#endif
            const X* src = (const X*)F.A.in + w.obase_in + (int64_t)w.grp * CA::Q * F.A.in_gs;
            for (int p = lane; p < CA::N; p += nlanes)
                tma_load_1d(buf + (size_t)p * CA::Q, src + (int64_t)p * CB::N, (uint32_t)(CA::Q * sizeof(X)), bar);
        } else if (lane == 0) {
            mbar_expect_tx(bar, BYTES_B);
            const X* src = (const X*)F.B.in + w.obase_in + (int64_t)w.grp * CB::Q * CB::N;
            tma_load_1d(buf, src, BYTES_B, bar);
        }
    }
    // thread 0: hold back nothing, then wait until the tile may run (the slow path: its unit was not ready when looked at)
    B2_D static void wait_runnable(const b2_fused_params& F, Sched& S, uint32_t kind, uint32_t unit) {
        publish(F, S);
        after_publish(F, S);
        for (;;) {
            S.ready = fz_ld_relaxed(F.ctl + B2_FCTL_READY_UNITS);
            S.freed = fz_ld_relaxed(F.ctl + B2_FCTL_FREED_UNITS);
            if (runnable(F, S, kind, unit)) return;
            fz_sleep(100);
        }
    }

    // first-stage legs from the tile as TMA delivered it
    template <class E, class C>
    B2_D static void load_raw(X* x, const X* sm, int q, int t) {
        using Sch = typename C::Sch;
        constexpr int r = Sch::r(0), NB = C::N / r, BPT = E::template bpt<0>();
#pragma unroll
        for (int m = 0; m < BPT; ++m) {
            const int b = t + m * C::TPL;
            if (E::template guarded<0>() && b >= NB) continue;
#pragma unroll
            for (int k = 0; k < r; ++k) {
                const int p = b + k * NB;
                X a = B2_SMEM_LD(sm, C::LAYOUT == LAY_ELEM ? p * C::Q + q : q * C::N + p);
                x[m * r + k] = C::INV ? swp(a) : a;
            }
        }
    }

    // one tile of pass `C` from buffer `sm`.  `on_dead` runs (thread 0 .. 31) once every thread has read its last-stage legs
    template <class E, class C, class Dead>
    B2_D static void process(const b2_fused_params& F, const b2_pass_params& P, X* sm, const Where& w, Sched& S, bool discard, Dead on_dead) {
        using Sch = typename C::Sch;
        constexpr int NS = Sch::ns;
        const int tid = threadIdx.x;
        const X* __restrict__ lut = (const X*)P.lut;
        int ql, tl;
        E::template tmap<C::LMAP>(tid, ql, tl);
        {
            X x[E::template bpt<0>() * Sch::r(0)];
            load_raw<E, C>(x, sm, ql, tl);
            if (tid == 0) sched_issue(F, S);         // atomics + limit loads in flight behind the first butterflies
            E::template compute<0>(x, lut, tl);
            __syncthreads();                         // every raw read of the buffer is done: switch to the stage layout
            if constexpr (C::LAYOUT == LAY_LINE) { if (discard) E::discard_tile(P, w.obase_in, w.grp, tid); }   // pass B: the scratch lines are dead in L2 as well
            E::template store_smem<0>(x, sm, ql, tl);
        }
        __syncthreads();
        E::template middle<1>(sm, lut, tid);
        {
            constexpr int s = NS - 1;
            int qs, ts;
            E::template tmap<C::SMAP>(tid, qs, ts);
            const uint32_t gs = w.grp * C::Q + qs;
            X x[E::template bpt<s>() * Sch::r(s)];
            E::template load_smem<s>(x, sm, qs, ts);
            if constexpr (NBUF == 1) {
                fence_proxy_async_smem();
                __syncthreads();                     // the buffer is dead: refill it while the last stage runs
                on_dead();
            }
            E::template compute<s>(x, lut, ts);
            X* out_line = (X*)P.out + w.obase_out + (int64_t)gs * P.out_gs;
            E::template store_global<s>(x, out_line, P.out_es, ts, gs < P.G, P, E::twl(P, gs, w.o0, w.o1, w.o2), (uint32_t)qs);
        }
    }

    B2_D static void run(const b2_fused_params& F, unsigned char* smem_raw) {
        uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NBUF * TILE_BYTES);
        volatile uint32_t* mail = reinterpret_cast<volatile uint32_t*>(smem_raw + (size_t)NBUF * TILE_BYTES + 16);   // [slot][kind, unit, tile]
        const int tid = threadIdx.x;
#if defined(__CUDA_ARCH__)
        const bool w0 = tid < 32;                    // warp 0 issues the copies
#else
        const bool w0 = tid == 0;
#endif
        auto bufp = [&](uint32_t i) { return reinterpret_cast<X*>(smem_raw + (size_t)(i % NBUF) * TILE_BYTES); };
        Sched S;
        S.ready = S.freed = 0; S.pend_kind = S.done_kind = NONE; S.pend_unit = S.done_unit = 0; S.newT = S.dcount = 0;
        // thread 0: the tile being transformed and the one after it (whose copy may already be running)
        uint32_t ck = NONE, cu = 0, ct = 0, nk = NONE, nu = 0, nt = 0;
        bool cissued = false, nissued = false;
        if (tid == 0) {
            for (int i = 0; i < NBUF; ++i) mbar_init(&bars[i], 1);
            mbar_init_fence();
            const uint32_t a = fz_add(F.ctl + B2_FCTL_NEXT_A, 1u), b = fz_add(F.ctl + B2_FCTL_NEXT_A, 1u);
            ck = decode(F, a, cu, ct);
            nk = decode(F, b, nu, nt);
            mail[0] = ck; mail[1] = cu; mail[2] = ct;
            if (ck != NONE) wait_runnable(F, S, ck, cu);
        }
        __syncthreads();                             // mbarriers initialised, mailbox written
        if (w0) { warp_issue(F, ck != NONE, ck, cu, ct, bufp(0), &bars[0]); cissued = true; }
        for (uint32_t it = 0;; ++it) {
            const uint32_t slot = it & 1u;
            const uint32_t kind = mail[3 * slot], unit = mail[3 * slot + 1], tile = mail[3 * slot + 2];
            if (kind == NONE) break;
            X* sm = bufp(it);
            uint64_t* bar = &bars[it % NBUF];
            if (w0) {
                // this tile's copy could not be started ahead (its unit was not ready): wait for it now
                if (tid == 0 && !cissued) wait_runnable(F, S, ck, cu);
                warp_issue(F, !cissued, ck, cu, ct, sm, bar);
                cissued = true;
                if constexpr (NBUF == 2) {           // the other buffer is free: start the next tile's copy if its unit is ready
                    const bool go = nk != NONE && runnable(F, S, nk, nu);
                    warp_issue(F, go, nk, nu, nt, bufp(it + 1), &bars[(it + 1) % NBUF]);
                    nissued = go;
                }
            }
            mbar_wait(bar, (it / NBUF) & 1u);
            const Where w = locate(F, kind, unit, tile);
            auto dead = [&]() {                      // one buffer: refill it as soon as the last stage has its legs
                if (w0) {
                    const bool go = nk != NONE && runnable(F, S, nk, nu);
                    warp_issue(F, go, nk, nu, nt, sm, bar);
                    nissued = go;
                }
            };
            if (kind == TILE_A) process<EA, CA>(F, F.A, sm, w, S, false, dead);
            else process<EB, CB>(F, F.B, sm, w, S, !(F.B.aux_u1 & 1u), dead);
            if (tid == 0) {
                after_publish(F, S);
                S.pend_kind = kind; S.pend_unit = unit;          // published behind the next tile's first stage (or on the wait / exit path)
                volatile uint32_t* m = mail + 3 * (slot ^ 1);
                m[0] = nk; m[1] = nu; m[2] = nt;
                ck = nk; cu = nu; ct = nt; cissued = nissued;
                nk = decode(F, S.newT, nu, nt);                  // the ticket fetched behind this tile's first stage
                nissued = false;
            }
            fence_proxy_async_smem();
            __syncthreads();      // stores issued; this buffer may be refilled; the next mailbox slot is visible
        }
        if (tid == 0) {
            publish(F, S);
            after_publish(F, S);
        }
    }
};

#if defined(__CUDACC__)
template <class CA, class CB, int NBUF>
__global__ void __launch_bounds__(CA::THREADS, Fused4<CA, CB, NBUF>::MINB) fused4_kernel(const __grid_constant__ b2_fused_params F) {
    extern __shared__ __align__(128) unsigned char b2_smem_fused[];
    Fused4<CA, CB, NBUF>::run(F, b2_smem_fused);
}
// control block: ticket counters, unit prefixes and per-unit done counters all start at zero
template <int DUMMY = 0>
__global__ void fused4_init_kernel(uint32_t* ctl, uint32_t words) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < words; i += gridDim.x * blockDim.x) ctl[i] = 0u;
}
#endif

}  // namespace b200fft
