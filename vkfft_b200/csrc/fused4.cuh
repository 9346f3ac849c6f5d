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

template <class CA, class CB>
struct Fused4 {
    // strides known at compile time: pass A walks columns of an n1 x n2 matrix (element stride n2 on both sides), pass B
    // stores transposed (element stride n1)
    using EA = Engine<CA, 0, CB::N, CB::N>;
    using EB = Engine<CB, XF_LDCG | XF_DISCARD, 0, CA::N>;
    static_assert(CA::THREADS == CB::THREADS, "both passes run in the same CTA shape");
    static_assert(CA::LAYOUT == LAY_ELEM && CB::LAYOUT == LAY_LINE, "pass A: interleaved columns, pass B: contiguous rows");
    static constexpr int THREADS = CA::THREADS;
    static constexpr int MINB = CA::MINB < CB::MINB ? CA::MINB : CB::MINB;
    static constexpr int TILE_BYTES = ((CA::SMEM_BYTES > CB::SMEM_BYTES ? CA::SMEM_BYTES : CB::SMEM_BYTES) + 15) / 16 * 16;
    static constexpr int SMEM_BYTES = TILE_BYTES + 32;     // + mailbox: two alternating slots of (kind, unit, tile, must-wait)
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
        uint32_t next;                 // ticket of the following tile (fetched one tile ahead)
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

    struct Hook {
        const b2_fused_params* F;
        Sched* S;
        B2_D void operator()() const {
            if (threadIdx.x == 0) {
                S->newT = fz_add(F->ctl + B2_FCTL_NEXT_A, 1u);
                S->ready = fz_ld_relaxed(F->ctl + B2_FCTL_READY_UNITS);
                S->freed = fz_ld_relaxed(F->ctl + B2_FCTL_FREED_UNITS);
                publish(*F, *S);
            }
        }
    };

    B2_D static void seq_coords(const b2_pass_params& P, uint32_t seq, uint32_t& o0, uint32_t& o1, uint32_t& o2) {
        o0 = seq % P.nb[0]; seq /= P.nb[0];
        o1 = seq % P.nb[1]; seq /= P.nb[1];
        o2 = seq;
    }

    B2_D static void run(const b2_fused_params& F, unsigned char* smem_raw) {
        volatile uint32_t* mail = reinterpret_cast<volatile uint32_t*>(smem_raw + TILE_BYTES);    // [slot][kind, unit, tile, wait]
        const int tid = threadIdx.x;
        const uint64_t NN = (uint64_t)CA::N * (uint64_t)CB::N;              // points per sequence
        const uint32_t ga = (F.A.G + CA::Q - 1) / CA::Q, gb = (F.B.G + CB::Q - 1) / CB::Q;   // tiles per sequence
        Sched S;
        S.next = 0; S.ready = S.freed = 0; S.pend_kind = S.done_kind = NONE; S.pend_unit = S.done_unit = 0; S.newT = S.dcount = 0;
        if (tid == 0) {
            const uint32_t t0 = fz_add(F.ctl + B2_FCTL_NEXT_A, 1u);
            S.next = fz_add(F.ctl + B2_FCTL_NEXT_A, 1u);
            uint32_t u, t;
            const uint32_t k = decode(F, t0, u, t);
            mail[0] = k; mail[1] = u; mail[2] = t; mail[3] = (k != NONE && !runnable(F, S, k, u)) ? 1u : 0u;
        }
        __syncthreads();
        uint32_t slot = 0;
        for (;;) {
            const uint32_t kind = mail[4 * slot], unit = mail[4 * slot + 1], tile = mail[4 * slot + 2], wait = mail[4 * slot + 3];
            if (kind == NONE) break;
            if (wait) {
                // the tile's unit is not ready yet.  Every store of the previous tile has been issued (barrier): publish it
                // first -- others may be waiting for exactly that -- then poll the two limits
                if (tid == 0) {
                    publish(F, S);
                    after_publish(F, S);
                    for (;;) {
                        S.ready = fz_ld_relaxed(F.ctl + B2_FCTL_READY_UNITS);
                        S.freed = fz_ld_relaxed(F.ctl + B2_FCTL_FREED_UNITS);
                        if (runnable(F, S, kind, unit)) break;
                        fz_sleep(100);
                    }
                }
                __syncthreads();
            }
            const Hook hk{&F, &S};
            if (kind == TILE_A) {
                const uint32_t sq = tile / ga, grp = tile % ga;
                uint32_t o0, o1, o2;
                seq_coords(F.A, unit * F.U + sq, o0, o1, o2);
                const int64_t obase_in = (int64_t)o0 * F.A.in_bs[0] + (int64_t)o1 * F.A.in_bs[1] + (int64_t)o2 * F.A.in_bs[2];
                const int64_t obase_out = (int64_t)(((uint64_t)(unit % F.R) * F.U + sq) * NN);
                EA::run_at(F.A, smem_raw, grp, o0, o1, o2, obase_in, obase_out, hk);
            } else {
                const uint32_t sq = tile / gb, grp = tile % gb;
                uint32_t o0, o1, o2;
                seq_coords(F.B, unit * F.U + sq, o0, o1, o2);
                const int64_t obase_in = (int64_t)(((uint64_t)(unit % F.R) * F.U + sq) * NN);
                const int64_t obase_out = (int64_t)o0 * F.B.out_bs[0] + (int64_t)o1 * F.B.out_bs[1] + (int64_t)o2 * F.B.out_bs[2];
                EB::run_at(F.B, smem_raw, grp, o0, o1, o2, obase_in, obase_out, hk);
            }
            if (tid == 0) {
                // the atomics issued behind the first loads have long returned
                after_publish(F, S);
                S.pend_kind = kind; S.pend_unit = unit;      // published behind the next tile's loads (or on the wait / exit path)
                uint32_t u, t;
                const uint32_t k = decode(F, S.next, u, t);
                S.next = S.newT;
                volatile uint32_t* m = mail + 4 * (slot ^ 1);
                m[0] = k; m[1] = u; m[2] = t; m[3] = (k != NONE && !runnable(F, S, k, u)) ? 1u : 0u;
            }
            __syncthreads();      // every store of this tile has been issued; its shared memory is dead; the next mailbox slot is visible
            slot ^= 1;
        }
        if (tid == 0) {
            publish(F, S);
            after_publish(F, S);
        }
    }
};

#if defined(__CUDACC__)
template <class CA, class CB>
__global__ void __launch_bounds__(CA::THREADS, Fused4<CA, CB>::MINB) fused4_kernel(const __grid_constant__ b2_fused_params F) {
    extern __shared__ __align__(16) unsigned char b2_smem_fused[];
    Fused4<CA, CB>::run(F, b2_smem_fused);
}
// control block: ticket counters, unit prefixes and per-unit done counters all start at zero
template <int DUMMY = 0>
__global__ void fused4_init_kernel(uint32_t* ctl, uint32_t words) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < words; i += gridDim.x * blockDim.x) ctl[i] = 0u;
}
#endif

}  // namespace b200fft
