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
//     done.  Pass B has priority, so the ring drains as fast as it fills.  Two counting semaphores (AVAIL_A / AVAIL_B)
//     make a claim two atomics and guarantee that a claimed tile never has to wait; the claim for the NEXT tile is
//     made while the current tile's first loads are in flight (Engine::run_at hook), so it costs nothing;
//   * pass B reads the scratch with ld.global.cg (it was written by other SMs in this launch) and, once the legs are in
//     registers, drops the lines from L2 (discard.global.L2) so that dead scratch is never written back to HBM.
// Both passes run the very same stage code as the stand-alone kernels (Engine<C>::run_at), so results are identical
// to the two-launch plan bit for bit.
#pragma once
#include "stockham.cuh"

namespace b200fft {

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
B2_D uint32_t fz_add(uint32_t* p, uint32_t v) {       // acq_rel fetch-add
    uint32_t o;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(o) : "l"(p), "r"(v) : "memory");
    return o;
}
B2_D uint32_t fz_cas(uint32_t* p, uint32_t cmp, uint32_t val) {
    uint32_t o;
    asm volatile("atom.acq_rel.gpu.global.cas.b32 %0, [%1], %2, %3;" : "=r"(o) : "l"(p), "r"(cmp), "r"(val) : "memory");
    return o;
}
B2_D void fz_sleep(unsigned ns) { __nanosleep(ns); }
#else
// CPU emulation: one CTA at a time, only thread 0 touches the control words
B2_D uint32_t fz_ld_acquire(const uint32_t* p) { return *p; }
B2_D uint32_t fz_ld_relaxed(const uint32_t* p) { return *p; }
B2_D uint32_t fz_add(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o + v; return o; }
B2_D uint32_t fz_cas(uint32_t* p, uint32_t cmp, uint32_t val) { uint32_t o = *p; if (o == cmp) *p = val; return o; }
B2_D void fz_sleep(unsigned) {}
#endif

template <class CA, class CB>
struct Fused4 {
    using EA = Engine<CA, 0>;
    using EB = Engine<CB, XF_LDCG | XF_DISCARD>;
    static_assert(CA::THREADS == CB::THREADS, "both passes run in the same CTA shape");
    static_assert(CA::LAYOUT == LAY_ELEM && CB::LAYOUT == LAY_LINE, "pass A: interleaved columns, pass B: contiguous rows");
    static constexpr int THREADS = CA::THREADS;
    static constexpr int MINB = CA::MINB < CB::MINB ? CA::MINB : CB::MINB;
    static constexpr int TILE_BYTES = ((CA::SMEM_BYTES > CB::SMEM_BYTES ? CA::SMEM_BYTES : CB::SMEM_BYTES) + 15) / 16 * 16;
    static constexpr int SMEM_BYTES = TILE_BYTES + 16;     // + mailbox (kind, ticket)
    enum { NONE = 0, TILE_A = 1, TILE_B = 2 };

    // ---- scheduler (thread 0 only) ------------------------------------------------------------------------------
    B2_D static bool try_sem(uint32_t* sem) {
        if ((int32_t)fz_ld_relaxed(sem) <= 0) return false;
        if ((int32_t)fz_add(sem, (uint32_t)-1) > 0) return true;
        fz_add(sem, 1u);
        return false;
    }
    B2_D static void claim_try(const b2_fused_params& F, uint32_t& kind, uint32_t& ticket) {
        uint32_t* c = F.ctl;
        kind = NONE; ticket = 0;
        if (try_sem(c + B2_FCTL_AVAIL_B)) { kind = TILE_B; ticket = fz_add(c + B2_FCTL_NEXT_B, 1u); return; }
        if (try_sem(c + B2_FCTL_AVAIL_A)) { kind = TILE_A; ticket = fz_add(c + B2_FCTL_NEXT_A, 1u); return; }
    }
    B2_D static void claim_blocking(const b2_fused_params& F, uint32_t& kind, uint32_t& ticket) {
        const uint32_t totalA = F.NU * F.TA, totalB = F.NU * F.TB;
        for (;;) {
            claim_try(F, kind, ticket);
            if (kind != NONE) return;
            // every tile of both passes has been handed out: nothing left for this CTA
            if (fz_ld_relaxed(F.ctl + B2_FCTL_NEXT_A) >= totalA && fz_ld_relaxed(F.ctl + B2_FCTL_NEXT_B) >= totalB) return;
            fz_sleep(256);
        }
    }
    // a unit finished a pass: advance the in-order prefix and release the tiles that depend on it
    B2_D static void advance(const b2_fused_params& F, uint32_t prefix_word, const uint32_t* done, uint32_t per_unit,
                             uint32_t sem_word, uint32_t release, uint32_t skew) {
        uint32_t* c = F.ctl;
        for (;;) {
            const uint32_t p = fz_ld_acquire(c + prefix_word);
            if (p >= F.NU) return;
            if (fz_ld_acquire(done + p) != per_unit) return;
            if (fz_cas(c + prefix_word, p, p + 1) == p && p + skew < F.NU) fz_add(c + sem_word, release);
        }
    }
    B2_D static void finish(const b2_fused_params& F, uint32_t kind, uint32_t unit) {
        uint32_t* doneA = F.ctl + B2_FCTL_WORDS;
        uint32_t* doneB = doneA + F.NU;
        if (kind == TILE_A) {
            if (fz_add(doneA + unit, 1u) + 1 == F.TA)
                advance(F, B2_FCTL_READY_UNITS, doneA, F.TA, B2_FCTL_AVAIL_B, F.TB, 0);
        } else {
            // unit p's slot is written next by unit p + R: release that unit's A tiles (if it exists)
            if (fz_add(doneB + unit, 1u) + 1 == F.TB)
                advance(F, B2_FCTL_FREED_UNITS, doneB, F.TB, B2_FCTL_AVAIL_A, F.TA, F.R);
        }
    }

    struct Prefetch {
        const b2_fused_params* F;
        volatile uint32_t* mail;
        B2_D void operator()() const {
            if (threadIdx.x == 0) {
                uint32_t k, t;
                claim_try(*F, k, t);
                mail[0] = k; mail[1] = t;
            }
        }
    };

    B2_D static void seq_coords(const b2_pass_params& P, uint32_t seq, uint32_t& o0, uint32_t& o1, uint32_t& o2) {
        o0 = seq % P.nb[0]; seq /= P.nb[0];
        o1 = seq % P.nb[1]; seq /= P.nb[1];
        o2 = seq;
    }

    B2_D static void run(const b2_fused_params& F, unsigned char* smem_raw) {
        volatile uint32_t* mail = reinterpret_cast<volatile uint32_t*>(smem_raw + TILE_BYTES);
        const int tid = threadIdx.x;
        const uint64_t NN = (uint64_t)CA::N * (uint64_t)CB::N;              // points per sequence
        const uint32_t ga = (F.A.G + CA::Q - 1) / CA::Q, gb = (F.B.G + CB::Q - 1) / CB::Q;   // tiles per sequence
        if (tid == 0) {
            uint32_t k, t;
            claim_blocking(F, k, t);
            mail[0] = k; mail[1] = t;
        }
        __syncthreads();
        for (;;) {
            const uint32_t kind = mail[0], ticket = mail[1];
            __syncthreads();                     // mailbox read by everyone; the previous tile's shared memory is dead
            if (kind == NONE) break;
            const Prefetch pf{&F, mail};
            uint32_t unit;
            if (kind == TILE_A) {
                unit = ticket / F.TA;
                const uint32_t idx = ticket % F.TA, sq = idx / ga, grp = idx % ga;
                uint32_t o0, o1, o2;
                seq_coords(F.A, unit * F.U + sq, o0, o1, o2);
                const int64_t obase_in = (int64_t)o0 * F.A.in_bs[0] + (int64_t)o1 * F.A.in_bs[1] + (int64_t)o2 * F.A.in_bs[2];
                const int64_t obase_out = (int64_t)(((uint64_t)(unit % F.R) * F.U + sq) * NN);
                EA::run_at(F.A, smem_raw, grp, o0, o1, o2, obase_in, obase_out, pf);
            } else {
                unit = ticket / F.TB;
                const uint32_t idx = ticket % F.TB, sq = idx / gb, grp = idx % gb;
                uint32_t o0, o1, o2;
                seq_coords(F.B, unit * F.U + sq, o0, o1, o2);
                const int64_t obase_in = (int64_t)(((uint64_t)(unit % F.R) * F.U + sq) * NN);
                const int64_t obase_out = (int64_t)o0 * F.B.out_bs[0] + (int64_t)o1 * F.B.out_bs[1] + (int64_t)o2 * F.B.out_bs[2];
                EB::run_at(F.B, smem_raw, grp, o0, o1, o2, obase_in, obase_out, pf);
            }
            __syncthreads();                     // every store of this tile has been issued by its thread
            if (tid == 0) {
                finish(F, kind, unit);           // release: the tile's stores (and discards) are visible before the count moves
                if (mail[0] == NONE) {           // nothing was claimable when the prefetch looked: wait for work (or the end)
                    uint32_t k, t;
                    claim_blocking(F, k, t);
                    mail[0] = k; mail[1] = t;
                }
            }
            __syncthreads();
        }
    }
};

#if defined(__CUDACC__)
template <class CA, class CB>
__global__ void __launch_bounds__(CA::THREADS, Fused4<CA, CB>::MINB) fused4_kernel(const __grid_constant__ b2_fused_params F) {
    extern __shared__ __align__(16) unsigned char b2_smem_fused[];
    Fused4<CA, CB>::run(F, b2_smem_fused);
}
// control block: all zero except the A semaphore, which starts with the tiles of the first R units
template <int DUMMY = 0>
__global__ void fused4_init_kernel(uint32_t* ctl, uint32_t words, uint32_t avail_a) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < words; i += gridDim.x * blockDim.x)
        ctl[i] = (i == B2_FCTL_AVAIL_A) ? avail_a : 0u;
}
#endif

}  // namespace b200fft
