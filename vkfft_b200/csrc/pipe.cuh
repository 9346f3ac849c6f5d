// Persistent, TMA-fed variant of the specialised Stockham kernels for contiguous input lines.
//
// One CTA stays resident and walks tiles (Q lines each) with a ring of NBUF shared-memory tile buffers:
//   * an elected thread issues `cp.async.bulk` (TMA, 1-D bulk copy, SASS UBLKCP) for tiles NBUF-1 ahead; completion
//     is signalled on an mbarrier per buffer (`complete_tx::bytes`), so HBM reads for the next tiles are in flight
//     while the current tile is being transformed -- latency hiding no longer depends on how many CTAs fit on an SM;
//   * the compute threads wait on the buffer's mbarrier, read the first-stage legs from the raw (dense) tile, run the
//     same radix stages in place in that buffer (padded layout from the first scatter on), and store the last stage
//     straight from registers to HBM (contiguous or transposed, as in stockham.cuh);
//   * after the last shared-memory read of a buffer (+ fence.proxy.async + barrier) it is refilled by TMA.
// The reference has no asynchronous copy path at all: every VkFFT_main does
// load -> compute -> store with plain loads.
#pragma once
#include <string.h>
#include "stockham.cuh"

namespace b200fft {

#if defined(__CUDA_ARCH__)
B2_D uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
B2_D void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
B2_D void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
B2_D void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
B2_D void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
B2_D void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
B2_D void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#else
// CPU emulation: the "TMA" is a synchronous copy done by the issuing thread; waiting is a block barrier
B2_D void mbar_init(uint64_t*, uint32_t) {}
B2_D void mbar_init_fence() {}
B2_D void mbar_expect_tx(uint64_t*, uint32_t) {}
B2_D void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t*) { memcpy(dst, src, bytes); }
B2_D void mbar_wait(uint64_t*, uint32_t) { __syncthreads(); }
B2_D void fence_proxy_async_smem() {}
#endif

template <class C, int NBUF>
struct PipeEngine : Engine<C> {
    using E = Engine<C>;
    using T = typename C::T;
    using X = cpx<T>;
    using Sch = typename C::Sch;
    static constexpr int N = C::N, TPL = C::TPL, Q = C::Q, V = C::V, NS = Sch::ns;
    static_assert(NS >= 2, "the pipelined kernel needs a shared-memory exchange");
    static_assert(C::LAYOUT == LAY_LINE && C::LMAP == MAP_TFAST && C::IN_UNIT, "contiguous input lines only");
    static constexpr int TILE_BYTES = ((C::SMEM_BYTES + 127) / 128) * 128;
    static constexpr int SMEM_BYTES = NBUF * TILE_BYTES + 8 * NBUF + 16;

    // first-stage legs from the raw tile as TMA delivered it: line q at q*N, no padding
    template <int s>
    B2_D static void load_raw(X* x, const X* sm, int q, int t) {
        constexpr int r = Sch::r(s), NB = E::template nbut<s>(), BPT = E::template bpt<s>();
#pragma unroll
        for (int m = 0; m < BPT; ++m) {
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const int b = V * (t + m * TPL) + v;
                if (E::template guarded<s>() && b >= NB) continue;
#pragma unroll
                for (int k = 0; k < r; ++k) {
                    X a = B2_SMEM_LD(sm, q * N + b + k * NB);
                    x[(m * V + v) * r + k] = C::INV ? swp(a) : a;
                }
            }
        }
    }

    struct TileCoord { uint32_t grp, o0, o1, o2; };
    B2_D static TileCoord decode(const b2_pass_params& P, uint32_t tile, uint32_t ngrp) {
        TileCoord c;
        uint32_t rest = tile;
        c.grp = rest % ngrp; rest /= ngrp;
        c.o0 = rest % P.nb[0]; rest /= P.nb[0];
        c.o1 = rest % P.nb[1]; rest /= P.nb[1];
        c.o2 = rest;
        return c;
    }

    B2_D static void issue(const b2_pass_params& P, uint32_t tile, uint32_t ngrp, X* buf, uint64_t* bar) {
        const TileCoord c = decode(P, tile, ngrp);
        const int64_t obase = (int64_t)c.o0 * P.in_bs[0] + (int64_t)c.o1 * P.in_bs[1] + (int64_t)c.o2 * P.in_bs[2];
        const uint32_t g0 = c.grp * Q;
        const uint32_t nvalid = (P.G - g0) < (uint32_t)Q ? (P.G - g0) : (uint32_t)Q;
        const X* src = (const X*)P.in + obase + (int64_t)g0 * P.in_gs;
        const uint32_t line_bytes = (uint32_t)(N * sizeof(X));
        mbar_expect_tx(bar, nvalid * line_bytes);
        if (P.in_gs == (int64_t)N) {
            tma_load_1d(buf, src, nvalid * line_bytes, bar);
        } else {
            for (uint32_t q = 0; q < nvalid; ++q) tma_load_1d(buf + q * N, src + (int64_t)q * P.in_gs, line_bytes, bar);
        }
    }

    B2_D static void run(const b2_pass_params& P, unsigned char* smem_raw) {
        const int tid = threadIdx.x;
        const uint32_t ngrp = (P.G + Q - 1) / Q;
        const uint32_t ntiles = ngrp * P.nb[0] * P.nb[1] * P.nb[2];
        const uint32_t stride = gridDim.x;
        uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NBUF * TILE_BYTES);
        const X* __restrict__ lut = (const X*)P.lut;
        if (tid == 0) {
            for (int i = 0; i < NBUF; ++i) mbar_init(&bars[i], 1);
            mbar_init_fence();
        }
        __syncthreads();
        if (tid == 0) {
            for (int i = 0; i < NBUF; ++i) {
                const uint32_t tile = blockIdx.x + (uint32_t)i * stride;
                if (tile < ntiles) issue(P, tile, ngrp, reinterpret_cast<X*>(smem_raw + (size_t)i * TILE_BYTES), &bars[i]);
            }
        }
        uint32_t it = 0;
        for (uint32_t tile = blockIdx.x; tile < ntiles; tile += stride, ++it) {
            const int b = (int)(it % NBUF);
            X* sm = reinterpret_cast<X*>(smem_raw + (size_t)b * TILE_BYTES);
            mbar_wait(&bars[b], (it / NBUF) & 1u);
            const TileCoord c = decode(P, tile, ngrp);
            const int64_t obase_out = (int64_t)c.o0 * P.out_bs[0] + (int64_t)c.o1 * P.out_bs[1] + (int64_t)c.o2 * P.out_bs[2];
            int ql, tl;
            E::template tmap<C::LMAP>(tid, ql, tl);
            {
                X x[E::template bpt<0>() * V * Sch::r(0)];
                load_raw<0>(x, sm, ql, tl);
                E::template compute<0>(x, lut, tl);
                __syncthreads();                       // every raw read of this buffer is done: switch to the padded layout
                E::template store_smem<0>(x, sm, ql, tl);
            }
            __syncthreads();
            E::template middle<1>(sm, lut, tid);
            {
                constexpr int s = NS - 1;
                int qs, ts;
                E::template tmap<C::SMAP>(tid, qs, ts);
                const uint32_t gs = c.grp * Q + qs;
                X x[E::template bpt<s>() * V * Sch::r(s)];
                E::template load_smem<s>(x, sm, qs, ts);
                // this buffer is dead once every thread has its last-stage legs: hand it back to the TMA
                fence_proxy_async_smem();
                __syncthreads();
                if (tid == 0) {
                    const uint32_t next = tile + (uint32_t)NBUF * stride;
                    if (next < ntiles) issue(P, next, ngrp, sm, &bars[b]);
                }
                E::template compute<s>(x, lut, ts);
                X* out_line = (X*)P.out + obase_out + (int64_t)gs * P.out_gs;
                E::template store_global<s>(x, out_line, P.out_es, ts, gs < P.G, P, E::twl(P, gs, c.o0, c.o1, c.o2), (uint32_t)qs);
            }
        }
    }
};

#if defined(__CUDACC__)
template <class C, int NBUF>
__global__ void __launch_bounds__(C::THREADS, C::MINB) stockham_pipe_kernel(const __grid_constant__ b2_pass_params P) {
    extern __shared__ __align__(128) unsigned char b2_smem_pipe[];
    PipeEngine<C, NBUF>::run(P, b2_smem_pipe);
}
#endif

}  // namespace b200fft
