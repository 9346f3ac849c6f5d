// Runtime-scheduled Stockham kernel: any length whose radices are in {2..16} (so every 2/3/5/7/11/13-smooth
// N), any line addressing, with the real-data transforms fused into its load and store phases.
//
// One CTA owns Q lines of length n in shared memory (two buffers, ping-pong):
//     load phase  : HBM -> smem   (B2_IO_* decides how a line is assembled: complex copy, zero padding,
//                                  Bluestein chirp, C2R pre-pass, DCT permutation / phase merge ...)
//     stages      : for each radix r (runtime list) the butterflies of the stage, dispatched once per stage
//                   to the compile-time dft<r>; twiddles from the per-stage LUT (same layout as the fast path)
//     store phase : smem -> HBM   (complex copy / truncation, four-step phase, Bluestein filter,
//                                  R2C Hermitian post-pass, DCT split + phase ...)
// It trades two extra shared-memory passes for generality; the power-of-two C2C hot path uses the fully
// specialised kernels in stockham.cuh instead.  The reference reaches the same coverage by generating a new
// kernel text per plan (shaderGen_FFT, vkFFT_FFT.h:48; R2C :vkFFT_R2C.h:178,450; R2R vkFFT_R2R.h;
// Bluestein vkFFT_Bluestein.h:32,201).
#pragma once
#include "pass_params.h"
#include "radix.cuh"
#include "stockham.cuh"

namespace b200fft {

// RMAX = largest radix this instantiation can dispatch to (8, 11 or 16): the register allocation of the kernel is the
// maximum over all butterflies it contains, so schedules made of small radices get a leaner kernel and more CTAs per SM.
template <typename T, int RMAX = 16>
struct Generic {
    using X = cpx<T>;

    B2_D static int pad(int p) { return p + (p >> (sizeof(T) == 4 ? 4 : 3)); }

    // ---- one radix stage over the CTA's lines: src -> dst (both smem) ---------------------------------------
    template <int R>
    B2_D static void stage(const X* src, X* dst, int n, int S, const X* __restrict__ lut, int q, int t, int tpl,
                           int ls) {
        const int nb = n / R;
        const X* s = src + q * ls;
        X* d = dst + q * ls;
        for (int b = t; b < nb; b += tpl) {
            X x[R];
#pragma unroll
            for (int k = 0; k < R; ++k) x[k] = B2_SMEM_LD(s, pad(b + k * nb));
            const int j = b % S;
            if (S > 1) {
#pragma unroll
                for (int k = 1; k < R; ++k) x[k] = x[k] * ld_lut(lut + (k - 1) * S + j);
            }
            dft<R, T>(x);
            const int base = (b - j) * R + j;
#pragma unroll
            for (int k = 0; k < R; ++k) B2_SMEM_ST(d, pad(base + k * S), x[k]);
        }
    }

    // ---- first stage with its legs straight from HBM and/or last stage with its outputs straight to HBM ------------------
    // (plain complex lines only).  The separate load / store phases cost one shared-memory write + read each per element; here
    // the first butterflies read x[b + k n/R] from the line in global memory (consecutive threads, consecutive b: coalesced for
    // every k) and the last butterflies write y[b + k S] (S = n/R, again consecutive b), with the load / store operators of
    // load_line / store_line applied in registers.  FIN / FOUT: this stage is the first / the last one of the schedule.
    template <int R, bool FIN, bool FOUT>
    B2_D static void stage_io(const b2_pass_params& P, const X* src, X* dst, int S, const X* __restrict__ lut, int q, int t, int tpl,
                              int ls, int64_t in_off, int64_t out_off, uint32_t twline, bool valid) {
        const int n = (int)P.n, nb = n / R;
        const bool inv = P.inverse != 0, inner = P.inner_inverse != 0;
        const X zero = mk<T>(T(0), T(0));
        const X* s = src + q * ls;
        X* d = dst + q * ls;
        const X* in = (const X*)P.in + in_off;
        X* out = (X*)P.out + out_off;
        const T sc = (T)P.scale;
        const bool do_scale = (P.ops & B2_OP_SCALE) != 0;
        const int lim = (int)P.out_len < n ? (int)P.out_len : n;
        if (FOUT && !valid) return;
        for (int b = t; b < nb; b += tpl) {
            X x[R];
            if constexpr (FIN) {
#pragma unroll
                for (int k = 0; k < R; ++k) {
                    const int p = b + k * nb;
                    X v = zero;
                    if (valid && p < (int)P.in_len) {
                        v = in[(int64_t)p * P.in_es];
                        if (inv) v = swp(v);
                        if (P.ops & B2_OP_MUL_IN) v = v * ld_lut((const X*)P.aux0 + p);
                        if (inner) v = swp(v);
                    }
                    x[k] = v;
                }
            } else {
#pragma unroll
                for (int k = 0; k < R; ++k) x[k] = B2_SMEM_LD(s, pad(b + k * nb));
            }
            const int j = b % S;
            if (!FIN && S > 1) {
#pragma unroll
                for (int k = 1; k < R; ++k) x[k] = x[k] * ld_lut(lut + (k - 1) * S + j);
            }
            dft<R, T>(x);
            const int base = (b - j) * R + j;
            if constexpr (FOUT) {
#pragma unroll
                for (int k = 0; k < R; ++k) {
                    const int p = base + k * S;
                    if (p >= lim) continue;
                    X v = x[k];
                    if (inner) v = swp(v);
                    if (P.ops & B2_OP_MUL_OUT) v = v * ld_lut((const X*)P.aux1 + p);
                    if (P.ops & B2_OP_TWIDDLE_OUT) {
                        const uint64_t e = (uint64_t)twline * (uint64_t)p;
                        v = v * twiddle2<T>((const X*)P.tw_hi, (const X*)P.tw_lo, P.tw_shift, e);
                    }
                    if (do_scale) v = v * sc;
                    if (inv) v = swp(v);
                    out[(int64_t)p * P.out_es] = v;
                }
            } else {
#pragma unroll
                for (int k = 0; k < R; ++k) B2_SMEM_ST(d, pad(base + k * S), x[k]);
            }
        }
    }
    template <bool FIN, bool FOUT>
    B2_D static void run_stage_io(int r, const b2_pass_params& P, const X* src, X* dst, int S, const X* lut, int q, int t, int tpl, int ls,
                                  int64_t in_off, int64_t out_off, uint32_t twline, bool valid) {
#define B2_GEN_IO(RR) case RR: stage_io<RR, FIN, FOUT>(P, src, dst, S, lut, q, t, tpl, ls, in_off, out_off, twline, valid); break;
        switch (r) {
            B2_GEN_IO(2) B2_GEN_IO(3) B2_GEN_IO(4) B2_GEN_IO(5) B2_GEN_IO(6) B2_GEN_IO(7) B2_GEN_IO(8)
            default:
                if constexpr (RMAX > 8) {
                    switch (r) {
                        B2_GEN_IO(9) B2_GEN_IO(10) B2_GEN_IO(11)
                        default:
                            if constexpr (RMAX > 11) {
                                switch (r) {
                                    B2_GEN_IO(12) B2_GEN_IO(13) B2_GEN_IO(14) B2_GEN_IO(15) B2_GEN_IO(16)
                                    default: break;
                                }
                            }
                            break;
                    }
                }
                break;
        }
#undef B2_GEN_IO
    }

    // ---- Rader stage for a prime radix p > 16 ("mult" form: the length-(p-1) cyclic convolution is evaluated
    //      directly, vkFFT_RaderKernels.h:1278).  X_0 = sum of the legs;  X_{g^-q} = x_0 + sum_m a_m b_{(q-m) mod (p-1)},
    //      a_m = leg_{g^m}.  The legs are twiddled in place first (phase 1), then every thread produces outputs (phase 2).
    B2_D static void rader_stage(int p, X* src, X* dst, int n, int S, const X* __restrict__ lut, const X* __restrict__ rt,
                                 int q, int t, int tpl, int ls) {
        const int nb = n / p, pm1 = p - 1;
        X* s = src + q * ls;
        X* d = dst + q * ls;
        if (S > 1) {
            for (int idx = t; idx < nb * pm1; idx += tpl) {
                const int b = idx % nb, i = 1 + idx / nb;
                const int a = pad(b + i * nb);
                B2_SMEM_ST(s, a, B2_SMEM_LD(s, a) * ld_lut(lut + (i - 1) * S + (b % S)));
            }
        }
        __syncthreads();
        const X* bt = rt;             // b_m
        const X* perm = rt + pm1;     // (g^m, g^-m)
        for (int idx = t; idx < nb * p; idx += tpl) {
            const int b = idx % nb, qq = idx / nb;
            const int j = b % S, base = (b - j) * p + j;
            const X x0 = B2_SMEM_LD(s, pad(b));
            X acc = x0;
            int ko = 0;
            if (qq == pm1) {
                for (int i = 1; i < p; ++i) acc = acc + B2_SMEM_LD(s, pad(b + i * nb));
            } else {
                ko = (int)ld_lut(perm + qq).y;
                int bi = qq;                                   // g^(m-q) = g^-(q-m): index (q - m) mod (p-1), m = 0
                for (int m = 0; m < pm1; ++m) {
                    const int leg = (int)ld_lut(perm + m).x;
                    acc = acc + B2_SMEM_LD(s, pad(b + leg * nb)) * ld_lut(bt + bi);
                    bi = (bi == 0) ? pm1 - 1 : bi - 1;
                }
            }
            B2_SMEM_ST(d, pad(base + ko * S), acc);
        }
    }

    B2_D static void run_stage(int r, const X* src, X* dst, int n, int S, const X* lut, int q, int t, int tpl, int ls) {
        switch (r) {
            case 2: stage<2>(src, dst, n, S, lut, q, t, tpl, ls); break;
            case 3: stage<3>(src, dst, n, S, lut, q, t, tpl, ls); break;
            case 4: stage<4>(src, dst, n, S, lut, q, t, tpl, ls); break;
            case 5: stage<5>(src, dst, n, S, lut, q, t, tpl, ls); break;
            case 6: stage<6>(src, dst, n, S, lut, q, t, tpl, ls); break;
            case 7: stage<7>(src, dst, n, S, lut, q, t, tpl, ls); break;
            case 8: stage<8>(src, dst, n, S, lut, q, t, tpl, ls); break;
            default:
                if constexpr (RMAX > 8) {
                    switch (r) {
                        case 9: stage<9>(src, dst, n, S, lut, q, t, tpl, ls); break;
                        case 10: stage<10>(src, dst, n, S, lut, q, t, tpl, ls); break;
                        case 11: stage<11>(src, dst, n, S, lut, q, t, tpl, ls); break;
                        default:
                            if constexpr (RMAX > 11) {
                                switch (r) {
                                    case 12: stage<12>(src, dst, n, S, lut, q, t, tpl, ls); break;
                                    case 13: stage<13>(src, dst, n, S, lut, q, t, tpl, ls); break;
                                    case 14: stage<14>(src, dst, n, S, lut, q, t, tpl, ls); break;
                                    case 15: stage<15>(src, dst, n, S, lut, q, t, tpl, ls); break;
                                    case 16: stage<16>(src, dst, n, S, lut, q, t, tpl, ls); break;
                                    default: break;
                                }
                            }
                            break;
                    }
                }
                break;
        }
    }

    // Makhoul permutation shared by DCT-II (gather on load) and DCT-III (scatter on store)
    B2_D static int makhoul(int p, int n) { return (p < (n + 1) / 2) ? 2 * p : 2 * (n - 1 - p) + 1; }

    // DST wrappers: logical real length L, input index i -> (source index, sign)
    B2_D static int dst_src(const b2_pass_params& P, int i, int L) { return (P.dst_flags & B2_DST_REV_IN) ? L - 1 - i : i; }
    B2_D static T dst_sgn_in(const b2_pass_params& P, int i) { return ((P.dst_flags & B2_DST_NEG_ODD_IN) && (i & 1)) ? T(-1) : T(1); }
    B2_D static int dst_dst(const b2_pass_params& P, int k, int L) { return (P.dst_flags & B2_DST_REV_OUT) ? L - 1 - k : k; }
    B2_D static T dst_sgn_out(const b2_pass_params& P, int k) { return ((P.dst_flags & B2_DST_ALT_OUT) && (k & 1)) ? T(-1) : T(1); }

    // ---- load phase ---------------------------------------------------------------------------------------------
    // fills smem line `q` (local index) from the global line(s) it represents
    B2_D static void load_line(const b2_pass_params& P, X* sl, int64_t line_off, uint32_t gline, bool valid, int t,
                               int step) {
        const int n = (int)P.n;
        const bool inv = P.inverse != 0, inner = P.inner_inverse != 0;
        const X zero = mk<T>(T(0), T(0));
        switch (P.load_io) {
            default:
            case B2_IO_C2C:
            case B2_IO_R2C_EVEN: {   // (R2C even: the real line simply is n complex values)
                const X* in = (const X*)P.in + line_off;
                const X* mul = (const X*)P.aux0;
                for (int p = t; p < n; p += step) {
                    X v = zero;
                    if (valid && p < (int)P.in_len) {
                        v = in[(int64_t)p * P.in_es];
                        if (inv) v = swp(v);
                        if (P.ops & B2_OP_MUL_IN) v = v * ld_lut(mul + p);
                        if (inner) v = swp(v);
                    }
                    B2_SMEM_ST(sl, pad(p), v);
                }
            } break;
            case B2_IO_REAL: {       // real line -> (x, 0)
                const T* in = (const T*)P.in + line_off;
                const X* mul = (const X*)P.aux0;
                for (int p = t; p < n; p += step) {
                    X v = zero;
                    if (valid && p < (int)P.in_len) {
                        v.x = in[(int64_t)p * P.in_es];
                        if (P.ops & B2_OP_MUL_IN) v = v * ld_lut(mul + p);
                        if (inner) v = swp(v);
                    }
                    B2_SMEM_ST(sl, pad(p), v);
                }
            } break;
            case B2_IO_HERM: {       // half spectrum (in_len = L/2+1 of logical length L = aux_u1) -> full spectrum
                const X* in = (const X*)P.in + line_off;
                const X* mul = (const X*)P.aux0;
                const int L = (int)P.aux_u1;
                for (int p = t; p < n; p += step) {
                    X v = zero;
                    if (valid && p < L) {
                        v = (p <= L / 2) ? in[(int64_t)p * P.in_es] : conj(in[(int64_t)(L - p) * P.in_es]);
                        if (P.ops & B2_OP_MUL_IN) v = v * ld_lut(mul + p);
                        if (inner) v = swp(v);
                    }
                    B2_SMEM_ST(sl, pad(p), v);
                }
            } break;
            case B2_IO_C2R_EVEN: {
                // Zin[k] = (X[k] + conj X[n-k]) + i e^{+2 pi i k/N} (X[k] - conj X[n-k]),  N = 2n ; then the
                // unnormalised inverse FFT_n gives z[m] = x[2m] + i x[2m+1]   (aux0[k] = e^{-2 pi i k/N})
                const X* in = (const X*)P.in + line_off;
                const X* w = (const X*)P.aux0;
                for (int k = t; k < n; k += step) {
                    X v = zero;
                    if (valid) {
                        const X a = in[(int64_t)k * P.in_es];
                        const X b = conj(in[(int64_t)(n - k) * P.in_es]);
                        const X s = a + b, d = mulc(a - b, ld_lut(w + k));   // d = (a-b) * conj(w)
                        v = mk<T>(s.x - d.y, s.y + d.x);                      // s + i d
                        if (inner) v = swp(v);                                // planner sets inner_inverse = 1
                    }
                    B2_SMEM_ST(sl, pad(k), v);
                }
            } break;
            case B2_IO_DCT2: {
                // two real lines (2*gline, 2*gline+1) -> v[p] = xa[perm p] + i xb[perm p]
                const T* in = (const T*)P.in + line_off;
                const bool vb = valid && (2 * gline + 1 < P.aux_u0);
                for (int p = t; p < n; p += step) {
                    X v = zero;
                    if (valid) {
                        const int src = makhoul(p, n);
                        const int64_t s = (int64_t)dst_src(P, src, n) * P.in_es;
                        const T sg = dst_sgn_in(P, src);
                        v.x = sg * in[s];
                        if (vb) v.y = sg * in[s + P.in_gs];
                    }
                    B2_SMEM_ST(sl, pad(p), v);
                }
            } break;
            case B2_IO_DCT3: {
                // V'[k] = conj(c_k) [ (Xa[k] + Xb[n-k]) + i (Xb[k] - Xa[n-k]) ],  X[n] := 0, c_k = e^{-i pi k/2n}
                // followed by the unnormalised inverse FFT (aux0[k] = c_k)
                const T* in = (const T*)P.in + line_off;
                const X* c = (const X*)P.aux0;
                const bool vb = valid && (2 * gline + 1 < P.aux_u0);
                for (int k = t; k < n; k += step) {
                    X v = zero;
                    if (valid) {
                        const int64_t s0 = (int64_t)dst_src(P, k, n) * P.in_es;
                        const int64_t s1 = (int64_t)dst_src(P, k == 0 ? 0 : n - k, n) * P.in_es;
                        const T a0 = in[s0], a1 = (k == 0) ? T(0) : in[s1];
                        const T b0 = vb ? in[s0 + P.in_gs] : T(0), b1 = (vb && k != 0) ? in[s1 + P.in_gs] : T(0);
                        v = mulc(mk<T>(a0 + b1, b0 - a1), ld_lut(c + k));
                        if (inner) v = swp(v);
                    }
                    B2_SMEM_ST(sl, pad(k), v);
                }
            } break;
            case B2_IO_DCT1: {
                // even extension of two real lines of length L = aux_u1 to n = 2L-2
                const T* in = (const T*)P.in + line_off;
                const int L = (int)P.aux_u1;
                const bool vb = valid && (2 * gline + 1 < P.aux_u0);
                for (int p = t; p < n; p += step) {
                    X v = zero;
                    if (valid) {
                        const int64_t s = (int64_t)(p < L ? p : n - p) * P.in_es;
                        v.x = in[s];
                        if (vb) v.y = in[s + P.in_gs];
                    }
                    B2_SMEM_ST(sl, pad(p), v);
                }
            } break;
            case B2_IO_DST1: {
                // odd extension of two real lines of length L = aux_u1 to n = 2L+2: e[0]=0, e[1..L]=x, e[L+1]=0, e[L+2..]=-x reversed
                const T* in = (const T*)P.in + line_off;
                const int L = (int)P.aux_u1;
                const bool vb = valid && (2 * gline + 1 < P.aux_u0);
                for (int p = t; p < n; p += step) {
                    X v = zero;
                    if (valid && p != 0 && p != L + 1) {
                        const int j = p <= L ? p - 1 : n - p - 1;
                        const T sg = p <= L ? T(1) : T(-1);
                        const int64_t s = (int64_t)j * P.in_es;
                        v.x = sg * in[s];
                        if (vb) v.y = sg * in[s + P.in_gs];
                    }
                    B2_SMEM_ST(sl, pad(p), v);
                }
            } break;
            case B2_IO_DCT4: {
                // one real line of length N = 2n:  z'[m] = (x[2m] + i x[N-1-2m]) e^{-i pi m/N}   (aux0[m])
                const T* in = (const T*)P.in + line_off;
                const X* w = (const X*)P.aux0;
                for (int m = t; m < n; m += step) {
                    X v = zero;
                    if (valid) {
                        v = mk<T>(in[(int64_t)dst_src(P, 2 * m, 2 * n) * P.in_es], in[(int64_t)dst_src(P, 2 * n - 1 - 2 * m, 2 * n) * P.in_es]);
                        v = v * ld_lut(w + m);
                    }
                    B2_SMEM_ST(sl, pad(m), v);
                }
            } break;
            case B2_IO_DCT4_ODD: {
                // one real line of odd length L = aux_u1, n = 2L:  y[p] = x[p] e^{-i pi p/(2L)} (aux0[p]) for p < L, 0 above.
                //   X_k = 2 sum x_p cos(pi (2p+1)(2k+1)/(4L)) = 2 Re( e^{-i pi (2k+1)/(4L)} FFT_2L(y)[k] ),  k < L
                const T* in = (const T*)P.in + line_off;
                const X* w = (const X*)P.aux0;
                const int L = (int)P.aux_u1;
                for (int p = t; p < n; p += step) {
                    X v = zero;
                    if (valid && p < L) v = ld_lut(w + p) * in[(int64_t)dst_src(P, p, L) * P.in_es];
                    B2_SMEM_ST(sl, pad(p), v);
                }
            } break;
        }
    }

    // ---- store phase --------------------------------------------------------------------------------------------
    B2_D static void store_line(const b2_pass_params& P, const X* sl, int64_t line_off, uint32_t gline, uint32_t twline,
                                bool valid, int t, int step) {
        if (!valid) return;
        const int n = (int)P.n;
        const bool inv = P.inverse != 0, inner = P.inner_inverse != 0;
        const T sc = (T)P.scale;
        const bool do_scale = (P.ops & B2_OP_SCALE) != 0;
        switch (P.store_io) {
            default:
            case B2_IO_C2C:
            case B2_IO_C2R_EVEN: {   // (C2R even: the n complex results simply are the 2n reals)
                X* out = (X*)P.out + line_off;
                const int lim = (int)P.out_len < n ? (int)P.out_len : n;
                for (int p = t; p < lim; p += step) {
                    X v = B2_SMEM_LD(sl, pad(p));
                    if (inner) v = swp(v);
                    if (P.ops & B2_OP_MUL_OUT) v = v * ld_lut((const X*)P.aux1 + p);
                    if (P.ops & B2_OP_TWIDDLE_OUT) {
                        const uint64_t e = (uint64_t)twline * (uint64_t)p;
                        v = v * twiddle2<T>((const X*)P.tw_hi, (const X*)P.tw_lo, P.tw_shift, e);
                    }
                    if (do_scale) v = v * sc;
                    if (inv) v = swp(v);
                    out[(int64_t)p * P.out_es] = v;
                }
            } break;
            case B2_IO_REAL: {       // real part of the (inverse) transform
                T* out = (T*)P.out + line_off;
                const int lim = (int)P.out_len < n ? (int)P.out_len : n;
                for (int p = t; p < lim; p += step) {
                    X v = B2_SMEM_LD(sl, pad(p));
                    if (inner) v = swp(v);
                    if (P.ops & B2_OP_MUL_OUT) v = v * ld_lut((const X*)P.aux1 + p);
                    if (do_scale) v = v * sc;
                    out[(int64_t)p * P.out_es] = v.x;
                }
            } break;
            case B2_IO_R2C_EVEN: {
                // X[k] = 1/2 (Zk + conj Zn-k) - i/2 e^{-2 pi i k/N} (Zk - conj Zn-k),  k = 0..n   (aux0[k] = e^{-2 pi i k/N})
                X* out = (X*)P.out + line_off;
                const X* w = (const X*)P.aux0;
                for (int k = t; k <= n; k += step) {
                    const X a = B2_SMEM_LD(sl, pad(k == n ? 0 : k));
                    const X b = conj(B2_SMEM_LD(sl, pad(k == 0 ? 0 : n - k)));
                    const X s = a + b, d = (a - b) * ld_lut(w + k);
                    X v = mk<T>(T(0.5) * (s.x + d.y), T(0.5) * (s.y - d.x));   // (s - i d)/2
                    if (do_scale) v = v * sc;
                    out[(int64_t)k * P.out_es] = v;
                }
            } break;
            case B2_IO_DCT2: {
                // Xa[k] = Re(c_k (V[k] + conj V[n-k])),  Xb[k] = Im(c_k (V[k] - conj V[n-k]))   (aux0[k] = c_k)
                T* out = (T*)P.out + line_off;
                const X* c = (const X*)P.aux0;
                const bool vb = (2 * gline + 1 < P.aux_u0);
                for (int k = t; k < n; k += step) {
                    const X a = B2_SMEM_LD(sl, pad(k));
                    const X b = conj(B2_SMEM_LD(sl, pad(k == 0 ? 0 : n - k)));
                    const X ck = ld_lut(c + k);
                    const X s = ck * (a + b), d = ck * (a - b);
                    T ya = s.x, yb = d.y;
                    if (do_scale) { ya *= sc; yb *= sc; }
                    const int64_t o = (int64_t)dst_dst(P, k, n) * P.out_es;
                    out[o] = ya;
                    if (vb) out[o + P.out_gs] = yb;
                }
            } break;
            case B2_IO_DCT3: {
                // u = unnormalised inverse FFT result (already un-swapped here): y[perm p] = u[p]
                T* out = (T*)P.out + line_off;
                const bool vb = (2 * gline + 1 < P.aux_u0);
                for (int p = t; p < n; p += step) {
                    X v = B2_SMEM_LD(sl, pad(p));
                    if (inner) v = swp(v);
                    if (do_scale) v = v * sc;
                    const int dj = makhoul(p, n);
                    const int64_t o = (int64_t)dj * P.out_es;
                    const T sg = dst_sgn_out(P, dj);
                    out[o] = sg * v.x;
                    if (vb) out[o + P.out_gs] = sg * v.y;
                }
            } break;
            case B2_IO_DCT1: {
                T* out = (T*)P.out + line_off;
                const int L = (int)P.aux_u1;
                const bool vb = (2 * gline + 1 < P.aux_u0);
                for (int k = t; k < L; k += step) {
                    X v = B2_SMEM_LD(sl, pad(k));
                    if (do_scale) v = v * sc;
                    const int64_t o = (int64_t)k * P.out_es;
                    out[o] = v.x;
                    if (vb) out[o + P.out_gs] = v.y;
                }
            } break;
            case B2_IO_DST1: {
                // E[k] = -i Ya[k-1] + Yb[k-1]  ->  Ya[k-1] = -Im E[k], Yb[k-1] = Re E[k]
                T* out = (T*)P.out + line_off;
                const int L = (int)P.aux_u1;
                const bool vb = (2 * gline + 1 < P.aux_u0);
                for (int k = t; k < L; k += step) {
                    X v = B2_SMEM_LD(sl, pad(k + 1));
                    if (do_scale) v = v * sc;
                    const int64_t o = (int64_t)k * P.out_es;
                    out[o] = -v.y;
                    if (vb) out[o + P.out_gs] = v.x;
                }
            } break;
            case B2_IO_DCT4: {
                // C_q = aux1[q] Z[q];  X[2q] = 2 Re C_q,  X[N-1-2q] = -2 Im C_q     (aux1[q] = e^{-i pi (4q+1)/(4N)})
                T* out = (T*)P.out + line_off;
                const X* w = (const X*)P.aux1;
                for (int q2 = t; q2 < n; q2 += step) {
                    X v = B2_SMEM_LD(sl, pad(q2)) * ld_lut(w + q2);
                    T y0 = T(2) * v.x, y1 = T(-2) * v.y;
                    if (do_scale) { y0 *= sc; y1 *= sc; }
                    out[(int64_t)(2 * q2) * P.out_es] = dst_sgn_out(P, 2 * q2) * y0;
                    out[(int64_t)(2 * n - 1 - 2 * q2) * P.out_es] = dst_sgn_out(P, 2 * n - 1 - 2 * q2) * y1;
                }
            } break;
            case B2_IO_DCT4_ODD: {
                T* out = (T*)P.out + line_off;
                const X* w = (const X*)P.aux1;
                const int L = (int)P.aux_u1;
                for (int k = t; k < L; k += step) {
                    const X v = B2_SMEM_LD(sl, pad(k)) * ld_lut(w + k);
                    T y = T(2) * v.x;
                    if (do_scale) y *= sc;
                    out[(int64_t)k * P.out_es] = dst_sgn_out(P, k) * y;
                }
            } break;
        }
    }

    B2_D static void run(const b2_pass_params& P, unsigned char* smem_raw) {
        const int tid = threadIdx.x;
        const int Q = (int)P.q, TPL = (int)P.tpl, n = (int)P.n, ls = (int)P.line_stride;
        const uint32_t ngrp = (P.G + Q - 1) / Q;
        uint32_t rest = blockIdx.x;
        const uint32_t grp = rest % ngrp; rest /= ngrp;
        const uint32_t o0 = rest % P.nb[0]; rest /= P.nb[0];
        const uint32_t o1 = rest % P.nb[1]; rest /= P.nb[1];
        const uint32_t o2 = rest;
        const int64_t obase_in = (int64_t)o0 * P.in_bs[0] + (int64_t)o1 * P.in_bs[1] + (int64_t)o2 * P.in_bs[2];
        const int64_t obase_out = (int64_t)o0 * P.out_bs[0] + (int64_t)o1 * P.out_bs[1] + (int64_t)o2 * P.out_bs[2];
        X* buf0 = reinterpret_cast<X*>(smem_raw);
        X* buf1 = buf0 + (size_t)Q * ls;
        // real-pair operators address two real lines per complex line
        const uint32_t twsel = P.tw_sel;
        const uint32_t twbase = P.tw_line0 + (twsel == 1 ? o0 : (twsel == 2 ? o1 : (twsel == 3 ? o2 : 0)));
        const int in_mult = (P.load_io == B2_IO_DCT2 || P.load_io == B2_IO_DCT3 || P.load_io == B2_IO_DCT1 || P.load_io == B2_IO_DST1) ? 2 : 1;
        const int out_mult = (P.store_io == B2_IO_DCT2 || P.store_io == B2_IO_DCT3 || P.store_io == B2_IO_DCT1 || P.store_io == B2_IO_DST1) ? 2 : 1;

        // plain complex lines (B2_GEN_FUSE_IN / _OUT, set by the planner): no separate load / store phase, see stage_io
        const bool fin = (P.gen_flags & B2_GEN_FUSE_IN) != 0, fout = (P.gen_flags & B2_GEN_FUSE_OUT) != 0;
        int ql, tl, qs, ts;
        if (P.load_qfast) { ql = tid % Q; tl = tid / Q; } else { tl = tid % TPL; ql = tid / TPL; }
        if (P.store_qfast) { qs = tid % Q; ts = tid / Q; } else { ts = tid % TPL; qs = tid / TPL; }
        const uint32_t gll = grp * Q + ql, gls = grp * Q + qs;
        if (!fin) {   // load
            load_line(P, buf0 + ql * ls, obase_in + (int64_t)gll * in_mult * P.in_gs, gll, gll < P.G, tl, TPL);
            __syncthreads();
        }
        const X* lut = (const X*)P.lut;
        {   // stages (t fastest: neighbouring lanes take neighbouring butterflies of one line)
            const int t = tid % TPL, q = tid / TPL;
            int S = 1;
            X* src = buf0;
            X* dst = buf1;
            // Rader tables follow the stage twiddles in the LUT
            const X* rader = lut;
            {
                int S2 = 1;
                for (uint32_t s = 0; s < P.nstages; ++s) { if (s > 0) rader += ((int)P.radix[s] - 1) * S2; S2 *= (int)P.radix[s]; }
            }
            for (uint32_t s = 0; s < P.nstages; ++s) {
                const int r = (int)P.radix[s];
                const bool first = fin && s == 0, last = fout && s + 1 == P.nstages;
                if (first && last) {       // one radix: HBM -> registers -> HBM, in the load-side thread map
                    run_stage_io<true, true>(r, P, src, dst, S, lut, ql, tl, TPL, ls, obase_in + (int64_t)gll * P.in_gs,
                                             obase_out + (int64_t)gll * P.out_gs, twbase + (twsel == 0 ? gll : 0), gll < P.G);
                } else if (first) {
                    run_stage_io<true, false>(r, P, src, dst, S, lut, ql, tl, TPL, ls, obase_in + (int64_t)gll * P.in_gs, 0, 0, gll < P.G);
                } else if (last) {
                    run_stage_io<false, true>(r, P, src, dst, S, lut, qs, ts, TPL, ls, 0, obase_out + (int64_t)gls * P.out_gs,
                                              twbase + (twsel == 0 ? gls : 0), gls < P.G);
                } else if (r > 16) { rader_stage(r, src, dst, n, S, lut, rader, q, t, TPL, ls); rader += 2 * (r - 1); }
                else run_stage(r, src, dst, n, S, lut, q, t, TPL, ls);
                if (s > 0) lut += (r - 1) * S;
                S *= r;
                if (last) return;
                __syncthreads();
                X* tmp = src; src = dst; dst = tmp;
            }
            buf0 = src;   // final data
        }
        // store
        store_line(P, buf0 + qs * ls, obase_out + (int64_t)gls * out_mult * P.out_gs, gls,
                   twbase + (twsel == 0 ? gls : 0), gls < P.G, ts, TPL);
    }
};

#if defined(__CUDACC__)
// at most 256 threads per CTA; resident CTAs the compiler must leave room for: 4 / 3 / 2 for the radix classes 8 / 11 / 16
template <typename T, int RMAX>
__global__ void __launch_bounds__(256, (RMAX <= 8 ? 4 : (RMAX <= 11 ? 3 : 2)))
generic_kernel(const __grid_constant__ b2_pass_params P) {
    extern __shared__ __align__(16) unsigned char b2_smem_raw[];
    Generic<T, RMAX>::run(P, b2_smem_raw);
}
#endif

}  // namespace b200fft
