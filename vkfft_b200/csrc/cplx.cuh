// Complex arithmetic + host/device portability macros shared by every kernel in the engine.
//
// The same headers are compiled three ways:
//   * nvcc for sm_100a            -> the product (libb200fft.so)
//   * g++ with tests/emu/cuda_emu.h -> a CPU "one OS thread per CUDA thread" emulation used only by
//                                    the CPU test-suite to check index maps / bank conflicts
//   * g++ for the host planner (only the POD parts)
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define B2_HD __host__ __device__ __forceinline__
#define B2_D __device__ __forceinline__
#else
#define B2_HD inline
#define B2_D inline
#endif

namespace b200fft {

template <typename T>
struct alignas(2 * sizeof(T)) cpx {
    T x, y;
};
using cf32 = cpx<float>;
using cf64 = cpx<double>;

template <typename T> B2_HD cpx<T> mk(T a, T b) { cpx<T> r; r.x = a; r.y = b; return r; }

// FP32 complex arithmetic on sm_100a uses the packed two-lane instructions (FADD2 / FMUL2 / FFMA2, PTX add/mul/fma.f32x2):
// one instruction per complex add, two per complex multiply, and quarter turns / conjugation are operand modifiers
// (the SASS operands take a .LO_HI swap, a per-half negation and a 32-bit broadcast), so an interleaved (re, im) pair
// is processed whole.  These kernels are issue-bound (profiles/README.md), and this halves their floating-point
// instruction count.  The reference emits scalar code for every backend (vkFFT_MathUtils.h).  Host code, the CPU
// emulation and FP64 use the plain component-wise form below; results agree to rounding (same operations, the product
// a.x*b.x is rounded before the fused multiply-add instead of a.y*b.y).
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000) && !defined(B2_NO_F32X2)
#define B2_F32X2 1
#define B2_PK(a) make_float2((a).x, (a).y)
template <typename T> B2_D cpx<T> b2_unpk(float2 v) { return mk<T>((T)v.x, (T)v.y); }
#else
#define B2_F32X2 0
#endif

template <typename T> B2_HD cpx<T> operator+(cpx<T> a, cpx<T> b) {
#if B2_F32X2
    if constexpr (sizeof(T) == 4) return b2_unpk<T>(__fadd2_rn(B2_PK(a), B2_PK(b)));
    else
#endif
    return mk<T>(a.x + b.x, a.y + b.y);
}
template <typename T> B2_HD cpx<T> operator-(cpx<T> a, cpx<T> b) {
#if B2_F32X2
    if constexpr (sizeof(T) == 4) return b2_unpk<T>(__fadd2_rn(B2_PK(a), make_float2(-b.x, -b.y)));
    else
#endif
    return mk<T>(a.x - b.x, a.y - b.y);
}
template <typename T> B2_HD cpx<T> operator*(cpx<T> a, cpx<T> b) {
#if B2_F32X2
    if constexpr (sizeof(T) == 4) {
        const float2 p = __fmul2_rn(B2_PK(a), make_float2(b.x, b.x));
        return b2_unpk<T>(__ffma2_rn(make_float2(-a.y, a.x), make_float2(b.y, b.y), p));
    } else
#endif
    return mk<T>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
template <typename T> B2_HD cpx<T> operator*(cpx<T> a, T s) {
#if B2_F32X2
    if constexpr (sizeof(T) == 4) return b2_unpk<T>(__fmul2_rn(B2_PK(a), make_float2(s, s)));
    else
#endif
    return mk<T>(a.x * s, a.y * s);
}
// a * conj(b)
template <typename T> B2_HD cpx<T> mulc(cpx<T> a, cpx<T> b) {
#if B2_F32X2
    if constexpr (sizeof(T) == 4) {
        const float2 p = __fmul2_rn(B2_PK(a), make_float2(b.x, b.x));
        return b2_unpk<T>(__ffma2_rn(make_float2(a.y, -a.x), make_float2(b.y, b.y), p));
    } else
#endif
    return mk<T>(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
// c + a * s  (real scalar s)
template <typename T> B2_HD cpx<T> fma_s(cpx<T> a, T s, cpx<T> c) {
#if B2_F32X2
    if constexpr (sizeof(T) == 4) return b2_unpk<T>(__ffma2_rn(B2_PK(a), make_float2(s, s), B2_PK(c)));
    else
#endif
    return mk<T>(c.x + a.x * s, c.y + a.y * s);
}
template <typename T> B2_HD cpx<T> conj(cpx<T> a) { return mk<T>(a.x, -a.y); }
// multiply by -i  (forward-transform quarter turn):  (x + iy)(-i) = y - ix
template <typename T> B2_HD cpx<T> mul_mi(cpx<T> a) { return mk<T>(a.y, -a.x); }
// multiply by +i
template <typename T> B2_HD cpx<T> mul_pi(cpx<T> a) { return mk<T>(-a.y, a.x); }
// swap real/imag: IFFT(x) = swap(FFT(swap(x)))  -- how every inverse plan runs forward code
template <typename T> B2_HD cpx<T> swp(cpx<T> a) { return mk<T>(a.y, a.x); }

}  // namespace b200fft
