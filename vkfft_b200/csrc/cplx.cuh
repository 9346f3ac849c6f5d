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
template <typename T> B2_HD cpx<T> operator+(cpx<T> a, cpx<T> b) { return mk<T>(a.x + b.x, a.y + b.y); }
template <typename T> B2_HD cpx<T> operator-(cpx<T> a, cpx<T> b) { return mk<T>(a.x - b.x, a.y - b.y); }
template <typename T> B2_HD cpx<T> operator*(cpx<T> a, cpx<T> b) {
    return mk<T>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
template <typename T> B2_HD cpx<T> operator*(cpx<T> a, T s) { return mk<T>(a.x * s, a.y * s); }
// a * conj(b)
template <typename T> B2_HD cpx<T> mulc(cpx<T> a, cpx<T> b) {
    return mk<T>(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
template <typename T> B2_HD cpx<T> conj(cpx<T> a) { return mk<T>(a.x, -a.y); }
// multiply by -i  (forward-transform quarter turn):  (x + iy)(-i) = y - ix
template <typename T> B2_HD cpx<T> mul_mi(cpx<T> a) { return mk<T>(a.y, -a.x); }
// multiply by +i
template <typename T> B2_HD cpx<T> mul_pi(cpx<T> a) { return mk<T>(-a.y, a.x); }
// swap real/imag: IFFT(x) = swap(FFT(swap(x)))  -- how every inverse plan runs forward code
template <typename T> B2_HD cpx<T> swp(cpx<T> a) { return mk<T>(a.y, a.x); }

}  // namespace b200fft
