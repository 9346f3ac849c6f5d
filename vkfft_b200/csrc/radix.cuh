// Register-resident small DFTs ("butterflies") of radix 2..16 and 32, forward sign (e^{-2*pi*i*nk/R}).
//
// What the reference emits as text per plan (vkFFT_RadixKernels.h:43 radix-2, :73 radix-3, :167 radix-4,
// :310 radix-5, :663 radix-7, :802 radix-8, :1439 radix-11, :1761 radix-13, :2361 radix-16, :2747 radix-32)
// is written here once as templates:
//   * odd primes use the real-symmetric form: fold x_j +- x_{P-j}, then an h x h real cos block and an
//     h x h real sin block (h = (P-1)/2) -- pure FMA chains, no generator permutation needed;
//   * everything composite is a compile-time Cooley-Tukey split R = R1*R2 whose inner twiddles are
//     immediates; quarter/eighth turns are strength-reduced (no multiply for -i, two for sqrt(1/2)(1-i)).
// All indexing is compile-time, so after unrolling the "transposes" are register renames.
// The inverse transform never needs its own butterflies: kernels swap re/im on load and store.
#pragma once
#include "cplx.cuh"
#include "radix_consts.cuh"

namespace b200fft {

template <typename T, int R> B2_HD constexpr T wcos_r(int k) {
    if constexpr (R == 1) return T(1);
    else if constexpr (R == 2) return (k & 1) ? T(-1) : T(1);
    else if constexpr (R == 3) return rc::wcos3<T>(k);
    else if constexpr (R == 4) return rc::wcos4<T>(k);
    else if constexpr (R == 5) return rc::wcos5<T>(k);
    else if constexpr (R == 6) return rc::wcos6<T>(k);
    else if constexpr (R == 7) return rc::wcos7<T>(k);
    else if constexpr (R == 8) return rc::wcos8<T>(k);
    else if constexpr (R == 9) return rc::wcos9<T>(k);
    else if constexpr (R == 10) return rc::wcos10<T>(k);
    else if constexpr (R == 11) return rc::wcos11<T>(k);
    else if constexpr (R == 12) return rc::wcos12<T>(k);
    else if constexpr (R == 13) return rc::wcos13<T>(k);
    else if constexpr (R == 14) return rc::wcos14<T>(k);
    else if constexpr (R == 15) return rc::wcos15<T>(k);
    else if constexpr (R == 16) return rc::wcos16<T>(k);
    else if constexpr (R == 17) return rc::wcos17<T>(k);
    else if constexpr (R == 19) return rc::wcos19<T>(k);
    else if constexpr (R == 23) return rc::wcos23<T>(k);
    else if constexpr (R == 29) return rc::wcos29<T>(k);
    else if constexpr (R == 31) return rc::wcos31<T>(k);
    else return rc::wcos32<T>(k);
}
template <typename T, int R> B2_HD constexpr T wsin_r(int k) {
    if constexpr (R <= 2) return T(0);
    else if constexpr (R == 3) return rc::wsin3<T>(k);
    else if constexpr (R == 4) return rc::wsin4<T>(k);
    else if constexpr (R == 5) return rc::wsin5<T>(k);
    else if constexpr (R == 6) return rc::wsin6<T>(k);
    else if constexpr (R == 7) return rc::wsin7<T>(k);
    else if constexpr (R == 8) return rc::wsin8<T>(k);
    else if constexpr (R == 9) return rc::wsin9<T>(k);
    else if constexpr (R == 10) return rc::wsin10<T>(k);
    else if constexpr (R == 11) return rc::wsin11<T>(k);
    else if constexpr (R == 12) return rc::wsin12<T>(k);
    else if constexpr (R == 13) return rc::wsin13<T>(k);
    else if constexpr (R == 14) return rc::wsin14<T>(k);
    else if constexpr (R == 15) return rc::wsin15<T>(k);
    else if constexpr (R == 16) return rc::wsin16<T>(k);
    else if constexpr (R == 17) return rc::wsin17<T>(k);
    else if constexpr (R == 19) return rc::wsin19<T>(k);
    else if constexpr (R == 23) return rc::wsin23<T>(k);
    else if constexpr (R == 29) return rc::wsin29<T>(k);
    else if constexpr (R == 31) return rc::wsin31<T>(k);
    else return rc::wsin32<T>(k);
}

// a * W_R^e  with W_R = exp(-2*pi*i/R), e known at compile time after unrolling.
template <typename T, int R>
B2_HD cpx<T> mul_wconst(cpx<T> a, int e) {
    e %= R;
    if (e == 0) return a;
    if (2 * e == R) return mk<T>(-a.x, -a.y);
    if (4 * e == R) return mul_mi(a);
    if (4 * e == 3 * R) return mul_pi(a);
    constexpr T h = T(7.07106781186547524400844362104849039e-01L);
    // eighth turns: one (packed) add with a quarter turn of the same value, one scale
    if (8 * e == R) return (a + mul_mi(a)) * h;              // (1-i)/sqrt2 : ((x+y), (y-x)) h
    if (8 * e == 3 * R) return (mul_mi(a) - a) * h;          // (-1-i)/sqrt2: ((y-x), -(x+y)) h
    if (8 * e == 5 * R) return (mul_pi(a) - a) * h;          // (-1+i)/sqrt2: (-(x+y), (x-y)) h
    if (8 * e == 7 * R) return (a + mul_pi(a)) * h;          // (1+i)/sqrt2 : ((x-y), (x+y)) h
    const T c = wcos_r<T, R>(e), s = wsin_r<T, R>(e);        // W = c - i s :  a c + (-i a) s
    return fma_s(mul_mi(a), s, a * c);
}

template <int R> struct radix_split { static constexpr int r1 = 1, r2 = R; };  // prime
template <> struct radix_split<4> { static constexpr int r1 = 2, r2 = 2; };
template <> struct radix_split<6> { static constexpr int r1 = 2, r2 = 3; };
template <> struct radix_split<8> { static constexpr int r1 = 2, r2 = 4; };
template <> struct radix_split<9> { static constexpr int r1 = 3, r2 = 3; };
template <> struct radix_split<10> { static constexpr int r1 = 2, r2 = 5; };
template <> struct radix_split<12> { static constexpr int r1 = 3, r2 = 4; };
template <> struct radix_split<14> { static constexpr int r1 = 2, r2 = 7; };
template <> struct radix_split<15> { static constexpr int r1 = 3, r2 = 5; };
template <> struct radix_split<16> { static constexpr int r1 = 4, r2 = 4; };
template <> struct radix_split<32> { static constexpr int r1 = 4, r2 = 8; };

template <int R, typename T> B2_HD void dft(cpx<T>* x);

// odd prime P, real-symmetric form
template <int P, typename T>
B2_HD void dft_prime(cpx<T>* x) {
    constexpr int h = (P - 1) / 2;
    cpx<T> s[h], d[h];
#pragma unroll
    for (int j = 0; j < h; ++j) {
        s[j] = x[j + 1] + x[P - 1 - j];
        d[j] = x[j + 1] - x[P - 1 - j];
    }
    cpx<T> x0 = x[0];
    cpx<T> sum = x0;
#pragma unroll
    for (int j = 0; j < h; ++j) sum = sum + s[j];
    x[0] = sum;
#pragma unroll
    for (int k = 1; k <= h; ++k) {
        cpx<T> a = x0, b = mk<T>(T(0), T(0));
#pragma unroll
        for (int j = 1; j <= h; ++j) {
            const T c = wcos_r<T, P>((j * k) % P), sn = wsin_r<T, P>((j * k) % P);
            a = fma_s(s[j - 1], c, a);
            b = fma_s(d[j - 1], sn, b);
        }
        // X_k = A - iB ; X_{P-k} = A + iB
        x[k] = a + mul_mi(b);
        x[P - k] = a + mul_pi(b);
    }
}

template <int R, typename T>
B2_HD void dft(cpx<T>* x) {
    if constexpr (R == 1) {
        return;
    } else if constexpr (R == 2) {
        cpx<T> a = x[0], b = x[1];
        x[0] = a + b;
        x[1] = a - b;
    } else if constexpr (radix_split<R>::r1 == 1) {
        dft_prime<R, T>(x);
    } else {
        constexpr int R1 = radix_split<R>::r1, R2 = radix_split<R>::r2;
        // n = R2*n1 + n2 ; k = k1 + R1*k2
        cpx<T> y[R];
#pragma unroll
        for (int n2 = 0; n2 < R2; ++n2) {
            cpx<T> a[R1];
#pragma unroll
            for (int n1 = 0; n1 < R1; ++n1) a[n1] = x[R2 * n1 + n2];
            dft<R1, T>(a);
#pragma unroll
            for (int k1 = 0; k1 < R1; ++k1) y[n2 * R1 + k1] = mul_wconst<T, R>(a[k1], n2 * k1);
        }
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) {
            cpx<T> b[R2];
#pragma unroll
            for (int n2 = 0; n2 < R2; ++n2) b[n2] = y[n2 * R1 + k1];
            dft<R2, T>(b);
#pragma unroll
            for (int k2 = 0; k2 < R2; ++k2) x[k1 + R1 * k2] = b[k2];
        }
    }
}

}  // namespace b200fft
