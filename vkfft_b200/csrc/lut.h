// Host-side twiddle tables, computed in long double and rounded once to the target precision.
// Layout of the stage LUT matches RList::lut_off():  for every stage s >= 1 (radix r, stageSize S)
// r-1 blocks of S entries,  block k-1 entry j = exp(-2*pi*i * j*k / (S*r)).
// (The reference builds equivalent per-stage tables in VkFFT_AllocateLUT, vkFFT_ManageLUT.h:675-821, but
//  only uses them by default in FP64; here they are always used, in both precisions.)
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

namespace b200fft {

static const long double B2_PI_L = 3.14159265358979323846264338327950288419716939937510L;

// exp(-2*pi*i * num/den) with exact octant reduction (num, den integers)
inline void unit_root(uint64_t num, uint64_t den, long double& c, long double& s) {
    num %= den;
    // reduce to first octant for accuracy
    const long double a = 2.0L * B2_PI_L * (long double)num / (long double)den;
    // use symmetry around multiples of pi/4 via integer comparisons
    uint64_t n8 = num * 8;
    if (n8 == 0) { c = 1; s = 0; return; }
    if (n8 == den * 2) { c = 0; s = -1; return; }
    if (n8 == den * 4) { c = -1; s = 0; return; }
    if (n8 == den * 6) { c = 0; s = 1; return; }
    c = cosl(a);
    s = -sinl(a);
}

// smallest primitive root of the prime p
inline uint64_t primitive_root(uint64_t p) {
    auto powmod = [&](uint64_t b, uint64_t e) { uint64_t r = 1; b %= p; while (e) { if (e & 1) r = r * b % p; b = b * b % p; e >>= 1; } return r; };
    std::vector<uint64_t> fac;
    uint64_t m = p - 1;
    for (uint64_t f = 2; f * f <= m; ++f)
        if (m % f == 0) { fac.push_back(f); while (m % f == 0) m /= f; }
    if (m > 1) fac.push_back(m);
    for (uint64_t g = 2; g < p; ++g) {
        bool ok = true;
        for (uint64_t f : fac) if (powmod(g, (p - 1) / f) == 1) { ok = false; break; }
        if (ok) return g;
    }
    return 1;
}

// Rader tables of one prime-radix stage (vkFFT_RaderKernels.h:1278 "mult" form; generator powers as in
// VkFFTGenerateRaderFFTKernel, vkFFT_RecursiveFFTGenerators.h:1073-1103):  p-1 entries  b_m = exp(-2 pi i g^{-m}/p)
// followed by p-1 entries (g^m mod p, g^{-m} mod p) stored as numbers.
template <typename T>
inline void append_rader_tables(std::vector<T>& out, uint64_t p) {
    const uint64_t g = primitive_root(p);
    std::vector<uint64_t> gp(p - 1), gi(p - 1);
    uint64_t v = 1;
    for (uint64_t m = 0; m < p - 1; ++m) { gp[m] = v; v = v * g % p; }
    for (uint64_t m = 0; m < p - 1; ++m) gi[m] = gp[(p - 1 - m) % (p - 1)];
    for (uint64_t m = 0; m < p - 1; ++m) {
        long double c, s;
        unit_root(gi[m], p, c, s);
        out.push_back((T)c); out.push_back((T)s);
    }
    for (uint64_t m = 0; m < p - 1; ++m) { out.push_back((T)gp[m]); out.push_back((T)gi[m]); }
}

template <typename T>
inline std::vector<T> make_stage_lut(const int* radices, int ns) {
    std::vector<T> out;
    if (ns <= 0) return out;
    uint64_t S = radices[0];
    for (int s = 1; s < ns; ++s) {
        const uint64_t r = radices[s];
        for (uint64_t k = 1; k < r; ++k)
            for (uint64_t j = 0; j < S; ++j) {
                long double c, sn;
                unit_root(j * k, S * r, c, sn);
                out.push_back((T)c);
                out.push_back((T)sn);
            }
        S *= r;
    }
    // Rader tables of the prime stages (radix > 16), in stage order, after all stage twiddles
    for (int s2 = 0; s2 < ns; ++s2)
        if (radices[s2] > 16) append_rader_tables<T>(out, (uint64_t)radices[s2]);
    return out;
}

// two-level table for W_M^m, m < M:  lo[i] = W^i (i < 2^shift), hi[i] = W^(i << shift)
template <typename T>
inline void make_twolevel(uint64_t M, uint32_t& shift, std::vector<T>& hi, std::vector<T>& lo) {
    uint32_t bits = 0;
    while ((1ull << bits) < M) ++bits;
    shift = (bits + 1) / 2;
    const uint64_t nlo = 1ull << shift;
    const uint64_t nhi = ((M - 1) >> shift) + 1;
    lo.resize(2 * nlo);
    hi.resize(2 * nhi);
    for (uint64_t i = 0; i < nlo; ++i) {
        long double c, s;
        unit_root(i, M, c, s);
        lo[2 * i] = (T)c; lo[2 * i + 1] = (T)s;
    }
    for (uint64_t i = 0; i < nhi; ++i) {
        long double c, s;
        unit_root(i << shift, M, c, s);
        hi[2 * i] = (T)c; hi[2 * i + 1] = (T)s;
    }
}


// ---- auxiliary tables of the fused real-transform / Bluestein operators ----------------------------------------
enum AuxKind {
    AUX_R2C = 1,          // a = N : e^{-2 pi i k/N},           k = 0..N/2           (vkFFT_ManageLUT.h:1418 R2C LUT)
    AUX_DCT23 = 2,        // a = n : e^{-i pi k/(2n)},           k = 0..n-1           (vkFFT_ManageLUT.h:800-806)
    AUX_DCT4_PRE = 3,     // a = N : e^{-i pi m/N},              m = 0..N/2-1         (vkFFT_ManageLUT.h:807-820)
    AUX_DCT4_POST = 4,
    AUX_DCT4ODD_PRE = 9,  // a = N : e^{-i pi n/(2N)},           n = 0..N-1   (odd-length DCT-IV through a 2N-point transform)
    AUX_DCT4ODD_POST = 10,// a = N : e^{-i pi (2k+1)/(4N)},      k = 0..N-1    // a = N : e^{-i pi (4q+1)/(4N)},      q = 0..N/2-1
    AUX_BLUE_CHIRP = 5,   // a = N : e^{-i pi n^2/N},            n = 0..N-1           (vkFFT_RecursiveFFTGenerators.h:140)
    AUX_BLUE_FILTER = 6,  // a = N, b = M : FFT_M(e^{+i pi m^2/N} wrapped) / M        (vkFFT_RecursiveFFTGenerators.h:241-298)
};

// mixed-radix DFT in long double for any length (host, table generation only): decimation in time by the smallest prime factor
inline void host_fft_any(std::vector<long double>& re, std::vector<long double>& im) {
    const size_t n = re.size();
    if (n <= 1) return;
    size_t p = 2;
    while (p * p <= n && n % p) ++p;
    if (n % p) p = n;
    const size_t m = n / p;
    std::vector<std::vector<long double>> sr(p, std::vector<long double>(m)), si(p, std::vector<long double>(m));
    for (size_t r = 0; r < p; ++r)
        for (size_t j = 0; j < m; ++j) { sr[r][j] = re[j * p + r]; si[r][j] = im[j * p + r]; }
    for (size_t r = 0; r < p; ++r) host_fft_any(sr[r], si[r]);
    for (size_t k = 0; k < m; ++k)
        for (size_t q = 0; q < p; ++q) {
            long double ar = 0, ai = 0;
            for (size_t r = 0; r < p; ++r) {
                long double c, s;
                unit_root((r * (k + m * q)) % n, n, c, s);      // e^{-2 pi i r (k + m q)/n}
                ar += sr[r][k] * c - si[r][k] * s;
                ai += sr[r][k] * s + si[r][k] * c;
            }
            re[k + m * q] = ar; im[k + m * q] = ai;
        }
}

// in-place radix-2 FFT in long double (host, table generation only); n must be a power of two
inline void host_fft_pow2(std::vector<long double>& re, std::vector<long double>& im) {
    const size_t n = re.size();
    for (size_t i = 1, j = 0; i < n; ++i) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { std::swap(re[i], re[j]); std::swap(im[i], im[j]); }
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        for (size_t i = 0; i < n; i += len)
            for (size_t k = 0; k < len / 2; ++k) {
                long double c, s;
                unit_root(k, len, c, s);
                const size_t a = i + k, b = i + k + len / 2;
                const long double tr = re[b] * c - im[b] * s, ti = re[b] * s + im[b] * c;
                re[b] = re[a] - tr; im[b] = im[a] - ti;
                re[a] += tr; im[a] += ti;
            }
    }
}

template <typename T>
inline std::vector<T> make_aux(int kind, uint64_t a, uint64_t b) {
    std::vector<T> out;
    auto push = [&](long double c, long double s) { out.push_back((T)c); out.push_back((T)s); };
    long double c, s;
    switch (kind) {
        case AUX_R2C:
            for (uint64_t k = 0; k <= a / 2; ++k) { unit_root(k, a, c, s); push(c, s); }
            break;
        case AUX_DCT23:
            for (uint64_t k = 0; k < a; ++k) { unit_root(k, 4 * a, c, s); push(c, s); }
            break;
        case AUX_DCT4_PRE:
            for (uint64_t m = 0; m < a / 2; ++m) { unit_root(m, 2 * a, c, s); push(c, s); }
            break;
        case AUX_DCT4_POST:
            for (uint64_t q = 0; q < a / 2; ++q) { unit_root(4 * q + 1, 8 * a, c, s); push(c, s); }
            break;
        case AUX_DCT4ODD_PRE:
            for (uint64_t n = 0; n < a; ++n) { unit_root(n, 4 * a, c, s); push(c, s); }
            break;
        case AUX_DCT4ODD_POST:
            for (uint64_t k = 0; k < a; ++k) { unit_root(2 * k + 1, 8 * a, c, s); push(c, s); }
            break;
        case AUX_BLUE_CHIRP:
            for (uint64_t n = 0; n < a; ++n) { unit_root((n * n) % (2 * a), 2 * a, c, s); push(c, s); }
            break;
        case AUX_BLUE_FILTER: {
            std::vector<long double> re(b, 0.0L), im(b, 0.0L);
            for (uint64_t m = 0; m < a; ++m) {
                unit_root((m * m) % (2 * a), 2 * a, c, s);   // e^{-i pi m^2/N}; the filter uses the conjugate
                re[m] = c; im[m] = -s;
                if (m) { re[b - m] = c; im[b - m] = -s; }
            }
            if ((b & (b - 1)) == 0) host_fft_pow2(re, im); else host_fft_any(re, im);
            for (uint64_t k = 0; k < b; ++k) push(re[k] / (long double)b, im[k] / (long double)b);
        } break;
        default: break;
    }
    return out;
}

}  // namespace b200fft
