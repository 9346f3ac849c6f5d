// Host-side twiddle tables, computed in long double and rounded once to the target precision.
// Layout of the stage LUT matches RList::lut_off():  for every stage s >= 1 (radix r, stageSize S)
// r-1 blocks of S entries,  block k-1 entry j = exp(-2*pi*i * j*k / (S*r)).
// (The reference builds equivalent per-stage tables in VkFFT_AllocateLUT, vkFFT_ManageLUT.h:675-821, but
//  only uses them by default in FP64; here they are always used, in both precisions.)
#pragma once
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

template <typename T>
inline std::vector<T> make_stage_lut(const int* radices, int ns) {
    std::vector<T> out;
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

}  // namespace b200fft
