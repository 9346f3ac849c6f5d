/* TEST INFRASTRUCTURE (oracle) -- never linked into or loaded by the product.
 *
 * Plain-C, double-precision restatement of the algorithm the reference's generated kernels implement for
 * C2C transforms: the DIT Stockham autosort FFT with the reference's stage bookkeeping
 *   - stageSize S starts at 1 and is multiplied by the stage radix after every stage,
 *     stageAngle starts at -pi (forward) / +pi (inverse)            vkFFT_FFT.h:151-239 (:154, :236-237)
 *   - in stage (S, r) butterfly b uses j = b mod S; input leg i is multiplied by
 *     exp(sign*2*pi*i * j*i / (S*r)) BEFORE the butterfly           vkFFT_RadixKernels.h:43-71, :336-390
 *   - outputs scatter to (b - j)*r + j + k*S                        vkFFT_RadixShuffle.h:34
 *   - the radix-r butterfly itself is the plain length-r DFT (the reference's hand-factored radix-2..13
 *     butterflies, vkFFT_RadixKernels.h:43-2126, are algebraically this DFT)
 * plus the Four-Step decomposition for long sequences          vkFFT_4step.h:31-119, API guide :495-551.
 * Pinned in tests/test_oracle.py against the O(N^2) definition (API guide :263-352), against pocketfft and
 * against the committed golden vectors produced by the reference's CUDA backend.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

static const double PI = 3.14159265358979323846264338327950288;

/* one Stockham transform of length n with the given radix list; in -> out (both n complex, interleaved).
 * work must hold n complex. sign = -1 forward, +1 inverse (unnormalised). */
static void stockham_1d(const double* in, double* out, double* work, long n, const int* radix, int nstage, int sign) {
    const double* src = in;
    double* bufs[2] = {work, out};
    /* choose ping-pong parity so that the last stage lands in `out` */
    int which = (nstage % 2 == 1) ? 1 : 0;
    long S = 1;
    for (int s = 0; s < nstage; ++s) {
        const int r = radix[s];
        const long nb = n / r;
        double* dst = bufs[which];
        for (long b = 0; b < nb; ++b) {
            const long j = b % S;
            double lr[64], li[64];
            for (int i = 0; i < r; ++i) {
                const double xr = src[2 * (b + i * nb)], xi = src[2 * (b + i * nb) + 1];
                const double ang = sign * 2.0 * PI * (double)(j * i) / (double)(S * r);
                const double c = cos(ang), sn = sin(ang);
                lr[i] = xr * c - xi * sn;
                li[i] = xr * sn + xi * c;
            }
            for (int k = 0; k < r; ++k) {
                double ar = 0, ai = 0;
                for (int i = 0; i < r; ++i) {
                    const double ang = sign * 2.0 * PI * (double)((long)i * k % r) / (double)r;
                    const double c = cos(ang), sn = sin(ang);
                    ar += lr[i] * c - li[i] * sn;
                    ai += lr[i] * sn + li[i] * c;
                }
                const long o = (b - j) * r + j + k * S;
                dst[2 * o] = ar;
                dst[2 * o + 1] = ai;
            }
        }
        src = dst;
        which ^= 1;
        S *= r;
    }
    if (nstage == 0) memcpy(out, in, sizeof(double) * 2 * n);
}

/* factor n into radices <= 13 (largest first, like vkFFT_Scheduler.h:3230-3237); returns count or -1 */
static int factor(long n, int* radix) {
    static const int cand[] = {13, 11, 8, 7, 5, 4, 3, 2};
    int ns = 0;
    for (int c = 0; c < 8; ++c)
        while (n % cand[c] == 0 && n > 1) {
            radix[ns++] = cand[c];
            n /= cand[c];
            if (ns >= 60) return -1;
        }
    return n == 1 ? ns : -1;
}

/* batched 1-D C2C, in place on `data` (batch x n complex doubles). returns 0, or -1 if n is not 13-smooth */
int oracle_stockham_c2c(double* data, long n, long batch, int inverse) {
    int radix[64];
    int ns = factor(n, radix);
    if (ns < 0) return -1;
    double* tmp = (double*)malloc(sizeof(double) * 2 * n);
    double* work = (double*)malloc(sizeof(double) * 2 * n);
    for (long b = 0; b < batch; ++b) {
        stockham_1d(data + 2 * b * n, tmp, work, n, radix, ns, inverse ? +1 : -1);
        memcpy(data + 2 * b * n, tmp, sizeof(double) * 2 * n);
    }
    free(tmp);
    free(work);
    return 0;
}

/* Four-Step for n = n1*n2 (API guide :495-551): strided length-n1 transforms, multiply by
 * exp(sign*2*pi*i * x*k/n), contiguous length-n2 transforms, transposed write -> natural order. */
int oracle_four_step_c2c(double* data, long n1, long n2, long batch, int inverse) {
    const long n = n1 * n2;
    const int sign = inverse ? +1 : -1;
    int r1[64], r2[64];
    int ns1 = factor(n1, r1), ns2 = factor(n2, r2);
    if (ns1 < 0 || ns2 < 0) return -1;
    const long m = n1 > n2 ? n1 : n2;
    double* col = (double*)malloc(sizeof(double) * 2 * m);
    double* res = (double*)malloc(sizeof(double) * 2 * m);
    double* work = (double*)malloc(sizeof(double) * 2 * m);
    double* tmp = (double*)malloc(sizeof(double) * 2 * n);
    for (long b = 0; b < batch; ++b) {
        double* x = data + 2 * b * n;
        for (long c = 0; c < n2; ++c) { /* upload 1: columns of stride n2 */
            for (long i = 0; i < n1; ++i) {
                col[2 * i] = x[2 * (i * n2 + c)];
                col[2 * i + 1] = x[2 * (i * n2 + c) + 1];
            }
            stockham_1d(col, res, work, n1, r1, ns1, sign);
            for (long k = 0; k < n1; ++k) {
                const double ang = sign * 2.0 * PI * (double)(c * k) / (double)n;
                const double cs = cos(ang), sn = sin(ang);
                tmp[2 * (k * n2 + c)] = res[2 * k] * cs - res[2 * k + 1] * sn;
                tmp[2 * (k * n2 + c) + 1] = res[2 * k] * sn + res[2 * k + 1] * cs;
            }
        }
        for (long k1 = 0; k1 < n1; ++k1) { /* upload 0: rows, transposed write */
            stockham_1d(tmp + 2 * k1 * n2, res, work, n2, r2, ns2, sign);
            for (long k2 = 0; k2 < n2; ++k2) {
                x[2 * (k1 + n1 * k2)] = res[2 * k2];
                x[2 * (k1 + n1 * k2) + 1] = res[2 * k2 + 1];
            }
        }
    }
    free(col); free(res); free(work); free(tmp);
    return 0;
}
