// A user program written against the reference API (same calls as the reference's samples,
// e.g. sample_11_precision_VkFFT_single.cpp:208-273), compiled against include/vkFFT.h instead.
// Forward FFT of a shifted impulse must be the phase ramp exp(-2*pi*i*s*k/N); inverse must return N*delta.
#include <cmath>
#include <cstdio>
#include <vector>

#include "vkFFT.h"

int main() {
    if (cuInit(0) != CUDA_SUCCESS) { printf("no driver\n"); return 2; }
    CUdevice dev;
    if (cuDeviceGet(&dev, 0) != CUDA_SUCCESS) return 2;
    cudaSetDevice(0);
    const uint64_t N = 1 << 16, B = 4, S = 3;
    std::vector<float> h(2 * N * B, 0.f);
    for (uint64_t b = 0; b < B; b++) h[2 * (b * N + S)] = 1.f;
    void* buffer = 0;
    if (cudaMalloc(&buffer, sizeof(float) * 2 * N * B) != cudaSuccess) return 2;
    cudaMemcpy(buffer, h.data(), sizeof(float) * 2 * N * B, cudaMemcpyHostToDevice);

    VkFFTConfiguration configuration = {};
    VkFFTApplication app = {};
    configuration.FFTdim = 1;
    configuration.size[0] = N;
    configuration.numberBatches = B;
    configuration.device = &dev;
    VkFFTResult res = initializeVkFFT(&app, configuration);
    if (res != VKFFT_SUCCESS) { printf("init: %s\n", getVkFFTErrorString(res)); return 1; }
    if (initializeVkFFT(&app, configuration) != VKFFT_ERROR_NONZERO_APP_INITIALIZATION) return 1;
    VkFFTLaunchParams launchParams = {};
    launchParams.buffer = &buffer;
    res = VkFFTAppend(&app, -1, &launchParams);
    if (res != VKFFT_SUCCESS) { printf("append: %s\n", getVkFFTErrorString(res)); return 1; }
    cudaDeviceSynchronize();
    cudaMemcpy(h.data(), buffer, sizeof(float) * 2 * N * B, cudaMemcpyDeviceToHost);
    double worst = 0;
    for (uint64_t b = 0; b < B; b++)
        for (uint64_t k = 0; k < N; k++) {
            const double a = -2.0 * M_PI * (double)((S * k) % N) / (double)N;
            worst = fmax(worst, hypot(h[2 * (b * N + k)] - cos(a), h[2 * (b * N + k) + 1] - sin(a)));
        }
    res = VkFFTAppend(&app, 1, &launchParams);
    cudaDeviceSynchronize();
    cudaMemcpy(h.data(), buffer, sizeof(float) * 2 * N * B, cudaMemcpyDeviceToHost);
    double worst_inv = 0;
    for (uint64_t b = 0; b < B; b++)
        for (uint64_t k = 0; k < N; k++)
            worst_inv = fmax(worst_inv, hypot(h[2 * (b * N + k)] - (k == S ? (double)N : 0.0), h[2 * (b * N + k) + 1]) / (double)N);
    deleteVkFFT(&app);
    cudaFree(buffer);
    printf("forward max abs err %.3e, inverse max rel err %.3e, version %d\n", worst, worst_inv, VkFFTGetVersion());
    return (worst < 2e-6 && worst_inv < 2e-6 && app.b200fftPlan == 0) ? 0 : 1;
}
