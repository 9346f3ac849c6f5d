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

    // convolution, the call sequence of the reference's sample_50_convolution_VkFFT_single_1d_matrix.cpp:100-330:
    // a kernel application transforms the kernel (here: a delta shifted by 5), the convolution application then does
    // FFT -> product -> iFFT in one VkFFTAppend.  Convolving with a shifted delta is a circular shift.
    const uint64_t M = 4096, C = 2, SH = 5;
    std::vector<float> hk(2 * M * C, 0.f), hx(2 * M * C), hy(2 * M * C);
    for (uint64_t c = 0; c < C; c++) hk[2 * (c * M + SH)] = 1.f;
    for (uint64_t i = 0; i < M * C; i++) { hx[2 * i] = (float)(i % 8) - 3.5f; hx[2 * i + 1] = (float)(i % 4) - 1.5f; }
    void *kernel = 0, *cbuf = 0;
    if (cudaMalloc(&kernel, sizeof(float) * 2 * M * C) != cudaSuccess || cudaMalloc(&cbuf, sizeof(float) * 2 * M * C) != cudaSuccess) return 2;
    cudaMemcpy(kernel, hk.data(), sizeof(float) * 2 * M * C, cudaMemcpyHostToDevice);
    cudaMemcpy(cbuf, hx.data(), sizeof(float) * 2 * M * C, cudaMemcpyHostToDevice);
    VkFFTConfiguration kcfg = {};
    VkFFTApplication app_kernel = {}, app_conv = {};
    kcfg.FFTdim = 1; kcfg.size[0] = M; kcfg.coordinateFeatures = C; kcfg.kernelConvolution = 1; kcfg.normalize = 1;
    kcfg.device = &dev; kcfg.buffer = &kernel;
    if ((res = initializeVkFFT(&app_kernel, kcfg)) != VKFFT_SUCCESS) { printf("kernel init: %s\n", getVkFFTErrorString(res)); return 1; }
    if ((res = VkFFTAppend(&app_kernel, -1, 0)) != VKFFT_SUCCESS) { printf("kernel fft: %s\n", getVkFFTErrorString(res)); return 1; }
    VkFFTConfiguration ccfg = kcfg;
    ccfg.kernelConvolution = 0; ccfg.performConvolution = 1; ccfg.buffer = &cbuf; ccfg.kernel = &kernel;
    if ((res = initializeVkFFT(&app_conv, ccfg)) != VKFFT_SUCCESS) { printf("conv init: %s\n", getVkFFTErrorString(res)); return 1; }
    if ((res = VkFFTAppend(&app_conv, -1, 0)) != VKFFT_SUCCESS) { printf("conv: %s\n", getVkFFTErrorString(res)); return 1; }
    cudaDeviceSynchronize();
    cudaMemcpy(hy.data(), cbuf, sizeof(float) * 2 * M * C, cudaMemcpyDeviceToHost);
    double worst_conv = 0;
    for (uint64_t c = 0; c < C; c++)
        for (uint64_t i = 0; i < M; i++) {
            const uint64_t src = c * M + (i + M - SH) % M;
            worst_conv = fmax(worst_conv, hypot(hy[2 * (c * M + i)] - hx[2 * src], hy[2 * (c * M + i) + 1] - hx[2 * src + 1]));
        }
    deleteVkFFT(&app_kernel);
    deleteVkFFT(&app_conv);
    cudaFree(kernel);
    cudaFree(cbuf);
    // several streams (num_streams = 3): work enqueued on stream[2] BEFORE the transform (the fill) and on stream[1] AFTER it
    // (the copy back) must be ordered with the transform without any host synchronisation in between
    double worst_ms = 0;
    {
        const uint64_t L = 1 << 12, LB = 64;
        cudaStream_t st[3];
        for (int i = 0; i < 3; i++) cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking);
        void* mb = 0;
        float* pinned = 0;
        if (cudaMalloc(&mb, sizeof(float) * 2 * L * LB) != cudaSuccess || cudaMallocHost((void**)&pinned, sizeof(float) * 2 * L * LB) != cudaSuccess) return 2;
        for (uint64_t i = 0; i < 2 * L * LB; i++) pinned[i] = 0.f;
        for (uint64_t b = 0; b < LB; b++) pinned[2 * (b * L + 1)] = 1.f;        // delta at 1 -> exp(-2 pi i k / L)
        VkFFTConfiguration mcfg = {};
        VkFFTApplication mapp = {};
        mcfg.FFTdim = 1; mcfg.size[0] = L; mcfg.numberBatches = LB; mcfg.device = &dev; mcfg.stream = st; mcfg.num_streams = 3;
        if ((res = initializeVkFFT(&mapp, mcfg)) != VKFFT_SUCCESS) { printf("multi-stream init: %s\n", getVkFFTErrorString(res)); return 1; }
        VkFFTLaunchParams mlp = {};
        mlp.buffer = &mb;
        cudaMemcpyAsync(mb, pinned, sizeof(float) * 2 * L * LB, cudaMemcpyHostToDevice, st[2]);
        if ((res = VkFFTAppend(&mapp, -1, &mlp)) != VKFFT_SUCCESS) { printf("multi-stream append: %s\n", getVkFFTErrorString(res)); return 1; }
        cudaMemcpyAsync(pinned, mb, sizeof(float) * 2 * L * LB, cudaMemcpyDeviceToHost, st[1]);
        cudaStreamSynchronize(st[1]);
        for (uint64_t b = 0; b < LB; b++)
            for (uint64_t k = 0; k < L; k++) {
                const double a = -2.0 * M_PI * (double)k / (double)L;
                worst_ms = fmax(worst_ms, hypot(pinned[2 * (b * L + k)] - cos(a), pinned[2 * (b * L + k) + 1] - sin(a)));
            }
        deleteVkFFT(&mapp);
        cudaFree(mb); cudaFreeHost(pinned);
        for (int i = 0; i < 3; i++) cudaStreamDestroy(st[i]);
    }
    printf("forward max abs err %.3e, inverse max rel err %.3e, convolution max abs err %.3e, multi-stream max abs err %.3e, version %d\n",
           worst, worst_inv, worst_conv, worst_ms, VkFFTGetVersion());
    return (worst < 2e-6 && worst_inv < 2e-6 && worst_conv < 2e-5 && worst_ms < 2e-6 && app.b200fftPlan == 0) ? 0 : 1;
}
