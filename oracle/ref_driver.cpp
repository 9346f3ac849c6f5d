// TEST / BASELINE INFRASTRUCTURE -- not part of the product; nothing under vkfft_b200/ links or loads this.
//
// Thin C wrapper around the UNMODIFIED reference (DTolm/VkFFT, CUDA backend, VKFFT_BACKEND=1).  It is compiled
// by oracle/Makefile straight from the reference headers where they lie (-I/root/reference/vkFFT, nothing is
// copied into this repository) into oracle/_ref/libvkfft_ref.so, which travels to the GPU box.  Uses:
//   * parity oracle: same input -> reference CUDA backend output (tests/test_gpu_parity_ref.py);
//   * GPU baseline: the reference's own FFT+iFFT timing loop (performVulkanFFTiFFT,
//     benchmark_scripts/vkFFT_scripts/src/utils_VkFFT.cpp:868-919) for bench.py.
// Needs libcuda, libnvrtc and a GPU at run time (the reference JIT-compiles its kernels with NVRTC).
#include "vkFFT.h"   // the reference's umbrella header

#include <chrono>
#include <cstdio>

#include "../include/b200fft.h"   // only for the plain-C plan description struct shared with the tests

namespace {
struct Handle {
    VkFFTApplication app;
    VkFFTConfiguration cfg;
    CUdevice dev;
    void* buffer;
    void* input;
    void* output;
    pfUINT bufferSize;
};

int fill(const b200fft_desc* d, Handle* h) {
    memset(&h->app, 0, sizeof h->app);
    memset(&h->cfg, 0, sizeof h->cfg);
    if (cuInit(0) != CUDA_SUCCESS) return VKFFT_ERROR_FAILED_TO_INITIALIZE;
    if (cudaSetDevice(d->device) != cudaSuccess) return VKFFT_ERROR_FAILED_TO_SET_DEVICE_ID;
    cudaFree(0);
    if (cuDeviceGet(&h->dev, d->device) != CUDA_SUCCESS) return VKFFT_ERROR_FAILED_TO_GET_DEVICE;
    VkFFTConfiguration& c = h->cfg;
    c.FFTdim = d->fft_dim;
    for (int i = 0; i < 4; ++i) {
        c.size[i] = d->size[i];
        c.bufferStride[i] = d->buffer_stride[i];
        c.inputBufferStride[i] = d->input_stride[i];
        c.outputBufferStride[i] = d->output_stride[i];
        c.omitDimension[i] = d->omit_dimension[i];
    }
    c.numberBatches = d->number_batches;
    c.coordinateFeatures = d->coordinate_features;
    c.doublePrecision = d->precision == B200FFT_F64;
    c.performR2C = d->perform_r2c;
    c.performDCT = d->perform_dct;
    c.performDST = d->perform_dst;
    c.normalize = d->normalize;
    c.disableReorderFourStep = d->disable_reorder_four_step;
    c.makeForwardPlanOnly = d->make_forward_plan_only;
    c.makeInversePlanOnly = d->make_inverse_plan_only;
    c.isInputFormatted = d->is_input_formatted;
    c.isOutputFormatted = d->is_output_formatted;
    c.inverseReturnToInputBuffer = d->inverse_return_to_input;
    c.useLUT = (pfINT)d->reserved[0];           // 0 auto (reference default: FP32 computes sincos on chip), 1 LUT
    c.device = &h->dev;
    return VKFFT_SUCCESS;
}
}  // namespace

extern "C" int vkref_version() { return VkFFTGetVersion(); }

extern "C" int vkref_open(const b200fft_desc* d, void** out) {
    Handle* h = new Handle();
    int rc = fill(d, h);
    if (rc == VKFFT_SUCCESS) rc = initializeVkFFT(&h->app, h->cfg);
    if (rc != VKFFT_SUCCESS) { delete h; *out = nullptr; return rc; }
    *out = h;
    return rc;
}

extern "C" int vkref_append(void* hv, int inverse, void* buffer, void* input, void* output) {
    Handle* h = (Handle*)hv;
    VkFFTLaunchParams lp = {};
    h->buffer = buffer; h->input = input; h->output = output;
    lp.buffer = &h->buffer;
    if (input) lp.inputBuffer = &h->input;
    if (output) lp.outputBuffer = &h->output;
    return VkFFTAppend(&h->app, inverse, &lp);
}

extern "C" void vkref_close(void* hv) {
    Handle* h = (Handle*)hv;
    if (!h) return;
    deleteVkFFT(&h->app);
    delete h;
}

// number of kernel launches ("uploads") the reference plan uses for axis 0 -- for reporting
extern "C" int vkref_axis0_uploads(void* hv) {
    Handle* h = (Handle*)hv;
    return h->app.localFFTPlan ? (int)h->app.localFFTPlan->numAxisUploads[0] : -1;
}

// one-shot: plan, run once, synchronise, delete
extern "C" int vkref_run(const b200fft_desc* d, int inverse, void* buffer, void* input, void* output) {
    void* h = nullptr;
    int rc = vkref_open(d, &h);
    if (rc != VKFFT_SUCCESS) return rc;
    rc = vkref_append(h, inverse, buffer, input, output);
    if (cudaDeviceSynchronize() != cudaSuccess && rc == VKFFT_SUCCESS) rc = VKFFT_ERROR_FAILED_TO_SYNCHRONIZE;
    vkref_close(h);
    return rc;
}

// The reference's benchmark loop: `iters` x (forward, inverse) back to back, one synchronise, wall clock AND
// CUDA events.  Returns ms per (FFT+iFFT) pair in ms_event / ms_wall.
extern "C" int vkref_bench_pairs(void* hv, void* buffer, int warmup, int iters, double* ms_event, double* ms_wall) {
    Handle* h = (Handle*)hv;
    int rc = VKFFT_SUCCESS;
    for (int i = 0; i < warmup && rc == VKFFT_SUCCESS; ++i) {
        rc = vkref_append(h, -1, buffer, nullptr, nullptr);
        if (rc == VKFFT_SUCCESS) rc = vkref_append(h, 1, buffer, nullptr, nullptr);
    }
    if (cudaDeviceSynchronize() != cudaSuccess) return VKFFT_ERROR_FAILED_TO_SYNCHRONIZE;
    if (rc != VKFFT_SUCCESS) return rc;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto t0 = std::chrono::steady_clock::now();
    cudaEventRecord(e0, 0);
    for (int i = 0; i < iters && rc == VKFFT_SUCCESS; ++i) {
        rc = vkref_append(h, -1, buffer, nullptr, nullptr);
        if (rc == VKFFT_SUCCESS) rc = vkref_append(h, 1, buffer, nullptr, nullptr);
    }
    cudaEventRecord(e1, 0);
    if (cudaDeviceSynchronize() != cudaSuccess) rc = VKFFT_ERROR_FAILED_TO_SYNCHRONIZE;
    auto t1 = std::chrono::steady_clock::now();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (ms_event) *ms_event = (double)ms / iters;
    if (ms_wall) *ms_wall = std::chrono::duration<double, std::milli>(t1 - t0).count() / iters;
    return rc;
}
