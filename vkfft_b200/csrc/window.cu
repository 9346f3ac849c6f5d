// Peer windows: one flat virtual address range over the slabs of every GPU of a box (include/b200fft.h).
//
// The reference is single-device (README.md:26-28); this is the memory model of the distributed Four-Step
// (SURVEY.md section 8e): every rank allocates its slab with the CUDA virtual-memory API, the ranks swap POSIX
// file descriptors of those allocations, and every rank maps all slabs back to back into one reserved address
// range.  A kernel then addresses the whole distributed sequence like a local array -- the Four-Step launches of
// planner.cpp run unchanged on a slice of their lines, and their strided loads / transposed stores ARE the
// all-to-all exchange (NVLink reads and writes issued tile by tile from inside the FFT kernels; no pack kernels,
// no separate collective).  Ranks meet at a device-side barrier over a small signal pad that is mapped the same way.
//
// Driver entry points are fetched through cudaGetDriverEntryPoint so that the library keeps linking against the
// CUDA runtime only.
#include <cuda.h>
#include <cuda_runtime.h>
#include <unistd.h>

#include <cstdint>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/b200fft.h"
#include "plan.h"

using namespace b200fft;

namespace {

struct DriverApi {
    CUresult (*GetGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*Create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
    CUresult (*Release)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*AddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*AddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*Map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*Unmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*SetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
    CUresult (*Export)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
    CUresult (*Import)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
    bool ok = false;
};

template <typename F>
bool fetch(const char* name, F& fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) return false;
    fn = reinterpret_cast<F>(p);
    return true;
}

const DriverApi& drv() {
    static DriverApi d = [] {
        DriverApi a;
        a.ok = fetch("cuMemGetAllocationGranularity", a.GetGranularity) && fetch("cuMemCreate", a.Create) &&
               fetch("cuMemRelease", a.Release) && fetch("cuMemAddressReserve", a.AddressReserve) &&
               fetch("cuMemAddressFree", a.AddressFree) && fetch("cuMemMap", a.Map) && fetch("cuMemUnmap", a.Unmap) &&
               fetch("cuMemSetAccess", a.SetAccess) && fetch("cuMemExportToShareableHandle", a.Export) &&
               fetch("cuMemImportFromShareableHandle", a.Import);
        return a;
    }();
    return d;
}

CUmemAllocationProp slab_prop(int device) {
    CUmemAllocationProp p;
    memset(&p, 0, sizeof p);
    p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    p.location.id = device;
    p.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    return p;
}

struct Guard {
    int prev = -1;
    explicit Guard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); }
    ~Guard() { int cur = -1; if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev); }
};

constexpr uint64_t BARRIER_TIMEOUT_NS = 4000000000ull;

// one thread per peer: tell peer t "rank reached `epoch`" (release: everything this rank's earlier launches stored,
// to any slab, is visible first), then wait until peer t has told us the same (acquire)
__global__ void window_barrier_kernel(uint32_t* pads, uint64_t slot_words, uint32_t rank, uint32_t world, uint32_t epoch, int* err) {
    const uint32_t t = threadIdx.x;
    if (t >= world) return;
    __threadfence_system();
    uint32_t* remote = pads + (uint64_t)t * slot_words + rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(epoch) : "memory");
    const uint32_t* mine = pads + (uint64_t)rank * slot_words + t;
    uint64_t t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
        if ((int32_t)(v - epoch) >= 0) break;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > BARRIER_TIMEOUT_NS) { *err = 1; break; }
        __nanosleep(64);
    }
    __threadfence_system();
}

}  // namespace

struct b200fft_window {
    int device = 0;
    uint32_t world = 1, rank = 0;
    uint64_t slab_bytes = 0, pad_bytes = 0;
    CUdeviceptr data_va = 0, pad_va = 0;
    std::vector<CUmemGenericAllocationHandle> data_h, pad_h;
    std::vector<char> mapped;
    uint32_t epoch = 0;
    int* d_err = nullptr;            // device alias of h_err
    volatile int* h_err = nullptr;   // mapped pinned host word: a barrier that timed out sets it; read without synchronising
};

extern "C" uint64_t b200fft_window_granularity(int device) {
    const DriverApi& a = drv();
    if (!a.ok) return 0;
    Guard g(device);
    if (cudaFree(0) != cudaSuccess) { cudaGetLastError(); return 0; }
    CUmemAllocationProp p = slab_prop(device);
    size_t gran = 0;
    if (a.GetGranularity(&gran, &p, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS) return 0;
    return gran;
}

static int map_slot(b200fft_window* w, uint32_t slot, CUmemGenericAllocationHandle dh, CUmemGenericAllocationHandle ph) {
    const DriverApi& a = drv();
    if (a.Map(w->data_va + (CUdeviceptr)slot * w->slab_bytes, w->slab_bytes, 0, dh, 0) != CUDA_SUCCESS) return R_FAILED_TO_ALLOCATE;
    if (a.Map(w->pad_va + (CUdeviceptr)slot * w->pad_bytes, w->pad_bytes, 0, ph, 0) != CUDA_SUCCESS) {
        a.Unmap(w->data_va + (CUdeviceptr)slot * w->slab_bytes, w->slab_bytes);
        return R_FAILED_TO_ALLOCATE;
    }
    w->data_h[slot] = dh; w->pad_h[slot] = ph; w->mapped[slot] = 1;
    CUmemAccessDesc acc;
    memset(&acc, 0, sizeof acc);
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = w->device;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    if (a.SetAccess(w->data_va + (CUdeviceptr)slot * w->slab_bytes, w->slab_bytes, &acc, 1) != CUDA_SUCCESS) return R_FAILED_TO_ALLOCATE;
    if (a.SetAccess(w->pad_va + (CUdeviceptr)slot * w->pad_bytes, w->pad_bytes, &acc, 1) != CUDA_SUCCESS) return R_FAILED_TO_ALLOCATE;
    return R_SUCCESS;
}

extern "C" void b200fft_window_destroy(b200fft_window* w) {
    if (!w) return;
    const DriverApi& a = drv();
    Guard g(w->device);
    cudaDeviceSynchronize();
    for (uint32_t s = 0; s < w->world; ++s) {
        if (!w->mapped[s]) continue;
        a.Unmap(w->data_va + (CUdeviceptr)s * w->slab_bytes, w->slab_bytes);
        a.Unmap(w->pad_va + (CUdeviceptr)s * w->pad_bytes, w->pad_bytes);
        a.Release(w->data_h[s]);
        a.Release(w->pad_h[s]);
    }
    if (w->data_va) a.AddressFree(w->data_va, (size_t)w->world * w->slab_bytes);
    if (w->pad_va) a.AddressFree(w->pad_va, (size_t)w->world * w->pad_bytes);
    if (w->h_err) cudaFreeHost((void*)w->h_err);
    delete w;
}

extern "C" int b200fft_window_create(int device, uint32_t world, uint32_t rank, uint64_t slab_bytes, b200fft_window** out) {
    if (!out) return R_EMPTY_APP;
    *out = nullptr;
    if (world == 0 || world > 32 || rank >= world || slab_bytes == 0) return R_EMPTY_SIZE;
    const DriverApi& a = drv();
    if (!a.ok) return R_INVALID_DEVICE;
    const uint64_t gran = b200fft_window_granularity(device);
    if (gran == 0) return R_INVALID_DEVICE;
    if (slab_bytes % gran) return R_EMPTY_SIZE;
    b200fft_window* w = new (std::nothrow) b200fft_window();
    if (!w) return R_MALLOC_FAILED;
    w->device = device; w->world = world; w->rank = rank; w->slab_bytes = slab_bytes; w->pad_bytes = gran;
    w->data_h.assign(world, 0); w->pad_h.assign(world, 0); w->mapped.assign(world, 0);
    Guard g(device);
    int rc = R_SUCCESS;
    CUmemAllocationProp p = slab_prop(device);
    CUmemGenericAllocationHandle dh = 0, ph = 0;
    if (a.AddressReserve(&w->data_va, (size_t)world * slab_bytes, gran, 0, 0) != CUDA_SUCCESS) rc = R_FAILED_TO_ALLOCATE;
    if (rc == R_SUCCESS && a.AddressReserve(&w->pad_va, (size_t)world * w->pad_bytes, gran, 0, 0) != CUDA_SUCCESS) rc = R_FAILED_TO_ALLOCATE;
    if (rc == R_SUCCESS && a.Create(&dh, slab_bytes, &p, 0) != CUDA_SUCCESS) rc = R_FAILED_TO_ALLOCATE;
    if (rc == R_SUCCESS && a.Create(&ph, w->pad_bytes, &p, 0) != CUDA_SUCCESS) { a.Release(dh); rc = R_FAILED_TO_ALLOCATE; }
    if (rc == R_SUCCESS) rc = map_slot(w, rank, dh, ph);
    if (rc == R_SUCCESS) {
        void* h = nullptr;
        if (cudaHostAlloc(&h, sizeof(int), cudaHostAllocMapped) != cudaSuccess) rc = R_FAILED_TO_ALLOCATE;
        else {
            w->h_err = (volatile int*)h;
            *w->h_err = 0;
            if (cudaHostGetDevicePointer((void**)&w->d_err, h, 0) != cudaSuccess) rc = R_FAILED_TO_ALLOCATE;
        }
    }
    if (rc == R_SUCCESS &&
        (cudaMemset((void*)(w->pad_va + (CUdeviceptr)rank * w->pad_bytes), 0, w->pad_bytes) != cudaSuccess ||
         cudaDeviceSynchronize() != cudaSuccess))
        rc = R_FAILED_TO_COPY;
    if (rc != R_SUCCESS) { cudaGetLastError(); b200fft_window_destroy(w); return rc; }
    *out = w;
    return R_SUCCESS;
}

extern "C" int b200fft_window_export(b200fft_window* w, int fds[2]) {
    if (!w || !fds) return R_EMPTY_APP;
    const DriverApi& a = drv();
    Guard g(w->device);
    int fd0 = -1, fd1 = -1;
    if (a.Export(&fd0, w->data_h[w->rank], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0) != CUDA_SUCCESS) return R_FAILED_TO_ALLOCATE;
    if (a.Export(&fd1, w->pad_h[w->rank], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0) != CUDA_SUCCESS) { close(fd0); return R_FAILED_TO_ALLOCATE; }
    fds[0] = fd0; fds[1] = fd1;
    return R_SUCCESS;
}

extern "C" int b200fft_window_import(b200fft_window* w, uint32_t peer, const int fds[2]) {
    if (!w || !fds) return R_EMPTY_APP;
    if (peer >= w->world || peer == w->rank || w->mapped[peer]) return R_EMPTY_SIZE;
    const DriverApi& a = drv();
    Guard g(w->device);
    CUmemGenericAllocationHandle dh = 0, ph = 0;
    if (a.Import(&dh, (void*)(uintptr_t)fds[0], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR) != CUDA_SUCCESS) return R_FAILED_TO_ALLOCATE;
    if (a.Import(&ph, (void*)(uintptr_t)fds[1], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR) != CUDA_SUCCESS) { a.Release(dh); return R_FAILED_TO_ALLOCATE; }
    return map_slot(w, peer, dh, ph);
}

extern "C" void* b200fft_window_base(b200fft_window* w) { return w ? (void*)w->data_va : nullptr; }
extern "C" void* b200fft_window_local(b200fft_window* w) { return w ? (void*)(w->data_va + (CUdeviceptr)w->rank * w->slab_bytes) : nullptr; }

extern "C" int b200fft_window_barrier(b200fft_window* w, void* stream) {
    if (!w) return R_EMPTY_APP;
    for (uint32_t s = 0; s < w->world; ++s)
        if (!w->mapped[s]) return R_PLAN_NOT_INITIALIZED;
    // a barrier of an earlier call gave up (a peer stopped): every later exchange would consume stale data -- refuse
    if (w->h_err && *w->h_err) return R_FAILED_TO_SYNCHRONIZE;
    Guard g(w->device);
    ++w->epoch;
    uint32_t* pads = (uint32_t*)w->pad_va;
    uint64_t slot_words = w->pad_bytes / 4;
    void* args[] = {&pads, &slot_words, &w->rank, &w->world, &w->epoch, &w->d_err};
    return cudaLaunchKernel((const void*)window_barrier_kernel, dim3(1), dim3(32), args, 0, (cudaStream_t)stream) == cudaSuccess
               ? R_SUCCESS : R_FAILED_TO_LAUNCH_KERNEL;
}

extern "C" int b200fft_window_status(b200fft_window* w) {
    if (!w) return R_EMPTY_APP;
    Guard g(w->device);
    if (cudaDeviceSynchronize() != cudaSuccess) return R_FAILED_TO_SYNCHRONIZE;
    return (w->h_err && *w->h_err) ? R_FAILED_TO_SYNCHRONIZE : R_SUCCESS;
}
