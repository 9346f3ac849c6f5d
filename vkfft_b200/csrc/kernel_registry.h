// Registry of ahead-of-time compiled kernel instantiations.  The planner looks kernels up by
// (kind, precision, length, direction, fused-operator set); nothing is generated or compiled at run time
// (the reference JIT-compiles every plan through NVRTC, vkFFT_CompileKernel.h:299-491).
#pragma once
#include "pass_params.h"

#ifdef __cplusplus
extern "C" {
#endif

enum {
    B2_KIND_ROWS = 0,       // lines contiguous in HBM on both sides
    B2_KIND_ROWS_TOUT = 1,  // contiguous lines in, neighbouring lines interleaved on the way out (four-step final pass)
    B2_KIND_COLS = 2,       // neighbouring lines interleaved (strided axis / four-step first pass)
    B2_KIND_GENERIC = 3,    // runtime-scheduled kernel (generic.cuh): n = 0 in the registry, any addressing
    B2_KIND_ELEMENTWISE = 4,// elementwise helper passes (ew.cuh)
    B2_KIND_COUNT
};
enum { B2_PREC_F32 = 0, B2_PREC_F64 = 1 };

typedef struct b2_kernel_info {
    int kind, prec, n, inv, ops;       // lookup key
    int variant;                       // 0 = default; further CTA shapes / schedules for the same key (tuning)
    int pipelined;                     // 1: persistent TMA-fed kernel (pipe.cuh): input pointer and line pitch must be 16-byte aligned
    int threads, q, tpl, v, smem_bytes;
    int ns;
    int radices[8];
    int lut_size;                      // complex entries of the stage-twiddle LUT
    // enqueue `grid` CTAs on `stream` (cudaStream_t); returns cudaError_t as int
    int (*launch)(const b2_pass_params* P, unsigned grid, void* stream);
    int (*prepare)(void);              // one-time cudaFuncSetAttribute (max dynamic smem)
    const char* name;
} b2_kernel_info;

void b2_register_kernel(const b2_kernel_info* k);
const b2_kernel_info* b2_find_kernel(int kind, int prec, int n, int inv, int ops);   // honours B200FFT_VARIANTS
const b2_kernel_info* b2_find_kernel_variant(int kind, int prec, int n, int inv, int ops, int variant);
int b2_kernel_count(void);
const b2_kernel_info* b2_kernel_at(int i);

// fused Four-Step kernels (fused4.cuh): both passes of n1 x n2 in one persistent launch
typedef struct b2_fused_info {
    int prec, n1, n2, inv;             // lookup key
    int variant;
    int threads, qa, qb, smem_bytes;   // CTA shape, columns per pass-A tile, rows per pass-B tile
    int regs;                          // register budget per thread (bounds the resident CTAs per SM)
    int ns_a, ns_b;
    int radices_a[8], radices_b[8];
    // enqueue the control-block reset + the persistent kernel (at most max_ctas CTAs; 0 = as many as are resident)
    int (*launch)(const b2_fused_params* F, unsigned max_ctas, void* stream);
    int (*prepare)(void);
    const char* name;
} b2_fused_info;
void b2_register_fused(const b2_fused_info* k);
const b2_fused_info* b2_find_fused(int prec, int n1, int n2, int inv);
const b2_fused_info* b2_find_fused_variant(int prec, int n1, int n2, int inv, int variant);
int b2_fused_count(void);
const b2_fused_info* b2_fused_at(int i);

#ifdef __cplusplus
}
#endif
