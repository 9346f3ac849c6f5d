// POD description of ONE kernel launch ("pass") of the engine: a set of FFT lines, how they are addressed
// in HBM on the way in and on the way out, and which pre/post operators are fused into the pass.
// Plain C so the host planner (g++), the CUDA kernels (nvcc) and the CPU emulation tests agree on it.
//
// Replaces the per-axis push-constant / specialization-constant state the reference keeps in
// VkFFTSpecializationConstantsLayout (vkFFT_Structs.h:719-1014) and the stride bookkeeping of
// VkFFTPlanAxis (vkFFT_Plan_FFT.h:252-417).
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    B2_MAX_OUTER = 3,
};

// fused operators (bit flags)
enum {
    B2_OP_NONE = 0,
    B2_OP_TWIDDLE_OUT = 1,   // four-step phase  W_M^(line*elem) applied on store  (vkFFT_4step.h:31-119)
    B2_OP_SCALE = 2,         // multiply by `scale` on store (normalize=1, vkFFT_Structs.h:220)
};

typedef struct b2_pass_params {
    const void* in;
    void* out;
    const void* lut;      // per-stage radix twiddles (complex T), layout documented in lut.h
    const void* tw_hi;    // two-level four-step table: W_M^(hi << tw_shift)
    const void* tw_lo;    // W_M^lo , lo < 2^tw_shift
    // auxiliary tables used by fused real-transform operators (R2C split, DCT phases, Bluestein chirps)
    const void* aux0;
    const void* aux1;

    int64_t in_es, out_es;                 // element stride inside a line (complex elements)
    int64_t in_gs, out_gs;                 // stride between neighbouring lines of the grouped dimension
    int64_t in_bs[B2_MAX_OUTER];           // strides of the outer (batch) dimensions
    int64_t out_bs[B2_MAX_OUTER];
    uint32_t nb[B2_MAX_OUTER];             // extents of the outer dimensions (>=1)
    uint32_t G;                            // number of lines along the grouped dimension
    uint32_t n;                            // FFT length of this pass
    uint32_t tw_shift;                     // log2(size of tw_lo)
    uint32_t tw_line0;                     // offset added to the line index before forming line*elem
    uint32_t ops;                          // B2_OP_* flags
    uint32_t inverse;                      // 1: swap re/im on load+store (inverse transform)
    uint32_t aux_u0, aux_u1;               // operator specific (e.g. logical real length)
    double scale;
} b2_pass_params;

#ifdef __cplusplus
}
#endif
