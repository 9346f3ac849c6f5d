// TEST INFRASTRUCTURE ONLY: instantiates part B2_EMU_PART of B2_EMU_PARTS of the kernel lists for the CPU emulation.
#include "kernel_inst.cuh"
#include "kernel_list.def"
#include "kernel_list_nonpow2.def"
#include "kernel_list_blue1.def"
#include "kernel_list_fused.def"
// test-only: half-precision storage variants (the product instantiates them at plan time, vkfft_b200/csrc/jit.cpp)
B2_KH(0, ROWS, 0, 8, 16, 1, 64, 8, 8)                     // 64
B2_KH(1, ROWS, 0, 10, 12, 1, 80, 10, 10)                  // 100
B2_KH(2, ROWS, 0, 32, 4, 1, 128, 32, 32)                  // 1024
B2_KH(3, ROWS, 0, 1, 64, 1, 64, 8)                        // 8 (one radix, no shared memory)
B2_KH(4, COLS, 0, 8, 16, 1, 64, 8, 8)                     // 64, strided axis
B2_KH(5, COLS, B2_OP_TWIDDLE_OUT, 8, 16, 1, 64, 8, 8)     // 64, Four-Step first launch
B2_KH(6, ROWS_TOUT, 0, 8, 16, 1, 64, 8, 8)                // 64, Four-Step last launch
