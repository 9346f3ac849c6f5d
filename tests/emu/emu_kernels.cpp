// TEST INFRASTRUCTURE ONLY: instantiates part B2_EMU_PART of B2_EMU_PARTS of the kernel lists for the CPU emulation.
#include "kernel_inst.cuh"
#include "kernel_list.def"
#include "kernel_list_nonpow2.def"
#include "kernel_list_blue1.def"
#include "kernel_list_fused.def"
