// The runtime-scheduled kernel (generic.cuh), one instantiation per precision.
#define B2_SHARD 1000
#include "kernel_inst.cuh"

static ::b200fft::GenericRegistrar<float> b2_generic_f32("generic<float>");
static ::b200fft::GenericRegistrar<double> b2_generic_f64("generic<double>");
