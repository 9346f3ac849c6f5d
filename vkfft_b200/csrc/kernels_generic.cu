// The runtime-scheduled kernel (generic.cuh), one instantiation per precision.
#define B2_SHARD 1000
#include "kernel_inst.cuh"

static ::b200fft::GenericRegistrar<float> b2_generic_f32("generic<float>");
static ::b200fft::GenericRegistrar<double> b2_generic_f64("generic<double>");
static ::b200fft::ElementwiseRegistrar<float> b2_ew_f32("elementwise<float>");
static ::b200fft::ElementwiseRegistrar<double> b2_ew_f64("elementwise<double>");
