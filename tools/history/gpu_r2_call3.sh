#!/bin/bash
# half-precision storage: parity on the GPU + the reference's half-precision benchmark program (sample_2) on both engines
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_half_storage.py -m gpu -x -q -s 2>&1 | tail -n 8
cd oracle/_ref
( timeout 600 ./VkFFT_TestSuite_b200 -vkfft 2 ) > ../../gpurun_out/sample2_half_b200.log 2>&1; echo "sample_2 b200 rc=$?"
( timeout 600 ./VkFFT_TestSuite_ref -vkfft 2 ) > ../../gpurun_out/sample2_half_ref.log 2>&1; echo "sample_2 ref rc=$?"
cd ../..
grep -E "Benchmark score|VkFFT System" gpurun_out/sample2_half_b200.log | tail -n 28
grep -E "Benchmark score|VkFFT System" gpurun_out/sample2_half_ref.log | tail -n 28
