#!/bin/bash
# plan-time instantiated templates: parity on the GPU, A/B against the runtime-scheduled kernel and the reference
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_jit.py -m gpu -x -q 2>&1 | tail -n 6
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "non_pow2 or r2c or dct or dst or nd_vs" 2>&1 | tail -n 4
timeout 900 python tools/bench_bluestein.py > gpurun_out/bench_other_lengths_ab.log 2>&1; echo "ab rc=$?"; cat gpurun_out/bench_other_lengths_ab.log
