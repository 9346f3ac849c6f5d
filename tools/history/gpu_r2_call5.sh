#!/bin/bash
# the GPU tests that come after the first failure of the final run (-x): half storage (memory-only, timing) and the plan-time kernels
timeout 900 python -m pytest tests/test_half_storage.py tests/test_jit.py -m gpu -q -s 2>&1 | tail -n 8 | tee gpurun_out/pytest_gpu_half_jit.log
