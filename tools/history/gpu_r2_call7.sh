#!/bin/bash
# ncu evidence for the kernels added in the last session: a plan-time instantiated template (N = 1100) and the half-storage 4096-point kernel
mkdir -p gpurun_out
B200FFT_JIT_LINEINFO=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:b2_jit -s 3 -c 1 -o gpurun_out/prof_jit_1100 \
    python tools/run_one.py 1100 26 3 > gpurun_out/ncu_full_jit_1100.log 2>&1; echo "ncu jit 1100 rc=$?"
B200FFT_JIT_LINEINFO=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:b2_jit -s 3 -c 1 -o gpurun_out/prof_half_4096 \
    python tools/run_one.py 4096 27 3 0 0 1 > gpurun_out/ncu_full_half_4096.log 2>&1; echo "ncu half 4096 rc=$?"
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "known_answer or api_errors" 2>&1 | tail -n 2
python -c "import __graft_entry__ as g; g.smoke()"
