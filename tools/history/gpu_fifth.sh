#!/bin/bash
mkdir -p gpurun_out
python tools/ktune.py 0 256,512,1024,2048,4096,8192,16384 > gpurun_out/ktune_f32_c.log 2>&1; grep -E "kind=[01] " gpurun_out/ktune_f32_c.log | cut -c1-150
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "c2c_1d_f32" 2>&1 | tail -3
