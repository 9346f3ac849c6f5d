#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python tools/ktune.py 0 ) > gpurun_out/ktune_f32.log 2>&1
echo "ktune rc=$?"; cat gpurun_out/ktune_f32.log | tail -80
( time timeout 600 python tools/l2_chunk_exp.py ) > gpurun_out/l2_chunk.log 2>&1
echo "l2 rc=$?"; cat gpurun_out/l2_chunk.log | tail -12
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu2.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu2.log
