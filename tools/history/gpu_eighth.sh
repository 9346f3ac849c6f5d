#!/bin/bash
mkdir -p gpurun_out
python tools/ktune.py 0 512,1024,2048,4096,8192,16384 > gpurun_out/ktune_f32_d.log 2>&1; grep -E "kind=[012] .*v=[0123] " gpurun_out/ktune_f32_d.log | grep -v PIPE3 | cut -c1-125
python tools/ktune.py 1 128,256,512,1024,2048,4096,8192 > gpurun_out/ktune_f64_b.log 2>&1; cut -c1-125 gpurun_out/ktune_f64_b.log | grep kind
