#!/bin/bash
mkdir -p gpurun_out
echo "== COLS without phase multiply"; KTUNE_COLS_PLAIN=1 python tools/ktune.py 0 16,32,64,128,256,512,1024 2>&1 | grep "kind=2"
echo "== COLS+TW other=1040 (non power-of-two row stride)"; KTUNE_OTHER=1040 python tools/ktune.py 0 16,64,128,256,1024 2>&1 | grep "kind=2"
echo "== COLS+TW other=4096"; KTUNE_OTHER=4096 python tools/ktune.py 0 16,64,128,256 2>&1 | grep "kind=2"
echo "== COLS+TW other=256"; KTUNE_OTHER=256 python tools/ktune.py 0 16,64,128,256,1024 2>&1 | grep "kind=2"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stockham -s 4 -c 2 -o gpurun_out/prof_fourstep_65536 \
    python tools/run_one.py 65536 28 3 > gpurun_out/ncu_fourstep.log 2>&1
echo "ncu rc=$?"
