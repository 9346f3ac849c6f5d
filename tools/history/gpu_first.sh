#!/bin/bash
# first GPU contact: smoke, parity tests, bench, launch list, one full ncu capture of the dominant kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" 
tail -3 gpurun_out/smoke.log
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log
( time timeout 900 python bench.py --steps 5 --warmup 3 ) > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"
tail -c 3000 gpurun_out/bench.json
tail -5 gpurun_out/bench.err
python tests/golden/make_golden.py gpurun_out/golden > gpurun_out/golden.log 2>&1
echo "golden rc=$?"; tail -3 gpurun_out/golden.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-ref-gpu --no-cpu --e2e-steps 1 > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stockham -s 3 -c 2 -o gpurun_out/prof_n4096 \
    python tools/run_one.py 4096 28 3 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
