#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 ) > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
echo "rc=$?"; tail -c 1500 gpurun_out/bench_2gpu.json | cut -c1-1500; tail -5 gpurun_out/bench_2gpu.err
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 --no-ref-gpu ) > gpurun_out/bench_2gpu_ref.json 2> gpurun_out/bench_2gpu_ref.err
echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_2gpu_ref.json; tail -3 gpurun_out/bench_2gpu_ref.err
