#!/usr/bin/env python3
"""Run under torchrun on R GPUs: distributed single-sequence Four-Step (vkfft_b200.dist.DistributedFFT1D) --
parity against the single-GPU engine and timing for N = 2^26 (BASELINE.json config 5)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import vkfft_b200 as vk
from vkfft_b200.dist import DistributedFFT1D

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
out = {}
for logn in (16, 22, 26):
    n1 = 1 << ((logn + 1) // 2); n2 = 1 << (logn // 2); n = n1 * n2
    g = torch.Generator(device=dev).manual_seed(7)
    full = torch.empty(n, 2, dtype=torch.float32, device=dev).uniform_(-1, 1, generator=g)   # same on every rank (same seed)
    full = torch.view_as_complex(full)
    slab = full[rank * n // world:(rank + 1) * n // world].clone()
    f = DistributedFFT1D(n1, n2, dist, device=dev)
    y = f(slab)
    torch.cuda.synchronize()
    # reference: the whole sequence on this GPU through the engine
    ref = full.clone()
    app = vk.VkFFTApplication()
    assert vk.initializeVkFFT(app, vk.VkFFTConfiguration(FFTdim=1, size=[n], device=local)) == 0
    assert vk.VkFFTAppend(app, -1, vk.VkFFTLaunchParams(buffer=ref)) == 0
    torch.cuda.synchronize()
    vk.deleteVkFFT(app)
    mine = ref[rank * n // world:(rank + 1) * n // world]
    err = ((y - mine).abs().double().norm() / mine.abs().double().norm()).item()
    # timing
    for _ in range(2):
        f(slab)
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    reps = 5
    for _ in range(reps):
        f(slab)
    b.record(); torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / reps], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    errs = torch.tensor([err], device=dev, dtype=torch.float64)
    dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    f.close()
    if rank == 0:
        out[f"2^{logn}"] = {"rel_err_vs_single_gpu_engine": errs.item(), "ms": ms.item(), "gflops": 5 * n * logn / (ms.item() * 1e-3) / 1e9}
        print(json.dumps({f"2^{logn}": out[f"2^{logn}"], "world": world}), flush=True)
dist.destroy_process_group()
