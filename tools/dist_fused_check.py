#!/usr/bin/env python3
"""Run under torchrun on R GPUs: the fused distributed Four-Step (vkfft_b200.dist.FusedDistributedFFT1D: peer-window
loads/stores inside the FFT launches) -- parity against the single-GPU engine, timing, and the NCCL-collective
version (DistributedFFT1D) timed beside it."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import vkfft_b200 as vk
from vkfft_b200.dist import DistributedFFT1D, FusedDistributedFFT1D

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
sizes = [int(a) for a in sys.argv[1:]] or [21, 24, 26]


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / reps], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms.item()


for logn in sizes:
    n = 1 << logn
    g = torch.Generator(device=dev).manual_seed(7)
    full = torch.view_as_complex(torch.empty(n, 2, dtype=torch.float32, device=dev).uniform_(-1, 1, generator=g))
    lo, hi = rank * n // world, (rank + 1) * n // world
    f = FusedDistributedFFT1D(n, dist, local, normalize=True)
    f.local.copy_(full[lo:hi])
    torch.cuda.synchronize(); dist.barrier()
    f(inverse=False)
    f.check()
    y = f.local.clone()
    ref = full.clone()
    app = vk.VkFFTApplication()
    assert vk.initializeVkFFT(app, vk.VkFFTConfiguration(FFTdim=1, size=[n], device=local)) == 0
    assert vk.VkFFTAppend(app, -1, vk.VkFFTLaunchParams(buffer=ref)) == 0
    torch.cuda.synchronize()
    vk.deleteVkFFT(app)
    mine = ref[lo:hi]
    err = ((y - mine).abs().double().norm() / mine.abs().double().norm()).item()
    f(inverse=True)
    f.check()
    back = ((f.local - full[lo:hi]).abs().double().norm() / full[lo:hi].abs().double().norm()).item()
    errs = torch.tensor([err, back], device=dev, dtype=torch.float64)
    dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    del ref, mine, y

    def pair():
        f(inverse=False)
        f(inverse=True)
    ms_pair = timed(pair)
    f.check()
    dist.barrier()
    f.timed(False)
    breakdown = f.timed(False)
    info = vk.planInfo(f.app)
    f.close()
    # the collective version, forward only
    n1 = 1 << ((logn + 1) // 2); n2 = 1 << (logn // 2)
    ms_nccl = None
    if not os.environ.get("DIST_SKIP_NCCL"):
        slab = full[lo:hi].clone()
        c = DistributedFFT1D(n1, n2, dist, device=dev)
        ms_nccl = timed(lambda: c(slab), reps=5, warm=2)
        c.close()
    if rank == 0:
        ms = ms_pair / 2
        print(json.dumps({"n": f"2^{logn}", "world": world, "rel_err_vs_single_gpu_engine": errs[0].item(),
                          "roundtrip_rel_err": errs[1].item(), "fused_ms_per_transform": ms,
                          "fused_gflops": 5 * n * logn / (ms * 1e-3) / 1e9, "nccl_collective_ms_per_transform": ms_nccl,
                          "launches_per_transform": info["num_passes_forward"],
                          "rank0_breakdown_ms": breakdown}), flush=True)
        print(info["forward"], flush=True)
dist.destroy_process_group()
