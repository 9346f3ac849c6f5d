#!/usr/bin/env python3
"""How fast are the Four-Step launches when ALL their data is L2-resident?  (upper bound for a fused, L2-pinned
scheme: if this is not much faster than the HBM-streaming case the kernels are SM-bound and fusing cannot help)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vkfft_b200 as vk

for logn in (16, 18, 20, 22):
    n = 1 << logn
    for logtot in (22, 23, 28):                      # 32 MB, 64 MB (L2 resident incl. temp) and 2 GiB (HBM)
        pts = 1 << logtot
        if pts < n:
            continue
        buf = torch.zeros(pts, dtype=torch.complex64, device="cuda")
        torch.view_as_real(buf).uniform_(-1, 1)
        app = vk.VkFFTApplication()
        assert vk.initializeVkFFT(app, vk.VkFFTConfiguration(FFTdim=1, size=[n], numberBatches=pts // n, device=0)) == 0
        lp = vk.VkFFTLaunchParams(buffer=buf)
        reps = max(4, (1 << 29) // pts // 4)
        for _ in range(3):
            vk.VkFFTAppend(app, -1, lp)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            vk.VkFFTAppend(app, -1, lp)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        print(f"N=2^{logn} buffer={pts*8>>20:5d} MB  {ms*1e3:8.1f} us per transform -> {ms*1e3*(1<<28)/pts:8.1f} us per 2 GiB-equivalent", flush=True)
        vk.deleteVkFFT(app)
        del buf
