#!/usr/bin/env python3
"""Experiment: Four-Step with the intermediate kept L2-resident.  The 2 GiB batch is processed in chunks small
enough that pass 1's output (written to a small, reused temp buffer) is still in the 126 MB L2 when pass 2 reads
it.  All launches are captured in one CUDA graph so launch overhead does not mask the effect."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vkfft_b200 as vk

pts = 1 << 28
buf = torch.zeros(pts, dtype=torch.complex64, device="cuda")
torch.view_as_real(buf).uniform_(-1, 1)
tmp = torch.zeros(pts, dtype=torch.complex64, device="cuda")
side = torch.cuda.Stream()


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for logn in (15, 16, 18, 20, 22):
    n = 1 << logn
    app = vk.VkFFTApplication()
    assert vk.initializeVkFFT(app, vk.VkFFTConfiguration(FFTdim=1, size=[n], numberBatches=pts // n, device=0,
                                                         userTempBuffer=1, tempBufferSize=pts * 8)) == 0
    lp = vk.VkFFTLaunchParams(buffer=buf, tempBuffer=tmp)
    base = timed(lambda: vk.VkFFTAppend(app, -1, lp))
    vk.deleteVkFFT(app)
    line = f"N=2^{logn}: unchunked {base*1e3:7.1f} us |"
    for chunk_mb in (8, 16, 32, 48, 64):
        seqs = max(1, (chunk_mb << 20) // (n * 8))
        nchunks = (pts // n) // seqs
        if nchunks < 2 or (pts // n) % seqs:
            continue
        app = vk.VkFFTApplication()
        cfg = vk.VkFFTConfiguration(FFTdim=1, size=[n], numberBatches=seqs, device=0, userTempBuffer=1,
                                    tempBufferSize=seqs * n * 8, specifyOffsetsAtLaunch=1)
        assert vk.initializeVkFFT(app, cfg) == 0
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            st = torch.cuda.current_stream().cuda_stream
            for c in range(nchunks):
                l = vk.VkFFTLaunchParams(buffer=buf, tempBuffer=tmp, bufferOffset=c * seqs * n * 8, stream=st)
                assert vk.VkFFTAppend(app, -1, l) == 0
        t = timed(g.replay)
        line += f" {chunk_mb}MB:{t*1e3:7.1f}"
        vk.deleteVkFFT(app)
        del g
    print(line, flush=True)
