#!/usr/bin/env python3
"""Run one plan a few times (for ncu captures): python tools/run_one.py N [batch_log2_total=28] [iters=3] [inverse=0] [double=0] [half=0]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vkfft_b200 as vk

n = int(sys.argv[1]); tot = int(sys.argv[2]) if len(sys.argv) > 2 else 28
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
inv = int(sys.argv[4]) if len(sys.argv) > 4 else 0
dbl = int(sys.argv[5]) if len(sys.argv) > 5 else 0
half = int(sys.argv[6]) if len(sys.argv) > 6 else 0
pts = 1 << tot
if half:
    buf = torch.zeros(pts, dtype=torch.int32, device="cuda")            # (half re, half im) per element
else:
    buf = torch.zeros(pts, dtype=torch.complex128 if dbl else torch.complex64, device="cuda")
    torch.view_as_real(buf).uniform_(-1, 1)
app = vk.VkFFTApplication()
rc = vk.initializeVkFFT(app, vk.VkFFTConfiguration(FFTdim=1, size=[n], numberBatches=pts // n, device=0, doublePrecision=dbl, halfPrecision=half))
assert rc == 0, vk.getVkFFTErrorString(rc)
print(vk.planInfo(app)["forward"])
lp = vk.VkFFTLaunchParams(buffer=buf)
for _ in range(iters):
    vk.VkFFTAppend(app, 1 if inv else -1, lp)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(iters):
    vk.VkFFTAppend(app, 1 if inv else -1, lp)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / iters
print(f"n={n} ms={ms:.4f} alg GB/s={2*buf.element_size()*pts/ms/1e6:.1f}")
vk.deleteVkFFT(app)
