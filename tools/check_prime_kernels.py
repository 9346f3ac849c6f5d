#!/usr/bin/env python3
"""GPU spot check of the {17..31} x 2^k specialised kernels (and the N = 2 real-transform edge cases): parity vs torch.fft, one timing."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vkfft_b200 as vk
from vkfft_b200 import fft as vkfft

out = {}
for n in (136, 248, 496, 992, 1088, 1984, 2176, 3968):
    x = torch.randn(64, n, dtype=torch.complex64, device="cuda")
    y = vkfft.fftn(x, ndim=1, norm=0)
    ref = torch.fft.fft(x.to(torch.complex128), dim=-1)
    out[n] = float(((y - ref).abs().norm() / ref.abs().norm()).item())
r = torch.randn(5, 2 * 1088, dtype=torch.float32, device="cuda")
out["r2c_2176"] = float(((vkfft.rfftn(r, ndim=1, norm=0) - torch.fft.rfft(r.double(), dim=-1)).abs().norm() / torch.fft.rfft(r.double(), dim=-1).abs().norm()).item())
r2 = torch.randn(7, 2, dtype=torch.float32, device="cuda")
out["r2c_n2"] = float((vkfft.rfftn(r2, ndim=1, norm=0) - torch.fft.rfft(r2.double(), dim=-1)).abs().max().item())
n, batch = 1088, 1 << 17
buf = torch.randn(batch, n, dtype=torch.complex64, device="cuda")
app = vk.VkFFTApplication()
assert vk.initializeVkFFT(app, vk.VkFFTConfiguration(FFTdim=1, size=[n], numberBatches=batch, device=0)) == 0
lp = vk.VkFFTLaunchParams(buffer=buf)
for _ in range(2):
    vk.VkFFTAppend(app, -1, lp); vk.VkFFTAppend(app, 1, lp)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    vk.VkFFTAppend(app, -1, lp); vk.VkFFTAppend(app, 1, lp)
b.record(); torch.cuda.synchronize()
out["ms_pair_1088_batch_2p17"] = a.elapsed_time(b) / 5
out["plan"] = vk.planInfo(app)["forward"].strip()
print(json.dumps(out))
