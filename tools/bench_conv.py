#!/usr/bin/env python3
"""Convolution (performConvolution) on a 2 GiB buffer: plan with the fused launch vs the chain with a separate product launch."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vkfft_b200 as vk

PEAK = 6575.4
CASES = [((256,), None), ((1024,), None), ((4096,), None), ((8192,), None), ((4096, 256), None), ((512, 512), None), ((256, 256, 64), None)]
for shape, _ in CASES:
    n = 1
    for s_ in shape:
        n *= s_
    batch = (1 << 28) // n
    buf = torch.zeros((batch,) + tuple(reversed(shape)), dtype=torch.complex64, device="cuda")
    torch.view_as_real(buf).uniform_(-1, 1)
    ker = torch.fft.fftn(torch.randn((1,) + tuple(reversed(shape)), dtype=torch.complex64, device="cuda"), dim=tuple(range(1, len(shape) + 1)))
    res = {}
    for mode in ("fused", "chain"):
        if mode == "chain":
            os.environ["B200FFT_NO_FUSED_CONV"] = "1"
        else:
            os.environ.pop("B200FFT_NO_FUSED_CONV", None)
        app = vk.VkFFTApplication()
        assert vk.initializeVkFFT(app, vk.VkFFTConfiguration(FFTdim=len(shape), size=list(shape), numberBatches=batch, device=0, performConvolution=1, normalize=1)) == 0
        lp = vk.VkFFTLaunchParams(buffer=buf, kernel=ker)
        for _ in range(3):
            assert vk.VkFFTAppend(app, -1, lp) == 0
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            vk.VkFFTAppend(app, -1, lp)
        b.record(); torch.cuda.synchronize()
        res[mode] = a.elapsed_time(b) / 10
        res[mode + "_launches"] = vk.planInfo(app)["num_passes_forward"]
        vk.deleteVkFFT(app)
    os.environ.pop("B200FFT_NO_FUSED_CONV", None)
    res.update(shape=list(shape), fused_frac_of_copy_peak=round(2 * buf.numel() * 8 / (res["fused"] * 1e-3) / 1e9 / PEAK, 3))
    print(json.dumps(res), flush=True)
