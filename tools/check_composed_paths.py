#!/usr/bin/env python3
"""GPU spot check of the paths added after the GPU budget ran out: composed odd R2C, composed large-prime R2R, zero padding."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vkfft_b200 as vk
from vkfft_b200 import fft as vkfft
out = {}
def rel(a, b): return float(((a - b).abs().double().norm() / b.abs().double().norm()).item())
for shape, dt in (((3, 131), torch.float32), ((2, 4391), torch.float32), ((2, 5, 263), torch.float64)):
    x = torch.randn(*shape, dtype=dt, device="cuda")
    nd = len(shape) - 1
    ref = torch.fft.rfftn(x.double(), dim=tuple(range(1, nd + 1)))
    h = vkfft.rfftn(x, ndim=nd, norm=0)
    out["r2c" + str(shape)] = rel(h.to(torch.complex128), ref)
    out["c2r" + str(shape)] = rel(vkfft.irfftn(h, ndim=nd, norm=1, n_last=shape[-1]), x)
import scipy.fft as sf
for kind, n in ((2, 131), (4, 131), (1, 132)):
    x = torch.randn(4, n, dtype=torch.float32, device="cuda")
    c = vkfft.dctn(x, ndim=1, dct_type=kind, norm=0)
    r = torch.from_numpy(sf.dct(x.cpu().numpy().astype(np.float64), type=kind, axis=-1)).cuda()
    out[f"dct{kind}_{n}"] = rel(c.double(), r)
x = torch.randn(2, 8, 12, 16, dtype=torch.complex64, device="cuda")
clean = x.clone(); clean[..., 8:] = 0; clean[:, :, 6:, :] = 0
app = vk.VkFFTApplication()
assert vk.initializeVkFFT(app, vk.VkFFTConfiguration(FFTdim=3, size=[16, 12, 8], numberBatches=2, device=0, performZeropadding=[1, 1, 0],
                                                     fft_zeropad_left=[8, 6, 0], fft_zeropad_right=[16, 12, 0])) == 0
buf = x.clone()
assert vk.VkFFTAppend(app, -1, vk.VkFFTLaunchParams(buffer=buf)) == 0
torch.cuda.synchronize()
out["zeropad"] = rel(buf.to(torch.complex128), torch.fft.fftn(clean.to(torch.complex128), dim=(1, 2, 3)))
print(json.dumps(out))
