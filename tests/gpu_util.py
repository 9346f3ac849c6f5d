"""Helpers for the -m gpu tests: run a plan through the C ABI (ctypes) on torch-owned device memory."""
import numpy as np

import vkfft_b200 as vk


def torch_mod():
    import torch
    return torch


def run_c2c(x_np, size_xyz, batches=1, inverse=-1, double=False, **cfgkw):
    """x_np: numpy complex array [batch, ..., y, x] (contiguous).  Returns the transformed numpy array."""
    torch = torch_mod()
    t = torch.from_numpy(np.ascontiguousarray(x_np)).cuda()
    cfg = vk.VkFFTConfiguration(FFTdim=len(size_xyz), size=list(size_xyz), numberBatches=batches, device=0,
                                doublePrecision=int(double), **cfgkw)
    app = vk.VkFFTApplication()
    rc = vk.initializeVkFFT(app, cfg)
    assert rc == vk.VKFFT_SUCCESS, vk.getVkFFTErrorString(rc)
    try:
        rc = vk.VkFFTAppend(app, inverse, vk.VkFFTLaunchParams(buffer=t))
        assert rc == vk.VKFFT_SUCCESS, vk.getVkFFTErrorString(rc)
        torch.cuda.synchronize()
        out = t.cpu().numpy()
    finally:
        vk.deleteVkFFT(app)
    return out
