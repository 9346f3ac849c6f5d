"""GPU: performConvolution through the C ABI (the reference's samples 50-52 with random kernels), against numpy."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    return torch


def _rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def _run(torch, cfg_kw, buf_np, kernel_np, inp_np=None):
    import vkfft_b200 as vk
    buf = torch.from_numpy(buf_np).cuda()
    ker = torch.from_numpy(kernel_np).cuda()
    inp = torch.from_numpy(inp_np).cuda() if inp_np is not None else None
    app = vk.VkFFTApplication()
    rc = vk.initializeVkFFT(app, vk.VkFFTConfiguration(device=0, performConvolution=1, normalize=1, **cfg_kw))
    assert rc == 0, vk.getVkFFTErrorString(rc)
    assert vk.VkFFTAppend(app, -1, vk.VkFFTLaunchParams(buffer=buf)) == vk.VKFFT_ERROR_EMPTY_kernel
    assert vk.VkFFTAppend(app, -1, vk.VkFFTLaunchParams(buffer=buf, kernel=ker, inputBuffer=inp)) == 0
    assert vk.VkFFTAppend(app, 1, vk.VkFFTLaunchParams(buffer=buf, kernel=ker, inputBuffer=inp)) == vk.VKFFT_ERROR_ONLY_FORWARD_FFT_INITIALIZED
    torch.cuda.synchronize()
    vk.deleteVkFFT(app)
    return buf.cpu().numpy()


@pytest.mark.parametrize("shape", [(4096,), (256, 64), (1 << 18,), (1000,), (48, 32, 256), (8192,), (20, 2048)])
def test_convolution_c2c(torch_cuda, shape):
    import vkfft_b200 as vk
    torch = torch_cuda
    rng = np.random.default_rng(3)
    C, B = 2, 3
    np_shape = tuple(reversed(shape))
    axes = tuple(range(-len(shape), 0))
    cplx = lambda s: (rng.uniform(-1, 1, s) + 1j * rng.uniform(-1, 1, s)).astype(np.complex64)
    k, x = cplx((C,) + np_shape), cplx((B, C) + np_shape)
    # kernel application (kernelConvolution = 1): a forward transform with the same layout
    K = torch.from_numpy(k).cuda()
    app = vk.VkFFTApplication()
    assert vk.initializeVkFFT(app, vk.VkFFTConfiguration(FFTdim=len(shape), size=list(shape), coordinateFeatures=C, device=0,
                                                         kernelConvolution=1)) == 0
    assert vk.VkFFTAppend(app, -1, vk.VkFFTLaunchParams(buffer=K)) == 0
    torch.cuda.synchronize()
    vk.deleteVkFFT(app)
    out = _run(torch, dict(FFTdim=len(shape), size=list(shape), coordinateFeatures=C, numberBatches=B), x, K.cpu().numpy())
    ref = np.fft.ifftn(np.fft.fftn(x.astype(np.complex128), axes=axes) * np.fft.fftn(k.astype(np.complex128), axes=axes)[None], axes=axes)
    assert _rel(out, ref) < 2e-6


def test_matrix_convolution_3x3_symmetric(torch_cuda):
    rng = np.random.default_rng(4)
    n, M = 8192, 3
    cplx = lambda s: (rng.uniform(-1, 1, s) + 1j * rng.uniform(-1, 1, s)).astype(np.complex64)
    x, kfull = cplx((M, n)), cplx((M, M, n))
    for r in range(M):
        for c in range(r):
            kfull[r, c] = kfull[c, r]
    K = np.fft.fft(np.stack([kfull[r, c] for r in range(M) for c in range(r, M)]).astype(np.complex128), axis=-1).astype(np.complex64)
    out = _run(torch_cuda, dict(FFTdim=1, size=[n], coordinateFeatures=M, matrixConvolution=M, symmetricKernel=1), x, K)
    ref = np.fft.ifft(np.einsum("rcf,cf->rf", np.fft.fft(kfull.astype(np.complex128), axis=-1), np.fft.fft(x.astype(np.complex128), axis=-1)), axis=-1)
    assert _rel(out, ref) < 2e-6


@pytest.mark.parametrize("shape", [(64, 128), (256, 16, 64)])
def test_convolution_r2c_fused_last_axis(torch_cuda, shape):
    """padded in-place R2C layout; the last axis runs forward + product + inverse in one launch"""
    import vkfft_b200 as vk
    rng = np.random.default_rng(6)
    C, B = 2, 3
    np_shape = tuple(reversed(shape))
    axes = tuple(range(-len(shape), 0))
    x = rng.uniform(-1, 1, (B, C) + np_shape).astype(np.float32)
    k = rng.uniform(-1, 1, (C,) + np_shape).astype(np.float32)
    K = np.fft.rfftn(k.astype(np.float64), axes=axes).astype(np.complex64)
    buf = np.zeros((B, C) + np_shape[:-1] + (shape[0] + 2,), np.float32)
    buf[..., :shape[0]] = x
    cfg = dict(FFTdim=len(shape), size=list(shape), coordinateFeatures=C, numberBatches=B, performR2C=1)
    app = vk.VkFFTApplication()
    assert vk.initializeVkFFT(app, vk.VkFFTConfiguration(device=0, performConvolution=1, normalize=1, **cfg)) == 0
    info = vk.planInfo(app)
    vk.deleteVkFFT(app)
    assert "fused convolution" in info["forward"] and info["num_passes_forward"] == 2 * len(shape) - 1
    out = _run(torch_cuda, cfg, buf, K)
    ref = np.fft.irfftn(np.fft.rfftn(x.astype(np.float64), axes=axes) * K.astype(np.complex128)[None], s=np_shape, axes=axes)
    assert _rel(out[..., :shape[0]], ref) < 2e-6


def test_one_input_many_kernels_r2c(torch_cuda):
    rng = np.random.default_rng(5)
    nx, ny, C, NK = 64, 48, 2, 3
    x = rng.uniform(-1, 1, (C, ny, nx)).astype(np.float32)
    k = rng.uniform(-1, 1, (NK, C, ny, nx)).astype(np.float32)
    K = np.fft.rfft2(k.astype(np.float64)).astype(np.complex64)
    buf = np.zeros((NK, C, ny, nx + 2), np.float32)
    out = _run(torch_cuda, dict(FFTdim=2, size=[nx, ny], coordinateFeatures=C, performR2C=1, isInputFormatted=1, numberKernels=NK), buf, K, inp_np=x)
    ref = np.fft.irfft2(np.fft.rfft2(x.astype(np.float64))[None] * np.fft.rfft2(k.astype(np.float64)), s=(ny, nx))
    assert _rel(out[..., :nx], ref) < 2e-6
