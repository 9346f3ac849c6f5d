"""GPU: the pyvkfft-style convenience layer (vkfft_b200/fft.py) against torch.fft / the oracle."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    return torch


def _rel(a, b):
    return float(((a - b).abs().double().norm() / b.abs().double().norm()).item())


@pytest.mark.parametrize("shape,ndim", [((6, 4096), 1), ((3, 64, 128), 2), ((2, 16, 24, 40), 3), ((1 << 17,), 1)])
def test_fftn_ifftn_match_torch(torch_cuda, shape, ndim):
    torch = torch_cuda
    from vkfft_b200 import fft as vkfft
    x = torch.randn(*shape, dtype=torch.complex64, device="cuda")
    dims = tuple(range(len(shape) - ndim, len(shape)))
    ref = torch.fft.fftn(x.to(torch.complex128), dim=dims)
    y = vkfft.fftn(x, ndim=ndim)
    assert y.data_ptr() != x.data_ptr() and _rel(y.to(torch.complex128), ref) < 1e-6
    back = vkfft.ifftn(y, ndim=ndim)                       # out of place, norm=1 -> numpy convention
    assert _rel(back, x) < 1e-6
    z = x.clone()
    assert vkfft.fftn(z, z, ndim=ndim).data_ptr() == z.data_ptr()      # in place
    assert _rel(z.to(torch.complex128), ref) < 1e-6
    o = vkfft.fftn(x, ndim=ndim, norm="ortho")
    assert _rel(o.to(torch.complex128), torch.fft.fftn(x.to(torch.complex128), dim=dims, norm="ortho")) < 1e-6


@pytest.mark.parametrize("shape,ndim,dtype", [((5, 1024), 1, "float32"), ((2, 32, 64), 2, "float32"), ((3, 48, 20), 2, "float64")])
def test_rfftn_irfftn(torch_cuda, shape, ndim, dtype):
    torch = torch_cuda
    from vkfft_b200 import fft as vkfft
    rdt = getattr(torch, dtype)
    x = torch.randn(*shape, dtype=rdt, device="cuda")
    dims = tuple(range(len(shape) - ndim, len(shape)))
    ref = torch.fft.rfftn(x.double(), dim=dims)
    h = vkfft.rfftn(x, ndim=ndim)
    tol = 1e-6 if dtype == "float32" else 1e-12
    assert h.shape == ref.shape and _rel(h.to(torch.complex128), ref) < tol
    r = vkfft.irfftn(h, ndim=ndim, n_last=shape[-1])
    assert _rel(r, x) < tol


@pytest.mark.parametrize("kind", [1, 2, 3, 4])
def test_dctn_dstn_round_trip_and_definition(torch_cuda, kind):
    torch = torch_cuda
    import scipy.fft as sfft
    from vkfft_b200 import fft as vkfft
    x = torch.randn(3, 20, 32, dtype=torch.float32, device="cuda")
    c = vkfft.dctn(x, ndim=2, dct_type=kind, norm=0)
    ref = sfft.dctn(x.cpu().numpy().astype(np.float64), type=kind, axes=(1, 2))
    assert np.linalg.norm(c.cpu().numpy() - ref) / np.linalg.norm(ref) < 2e-6
    assert _rel(vkfft.idctn(c, ndim=2, dct_type=kind, norm=1), x) < 2e-6
    s = vkfft.dstn(x, ndim=2, dst_type=kind, norm=0)
    refs = sfft.dstn(x.cpu().numpy().astype(np.float64), type=kind, axes=(1, 2))
    assert np.linalg.norm(s.cpu().numpy() - refs) / np.linalg.norm(refs) < 2e-6
    assert _rel(vkfft.idstn(s, ndim=2, dst_type=kind, norm=1), x) < 2e-6
    vkfft.clear_cache()
