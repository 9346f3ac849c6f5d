"""GPU (-m gpu): same inputs through the engine and through the UNMODIFIED reference (CUDA backend, built into
oracle/_ref/libvkfft_ref.so by oracle/Makefile).  North-star tolerance: 1e-6 rel FP32 / 1e-12 rel FP64.
The reference's default FP32 path evaluates twiddles with __sincosf (its own error vs FFTW is up to ~1.4e-6,
README.md:76-80), so the comparison is norm-wise, and also run against the reference with useLUT=1."""
import numpy as np
import pytest

import vkfft_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    import torch
    assert torch.cuda.is_available()
    if not orc.ref_available():
        pytest.skip("oracle/_ref/libvkfft_ref.so not built (needs /root/reference at build time)")
    return orc.ref_lib()


def _both(torch, size_xyz, batch, inverse, double, use_lut):
    from gpu_util import run_c2c
    dt = np.complex128 if double else np.complex64
    x = orc.random_input((batch,) + tuple(reversed(size_xyz)), dt, seed=int(np.prod(size_xyz)) % 9973)
    mine = run_c2c(x, size_xyz, batch, inverse, double=double)
    t = torch.from_numpy(x.copy()).cuda()
    rc = orc.ref_run(orc.ref_desc(size_xyz, batch, double, use_lut=use_lut), inverse, t.data_ptr())
    assert rc == 0, rc
    theirs = t.cpu().numpy()
    return x, mine, theirs


@pytest.mark.parametrize("n", [8, 128, 1024, 4096, 8192, 1 << 15, 1 << 18, 1 << 20, 1 << 23])
@pytest.mark.parametrize("inverse", [-1, 1])
def test_c2c_f32_matches_reference(ref, n, inverse):
    import torch
    batch = max(1, (1 << 23) // n)
    x, mine, theirs = _both(torch, (n,), batch, inverse, False, use_lut=1)
    assert orc.error_metrics(mine, theirs)["l2_rel"] < 1e-6
    x, mine, theirs = _both(torch, (n,), batch, inverse, False, use_lut=0)
    # reference default (on-chip sincos) carries its own ~1e-6 error for large N; both must sit within 1e-6 of
    # the exact result's neighbourhood: |mine - theirs| <= |mine - exact| + |theirs - exact|
    exact = orc.c2c(x, 1, inverse == 1)
    e_m = orc.error_metrics(mine, exact)["l2_rel"]
    e_t = orc.error_metrics(theirs, exact)["l2_rel"]
    assert e_m < 1e-6 and e_m <= e_t * 1.05 + 1e-8
    assert orc.error_metrics(mine, theirs)["l2_rel"] < e_m + e_t + 1e-9


@pytest.mark.parametrize("size_xyz", [(4096,), (1 << 16,), (256, 256, 256)])
def test_c2c_f64_matches_reference(ref, size_xyz):
    import torch
    x, mine, theirs = _both(torch, size_xyz, 1, -1, True, use_lut=0)
    assert orc.error_metrics(mine, theirs)["l2_rel"] < 1e-12


def _ref_inplace(torch, arr, size_xyz, batch, inverse, double=False, **kw):
    t = torch.from_numpy(np.ascontiguousarray(arr)).cuda()
    rc = orc.ref_run(orc.ref_desc(size_xyz, batch, double, use_lut=1, **kw), inverse, t.data_ptr())
    assert rc == 0, rc
    return t.cpu().numpy()


def _mine_inplace(torch, arr, size_xyz, batch, inverse, double=False, **kw):
    import vkfft_b200 as vk
    t = torch.from_numpy(np.ascontiguousarray(arr)).cuda()
    app = vk.VkFFTApplication()
    cfg = vk.VkFFTConfiguration(FFTdim=len(size_xyz), size=list(size_xyz), numberBatches=batch, device=0,
                                doublePrecision=int(double), **kw)
    rc = vk.initializeVkFFT(app, cfg)
    assert rc == 0, vk.getVkFFTErrorString(rc)
    try:
        assert vk.VkFFTAppend(app, inverse, vk.VkFFTLaunchParams(buffer=t)) == 0
        torch.cuda.synchronize()
        return t.cpu().numpy()
    finally:
        vk.deleteVkFFT(app)


@pytest.mark.parametrize("size_xyz,batch", [((1000,), 8), ((2187,), 3), ((30030,), 2), ((17,), 64), ((509,), 8), ((105, 30), 2)])
@pytest.mark.parametrize("inverse", [-1, 1])
def test_non_pow2_matches_reference(ref, size_xyz, batch, inverse):
    import torch
    x = orc.random_input((batch,) + tuple(reversed(size_xyz)), np.complex64, seed=sum(size_xyz))
    mine = _mine_inplace(torch, x, size_xyz, batch, inverse)
    theirs = _ref_inplace(torch, x, size_xyz, batch, inverse)
    assert orc.error_metrics(mine, theirs)["l2_rel"] < 1e-6


@pytest.mark.parametrize("size_xyz,batch", [((64,), 8), ((4096,), 4), ((4096, 4096), 1), ((30, 4), 3)])
def test_r2c_c2r_matches_reference(ref, size_xyz, batch):
    import torch
    nx, H = size_xyz[0], size_xyz[0] // 2 + 1
    x = orc.random_input((batch,) + tuple(reversed(size_xyz)), np.float32, seed=sum(size_xyz))
    buf = np.zeros(x.shape[:-1] + (2 * H,), np.float32)
    buf[..., :nx] = x
    mine = _mine_inplace(torch, buf, size_xyz, batch, -1, performR2C=1)
    theirs = _ref_inplace(torch, buf, size_xyz, batch, -1, perform_r2c=1)
    assert orc.error_metrics(mine.view(np.complex64), theirs.view(np.complex64))["l2_rel"] < 1e-6
    mine2 = _mine_inplace(torch, theirs, size_xyz, batch, 1, performR2C=1)
    theirs2 = _ref_inplace(torch, theirs, size_xyz, batch, 1, perform_r2c=1)
    assert orc.error_metrics(mine2[..., :nx], theirs2[..., :nx])["l2_rel"] < 1e-6


@pytest.mark.parametrize("kind", [1, 2, 3, 4])
@pytest.mark.parametrize("size_xyz,batch", [((64,), 6), ((100,), 4), ((32, 16), 3), ((2048, 256), 1)])
@pytest.mark.parametrize("inverse", [-1, 1])
def test_dct_matches_reference(ref, kind, size_xyz, batch, inverse):
    import torch
    def smooth(n):
        for p in [2, 3, 5, 7, 11, 13] + [q for q in range(17, 128, 2) if all(q % r for r in range(3, 12, 2))]:
            while n % p == 0:
                n //= p
        return n == 1
    x = orc.random_input((batch,) + tuple(reversed(size_xyz)), np.float32, seed=kind + sum(size_xyz))
    mine = _mine_inplace(torch, x, size_xyz, batch, inverse, performDCT=kind)
    theirs = _ref_inplace(torch, x, size_xyz, batch, inverse, perform_dct=kind)
    # north-star 1e-6 between the two engines; where the transform's conditioning puts the reference itself further than
    # that from the exact result, this engine must be at least as close to it as the reference is
    d = orc.error_metrics(mine, theirs)["l2_rel"]
    if d >= 1e-6:
        exact = orc.dct(x, kind, len(size_xyz), inverse=(inverse == 1))
        e_m = orc.error_metrics(mine, exact)["l2_rel"]
        e_t = orc.error_metrics(theirs, exact)["l2_rel"]
        assert e_m <= 1.05 * e_t + 1e-8 and d < e_m + e_t + 1e-9, (d, e_m, e_t)
