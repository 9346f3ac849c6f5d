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
