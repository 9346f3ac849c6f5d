"""CPU: performConvolution plans (forward transform -> product with the kernel spectrum -> inverse transform behind one
VkFFTAppend(app, -1)) on the kernel-body emulation, against numpy.  Mirrors the reference's samples 50-52
(benchmark_scripts/vkFFT_scripts/src/sample_50/51/52_convolution_*.cpp) with random kernels instead of identities."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
import emu  # noqa: E402

T32 = 2e-6


def _rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def _rand(shape, seed, cplx=True):
    rng = np.random.default_rng(seed)
    x = rng.uniform(-1, 1, shape)
    return (x + 1j * rng.uniform(-1, 1, shape)).astype(np.complex64) if cplx else x.astype(np.float32)


@pytest.mark.parametrize("shape", [(256,), (48, 20), (16, 8, 12), (1 << 13,)])
def test_per_feature_convolution_c2c(shape):
    """coordinateFeatures = 2, 1x1 product per feature, two batches sharing the kernel"""
    C, B = 2, 2
    env = {"B200FFT_MAX_SINGLE_PASS": "1024"} if shape == (1 << 13,) else {}
    os.environ.update(env)
    try:
        nd = len(shape)
        np_shape = tuple(reversed(shape))
        axes = tuple(range(-nd, 0))
        k = _rand((C,) + np_shape, 1)
        x = _rand((B, C) + np_shape, 2)
        # kernel application: a plain forward plan with the same layout (kernelConvolution = 1)
        K = k.copy()
        rc, _ = emu.exec_plan(emu.make_desc(shape, 1, 0, coordinate_features=C, kernel_convolution=1), -1, K)
        assert rc == 0 and _rel(K, np.fft.fftn(k.astype(np.complex128), axes=axes)) < T32
        buf = x.copy()
        d = emu.make_desc(shape, B, 0, coordinate_features=C, perform_convolution=1, normalize=1)
        rc, _ = emu.exec_plan(d, -1, buf, kernel=K)
        assert rc == 0
        ref = np.fft.ifftn(np.fft.fftn(x.astype(np.complex128), axes=axes) * np.fft.fftn(k.astype(np.complex128), axes=axes)[None], axes=axes)
        assert _rel(buf, ref) < T32
        assert emu.exec_plan(d, 1, buf, kernel=K)[0] == 1006      # a convolution application only runs "forward"
    finally:
        for key in env:
            os.environ.pop(key, None)


@pytest.mark.parametrize("M,sym", [(2, 0), (2, 1), (3, 0), (3, 1)])
def test_matrix_vector_convolution(M, sym):
    """sample 50: out_r = sum_c K_rc (*) in_c, kernel stored row-major (or as its upper triangle when symmetric)"""
    n = 512
    x = _rand((M, n), 3)
    kfull = _rand((M, M, n), 4)
    if sym:
        for r in range(M):
            for c in range(r):
                kfull[r, c] = kfull[c, r]
        planes = [kfull[r, c] for r in range(M) for c in range(r, M)]
    else:
        planes = [kfull[r, c] for r in range(M) for c in range(M)]
    K = np.fft.fft(np.stack(planes).astype(np.complex128), axis=-1).astype(np.complex64)
    buf = x.copy()
    d = emu.make_desc((n,), 1, 0, coordinate_features=M, matrix_convolution=M, symmetric_kernel=sym, perform_convolution=1, normalize=1)
    rc, _ = emu.exec_plan(d, -1, buf, kernel=K)
    assert rc == 0
    X = np.fft.fft(x.astype(np.complex128), axis=-1)
    KK = np.fft.fft(kfull.astype(np.complex128), axis=-1)
    ref = np.fft.ifft(np.einsum("rcf,cf->rf", KK, X), axis=-1)
    assert _rel(buf, ref) < T32


def test_one_input_many_kernels_r2c_out_of_place():
    """sample 52: real input in its own unpadded buffer, numberKernels outputs in the padded R2C layout"""
    nx, ny, C, NK = 32, 24, 2, 3
    x = _rand((C, ny, nx), 5, cplx=False)
    k = _rand((NK, C, ny, nx), 6, cplx=False)
    K = np.fft.rfft2(k.astype(np.float64)).astype(np.complex64)                      # [NK][C][ny][nx/2+1]
    buf = np.zeros((NK, C, ny, nx + 2), np.float32)
    d = emu.make_desc((nx, ny), 1, 0, coordinate_features=C, perform_r2c=1, is_input_formatted=1, number_kernels=NK,
                      perform_convolution=1, normalize=1)
    rc, _ = emu.exec_plan(d, -1, buf, inp=x.copy(), kernel=K)
    assert rc == 0
    ref = np.fft.irfft2(np.fft.rfft2(x.astype(np.float64))[None] * np.fft.rfft2(k.astype(np.float64)), s=(ny, nx))
    assert _rel(buf[..., :nx], ref) < T32


def test_cross_correlation_and_phase_correlation():
    n = 1024
    x, k = _rand((n,), 7), _rand((n,), 8)
    K = np.fft.fft(k.astype(np.complex128)).astype(np.complex64)
    X = np.fft.fft(x.astype(np.complex128))
    buf = x.copy()
    rc, _ = emu.exec_plan(emu.make_desc((n,), 1, 0, perform_convolution=1, conjugate_convolution=2, normalize=1), -1, buf, kernel=K)
    assert rc == 0 and _rel(buf, np.fft.ifft(X * np.conj(K))) < T32                  # cross-correlation
    buf = x.copy()
    rc, _ = emu.exec_plan(emu.make_desc((n,), 1, 0, perform_convolution=1, conjugate_convolution=1, normalize=1), -1, buf, kernel=K)
    assert rc == 0 and _rel(buf, np.fft.ifft(np.conj(X) * K)) < T32
    buf = x.copy()
    d = emu.make_desc((n,), 1, 0, perform_convolution=1, conjugate_convolution=2, cross_power_spectrum_normalization=1, normalize=1)
    rc, _ = emu.exec_plan(d, -1, buf, kernel=K)
    P = X * np.conj(K)
    assert rc == 0 and _rel(buf, np.fft.ifft(P / np.abs(P))) < 1e-5                  # phase correlation


def test_convolution_rejects_what_the_reference_rejects():
    buf = np.zeros(256, np.complex64)
    assert emu.exec_plan(emu.make_desc((16, 16), 1, 0, perform_convolution=1, omit_dimension=[0, 1]), -1, buf, kernel=buf)[0] == 3005
    assert emu.exec_plan(emu.make_desc((256,), 1, 0, perform_convolution=1, coordinate_features=2, matrix_convolution=3), -1, buf, kernel=buf)[0] == 3002


@pytest.mark.parametrize("n,prec", [(64, 0), (128, 0), (256, 0), (512, 0), (1024, 0), (2048, 0), (4096, 0), (8192, 0),
                                    (64, 1), (256, 1), (512, 1), (1024, 1), (4096, 1)])
def test_fused_convolution_single_launch(n, prec):
    """1-D C2C with a palindromic schedule: FFT, kernel product and iFFT in ONE launch; same numbers as the three-launch chain"""
    C, B = 3, 2
    cdt = np.complex64 if prec == 0 else np.complex128
    k = _rand((C, n), 11).astype(cdt)
    x = _rand((B, C, n), 12).astype(cdt)
    K = np.fft.fft(k.astype(np.complex128), axis=-1).astype(cdt)
    d = emu.make_desc((n,), B, prec, coordinate_features=C, perform_convolution=1, normalize=1)
    assert "fused convolution" in emu.describe(d, -1)[1]
    buf = x.copy()
    rc, npass = emu.exec_plan(d, -1, buf, kernel=K)
    assert rc == 0 and npass == 1
    ref = np.fft.ifft(np.fft.fft(x.astype(np.complex128), axis=-1) * K.astype(np.complex128)[None], axis=-1)
    assert _rel(buf, ref) < (T32 if prec == 0 else 1e-13)
    os.environ["B200FFT_NO_FUSED_CONV"] = "1"
    try:
        chain = x.copy()
        rc, npass3 = emu.exec_plan(d, -1, chain, kernel=K)
    finally:
        os.environ.pop("B200FFT_NO_FUSED_CONV")
    assert rc == 0 and npass3 == 3 and _rel(buf, chain) < (T32 if prec == 0 else 1e-13)


def test_fused_convolution_options_and_ragged_lines():
    n, C, B = 1024, 1, 5          # 5 lines, 4 per CTA: the last CTA is ragged
    x, k = _rand((B, C, n), 13), _rand((C, n), 14)
    K = np.fft.fft(k.astype(np.complex128), axis=-1).astype(np.complex64)
    X = np.fft.fft(x.astype(np.complex128), axis=-1)
    for conj, xps, ref in ((2, 0, X * np.conj(K)[None]), (1, 0, np.conj(X) * K[None]), (2, 1, None)):
        buf = x.copy()
        d = emu.make_desc((n,), B, 0, coordinate_features=C, perform_convolution=1, conjugate_convolution=conj,
                          cross_power_spectrum_normalization=xps)          # no normalize: unnormalised inverse
        rc, npass = emu.exec_plan(d, -1, buf, kernel=K)
        assert rc == 0 and npass == 1
        if ref is None:
            P = X * np.conj(K.astype(np.complex128))[None]
            ref = P / np.abs(P)
        assert _rel(buf, np.fft.ifft(ref, axis=-1) * n) < 1e-5


@pytest.mark.parametrize("shape,r2c,prec", [((48, 64), 0, 0), ((20, 6, 256), 0, 0), ((64, 128), 1, 0), ((32, 8, 16), 1, 0), ((24, 512), 0, 1),
                                            ((16, 2048), 0, 0), ((40, 8), 1, 1)])
def test_fused_last_axis_of_nd_convolution(shape, r2c, prec):
    """N-D: the last axis runs forward + product + inverse in one launch, the other axes keep their passes: 2*nd-1 launches"""
    C, B = 2, 2
    nd = len(shape)
    np_shape = tuple(reversed(shape))
    axes = tuple(range(-nd, 0))
    tol = T32 if prec == 0 else 1e-13
    rdt, cdt = (np.float32, np.complex64) if prec == 0 else (np.float64, np.complex128)
    if r2c:
        k = _rand((C,) + np_shape, 21, cplx=False).astype(rdt)
        x = _rand((B, C) + np_shape, 22, cplx=False).astype(rdt)
        K = np.fft.rfftn(k.astype(np.float64), axes=axes).astype(cdt)
        buf = np.zeros((B, C) + np_shape[:-1] + (shape[0] + 2,), rdt)
        buf[..., :shape[0]] = x
        ref = np.fft.irfftn(np.fft.rfftn(x.astype(np.float64), axes=axes) * K.astype(np.complex128)[None], s=np_shape, axes=axes)
    else:
        k = _rand((C,) + np_shape, 21).astype(cdt)
        x = _rand((B, C) + np_shape, 22).astype(cdt)
        K = np.fft.fftn(k.astype(np.complex128), axes=axes).astype(cdt)
        buf = x.copy()
        ref = np.fft.ifftn(np.fft.fftn(x.astype(np.complex128), axes=axes) * K.astype(np.complex128)[None], axes=axes)
    d = emu.make_desc(shape, B, prec, coordinate_features=C, perform_convolution=1, perform_r2c=r2c, normalize=1)
    listing = emu.describe(d, -1)[1]
    assert "fused convolution" in listing, listing
    rc, npass = emu.exec_plan(d, -1, buf, kernel=K)
    assert rc == 0 and npass == 2 * nd - 1, (npass, listing)
    got = buf[..., :shape[0]] if r2c else buf
    assert _rel(got, ref) < tol


def test_zero_padded_convolution_like_sample_51():
    """sample_51_convolution_VkFFT_single_3d_matrix_zeropadding_r2c.cpp: an open system -- only the first half of every axis
    carries data, the rest is padding that may hold anything; circular convolution of the padded system = linear convolution"""
    nx, ny, C = 32, 16, 2
    rng = np.random.default_rng(51)
    x = rng.uniform(-1, 1, (C, ny, nx)).astype(np.float32)            # garbage in the padded half on purpose
    clean = x.copy(); clean[..., nx // 2:] = 0; clean[:, ny // 2:, :] = 0
    k = rng.uniform(-1, 1, (C, ny, nx)).astype(np.float32)
    K = np.fft.rfft2(k.astype(np.float64)).astype(np.complex64)
    buf = np.full((C, ny, nx + 2), 3.0, np.float32); buf[..., :nx] = x
    d = emu.make_desc((nx, ny), 1, 0, coordinate_features=C, perform_r2c=1, perform_convolution=1, normalize=1,
                      perform_zeropadding=[1, 1], zeropad_left=[nx // 2, ny // 2], zeropad_right=[nx, ny])
    rc, _ = emu.exec_plan(d, -1, buf, kernel=K)
    assert rc == 0
    ref = np.fft.irfft2(np.fft.rfft2(clean.astype(np.float64)) * K.astype(np.complex128), s=(ny, nx))
    assert _rel(buf[..., :nx], ref) < T32


def test_sample_51_configuration_3d_r2c_matrix_kernel_with_zero_padding():
    """the exact option set of sample_51_convolution_VkFFT_single_3d_matrix_zeropadding_r2c.cpp:77-206 on a smaller grid:
    kernel application (kernelConvolution, 9 features, R2C, zero padding), then 3x3 non-symmetric matrix convolution of a
    3-vector field with zero padding on all axes"""
    n = 16
    rng = np.random.default_rng(510)
    zp = dict(perform_zeropadding=[1, 1, 1], zeropad_left=[n // 2] * 3, zeropad_right=[n] * 3)
    # kernel: 9 real fields, garbage in the padded half, transformed by the kernel application
    kr = rng.uniform(-1, 1, (9, n, n, n)).astype(np.float32)
    kclean = kr.copy(); kclean[..., n // 2:] = 0; kclean[:, :, n // 2:, :] = 0; kclean[:, n // 2:, :, :] = 0
    kbuf = np.zeros((9, n, n, n + 2), np.float32); kbuf[..., :n] = kr
    dk = emu.make_desc((n, n, n), 1, 0, coordinate_features=9, perform_r2c=1, kernel_convolution=1, normalize=1, **zp)
    assert emu.exec_plan(dk, -1, kbuf)[0] == 0
    K = kbuf.view(np.complex64).copy()                                       # [9][n][n][n/2+1]
    assert _rel(K, np.fft.rfftn(kclean.astype(np.float64), axes=(1, 2, 3))) < T32
    x = rng.uniform(-1, 1, (3, n, n, n)).astype(np.float32)
    xclean = x.copy(); xclean[..., n // 2:] = 0; xclean[:, :, n // 2:, :] = 0; xclean[:, n // 2:, :, :] = 0
    buf = np.zeros((3, n, n, n + 2), np.float32); buf[..., :n] = x
    dc = emu.make_desc((n, n, n), 1, 0, coordinate_features=3, matrix_convolution=3, perform_r2c=1, perform_convolution=1, normalize=1, **zp)
    rc, _ = emu.exec_plan(dc, -1, buf, kernel=K)
    assert rc == 0
    X = np.fft.rfftn(xclean.astype(np.float64), axes=(1, 2, 3))
    KK = K.astype(np.complex128).reshape(3, 3, n, n, n // 2 + 1)
    ref = np.fft.irfftn(np.einsum("rczyx,czyx->rzyx", KK, X), s=(n, n, n), axes=(1, 2, 3))
    assert _rel(buf[..., :n], ref) < T32
