"""CPU: every registered *operator flavour* of the specialised kernels, one by one (test_emu_kernels.py does the plain C2C ones):
Four-Step phase on store, fused R2C / C2R, fused DCT-II / DCT-III on contiguous lines and on strided axes, Bluestein launches,
fused convolution.  Each case runs through the planner so that the kernel under test is the one a user would get."""
import os
import re
import sys

import numpy as np
import pytest
import scipy.fft as sfft

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
import emu  # noqa: E402

OP_TW, OP_REAL_EVEN, OP_DCT23, OP_BLUESTEIN, OP_CONV = 1, 16, 32, 256, 512


def _sel(ops, kind, inv=0):
    ks = [k for k in emu.kernels() if k["ops"] == ops and k["kind"] == kind and k["inv"] == inv and k["variant"] == 0]
    return sorted(ks, key=lambda k: (k["prec"], k["n"]))


def _id(k):
    return f"p{k['prec']}-n{k['n']}"


def _rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def _tol(k):
    return 3e-6 if k["prec"] == 0 else 1e-12


@pytest.mark.parametrize("k", [k for k in _sel(OP_TW, emu.KIND_COLS) if k["n"] <= 2048], ids=_id)
def test_phase_on_store(k):
    """COLS + TWIDDLE_OUT alone: out[p][g] = FFT(column g)[p] * W_M^(g*p), M = n*G"""
    n, q = k["n"], k["q"]
    G = q + 3 if n * q <= 8192 else q
    dt = np.complex64 if k["prec"] == 0 else np.complex128
    rng = np.random.default_rng(n)
    x = (rng.uniform(-1, 1, (n, G)) + 1j * rng.uniform(-1, 1, (n, G))).astype(dt)
    y = np.zeros_like(x)
    emu.run_pass(emu.KIND_COLS, k["prec"], n, 0, OP_TW, x, y, G, in_gs=1, out_gs=1, in_es=G, out_es=G, twM=n * G)
    p, g = np.meshgrid(np.arange(n), np.arange(G), indexing="ij")
    ref = np.fft.fft(x.astype(np.complex128), axis=0) * np.exp(-2j * np.pi * (p * g) / (n * G))
    assert _rel(y, ref) < (4e-7 if k["prec"] == 0 else 1e-15)


@pytest.mark.parametrize("k", [k for k in _sel(OP_REAL_EVEN, emu.KIND_ROWS) if k["n"] <= 2048], ids=_id)
def test_fused_r2c_c2r(k):
    N = 2 * k["n"]
    rdt, cdt = (np.float32, np.complex64) if k["prec"] == 0 else (np.float64, np.complex128)
    x = np.random.default_rng(N).uniform(-1, 1, (3, N)).astype(rdt)
    buf = np.zeros((3, N + 2), rdt); buf[:, :N] = x
    d = emu.make_desc((N,), 3, k["prec"], perform_r2c=1)
    assert "fused" in emu.describe(d, -1)[1]
    assert emu.exec_plan(d, -1, buf)[0] == 0
    assert _rel(buf.view(cdt), np.fft.rfft(x.astype(np.float64), axis=-1)) < _tol(k)
    assert emu.exec_plan(d, 1, buf)[0] == 0
    assert _rel(buf[:, :N], x.astype(np.float64) * N) < _tol(k)


@pytest.mark.parametrize("k", [k for k in _sel(OP_DCT23, emu.KIND_ROWS) if k["n"] <= 2048], ids=_id)
def test_fused_dct_rows(k):
    n = k["n"]
    rdt = np.float32 if k["prec"] == 0 else np.float64
    x = np.random.default_rng(n).uniform(-1, 1, (3, n)).astype(rdt)      # odd line count: the last complex line holds one real line
    for kind in (2, 3):
        d = emu.make_desc((n,), 3, k["prec"], perform_dct=kind)
        assert "fused" in emu.describe(d, -1)[1]
        buf = x.copy()
        assert emu.exec_plan(d, -1, buf)[0] == 0
        assert _rel(buf, sfft.dct(x.astype(np.float64), type=kind, axis=-1)) < _tol(k)


@pytest.mark.parametrize("k", [k for k in _sel(OP_DCT23, emu.KIND_COLS) if k["n"] <= 1280], ids=_id)
def test_fused_dct_strided_axis(k):
    n = k["n"]
    rdt = np.float32 if k["prec"] == 0 else np.float64
    x = np.random.default_rng(n).uniform(-1, 1, (2, n, 12)).astype(rdt)
    for kind in (2, 3):
        d = emu.make_desc((12, n), 2, k["prec"], perform_dct=kind)
        assert "DCT_COLS" in emu.describe(d, -1)[1]
        buf = x.copy()
        assert emu.exec_plan(d, -1, buf)[0] == 0
        assert _rel(buf, sfft.dctn(x.astype(np.float64), type=kind, axes=(1, 2))) < _tol(k)


@pytest.mark.parametrize("k", _sel(OP_BLUESTEIN, emu.KIND_ROWS), ids=_id)
def test_bluestein_launches(k):
    """pick the largest prime N with 2N-1 <= M so that the plan pads exactly to this kernel's length"""
    M = k["n"]
    N = (M + 1) // 2
    def prime(v):
        return v > 1 and all(v % f for f in range(2, int(v ** 0.5) + 1))
    while N > 130 and not prime(N):
        N -= 1
    if N <= 130:
        pytest.skip("no Bluestein length this short (primes up to 127 run as Rader stages)")
    dt = np.complex64 if k["prec"] == 0 else np.complex128
    d = emu.make_desc((N,), 3, k["prec"])
    m = re.search(r"n=(\d+)", emu.describe(d, -1)[1])
    if int(m.group(1)) != M:
        pytest.skip(f"a shorter padded length ({m.group(1)}) serves N={N}")
    rng = np.random.default_rng(N)
    x = (rng.uniform(-1, 1, (3, N)) + 1j * rng.uniform(-1, 1, (3, N))).astype(dt)
    for inv in (-1, 1):
        buf = x.copy()
        assert emu.exec_plan(d, inv, buf)[0] == 0
        ref = np.fft.fft(x.astype(np.complex128), axis=-1) if inv == -1 else np.fft.ifft(x.astype(np.complex128), axis=-1) * N
        assert _rel(buf, ref) < _tol(k)


@pytest.mark.parametrize("k", _sel(OP_CONV, emu.KIND_ROWS) + _sel(OP_CONV, emu.KIND_COLS), ids=lambda k: f"kind{k['kind']}-" + _id(k))
def test_fused_convolution_kernels(k):
    n = k["n"]
    dt = np.complex64 if k["prec"] == 0 else np.complex128
    shape = (n,) if k["kind"] == emu.KIND_ROWS else (16, n)
    rng = np.random.default_rng(n)
    npshape = tuple(reversed(shape)); axes = tuple(range(-len(shape), 0))
    x = (rng.uniform(-1, 1, (2, 2) + npshape) + 1j * rng.uniform(-1, 1, (2, 2) + npshape)).astype(dt)
    kk = (rng.uniform(-1, 1, (2,) + npshape) + 1j * rng.uniform(-1, 1, (2,) + npshape)).astype(dt)
    K = np.fft.fftn(kk.astype(np.complex128), axes=axes).astype(dt)
    d = emu.make_desc(shape, 2, k["prec"], coordinate_features=2, perform_convolution=1, normalize=1)
    assert "fused convolution" in emu.describe(d, -1)[1]
    buf = x.copy()
    assert emu.exec_plan(d, -1, buf, kernel=K)[0] == 0
    ref = np.fft.ifftn(np.fft.fftn(x.astype(np.complex128), axes=axes) * K.astype(np.complex128)[None], axes=axes)
    assert _rel(buf, ref) < _tol(k)
