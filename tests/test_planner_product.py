"""Planning in the PRODUCT library without a device (b200fft_debug_plan_text): unlike the emulation's planner it sees the
kernels that jit.cpp instantiates at plan time.  Sweeps: every length plans, smooth lengths outside the ahead-of-time lists
get a template (never the runtime-scheduled kernel) wherever jit.cpp says it serves them, half-precision plans only contain
half kernels, and the switch restores the old routing."""
import ctypes
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "emu"))


def _text(L, shape, batch=4, prec=0, inverse=-1, **kw):
    import emu                                   # only for the ctypes mirror of b200fft_desc
    d = emu.make_desc(shape, batch, prec, **kw)
    buf = ctypes.create_string_buffer(1 << 15)
    rc = L.b200fft_debug_plan_text(ctypes.byref(d), int(inverse), buf, len(buf))
    return rc, buf.value.decode()


@pytest.fixture(scope="module")
def lib():
    from vkfft_b200 import _lib
    L = _lib.load()
    if not L.b2_jit_available():
        pytest.skip("libnvrtc not loadable here: no plan-time kernels to plan with")
    return L


def _smooth(n, primes=(2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31)):
    for p in primes:
        while n % p == 0:
            n //= p
    return n == 1


def test_every_length_up_to_8300_plans_in_both_precisions(lib):
    """FP32 contiguous lines: up to 4096 points nothing but the tiny lengths below 18 runs on the runtime-scheduled kernel any more
    (ahead-of-time kernel, plan-time template, or Bluestein on specialised launches); 31-smooth lengths up to 8192 neither"""
    jit = 0
    for n in range(2, 8301):
        for prec in (0, 1):
            rc, txt = _text(lib, (n,), 3, prec)
            assert rc == 0 and txt.startswith("pass 0"), (n, prec, rc)
            if prec == 0:
                jit += "JIT_" in txt
                if "generic" in txt:
                    assert n < 18 or n > 8192 or (n > 4096 and not _smooth(n)), f"N={n} fell back to the runtime-scheduled kernel:\n{txt}"
                if 2048 < n <= 4096 and not _smooth(n):
                    assert txt.count("bluestein") == 2, txt          # two specialised launches on a padded length of 8192
    assert jit > 1000, jit


def test_real_and_cosine_transforms_pick_up_the_plan_time_kernels(lib):
    for n in (1100, 2002, 3080, 6006):
        rc, txt = _text(lib, (n,), 2, 0, perform_r2c=1)
        assert rc == 0 and "JIT_ROWS" in txt and "fused" in txt, txt
    for n in (1100, 1430, 770):
        for kind in (2, 3):
            rc, txt = _text(lib, (n, 154), 2, 0, perform_dct=kind)
            assert rc == 0 and txt.count("JIT_") == 2 and "generic" not in txt, txt
    rc, txt = _text(lib, (1100, 1430, 66), 1, 1)       # FP64: the 1430-point strided axis exceeds the 1024-point tile -> Four-Step along the stride
    assert rc == 0 and txt.count("JIT_") == 4 and "generic" not in txt, txt
    rc, txt = _text(lib, (1100, 1430, 66), 1, 0)
    assert rc == 0 and txt.count("JIT_") == 3 and "generic" not in txt, txt


def test_long_lengths_split_into_plan_time_factors(lib):
    for n in (1100 * 1430, 2002 * 1001, 154 * 154 * 154):
        rc, txt = _text(lib, (n,), 1, 0)
        assert rc == 0 and "generic" not in txt and "four-step" in txt, txt


@pytest.mark.parametrize("prec", [2])
def test_half_storage_plans_contain_only_half_kernels(lib, prec):
    for n in list(range(2, 300)) + [1000, 1100, 4096, 5000, 8192, 1 << 14, 1 << 16, 1 << 20, 1 << 24, 1 << 26, 10 ** 6]:
        rc, txt = _text(lib, (n,), 2, prec)
        if not _smooth(n) or n in (17, 19, 23, 29, 31):
            assert rc == 3002, (n, rc)           # Bluestein has no half variant (the bare primes 17...31 run through it as well)
            continue
        assert rc == 0, (n, rc)
        for line in txt.strip().split("\n"):
            assert "half in+out" in line, (n, line)
    rc, txt = _text(lib, (256, 256, 64), 2, prec)
    assert rc == 0 and txt.count("half in+out") == 3
    # halfPrecisionMemoryOnly: exactly the launch that touches inputBuffer converts
    for inv, tag in ((-1, "half in"), (1, "half out")):
        rc, txt = _text(lib, (1 << 16,), 2, 3, inv, is_input_formatted=1, inverse_return_to_input=1)
        assert rc == 0, rc
        lines = txt.strip().split("\n")
        assert sum(tag in l for l in lines) == 1 and sum("half" in l for l in lines) == 1, txt


def test_switch_restores_the_runtime_scheduled_kernel(lib, monkeypatch):
    monkeypatch.setenv("B200FFT_NO_JIT", "1")
    rc, txt = _text(lib, (1100,), 2, 0)
    assert rc == 0 and "generic" in txt and "JIT_" not in txt
    rc, _ = _text(lib, (1024,), 2, 2)
    assert rc == 3002                            # half storage exists only as plan-time kernels
