"""CPU: the fused Four-Step launch (csrc/fused4.cuh) on the kernel-body emulation.

One CTA (a group of one) walks every tile of the static schedule, so the tile bookkeeping -- phases, the two scratch slots,
the tile counters, the TMA-style tile copies (synchronous here) and the double-buffered tile loop -- is exercised exactly as
on the device, minus the concurrency.  Results are compared with the oracle and, bit for bit, with the two-launch plan
(both run the same stage code)."""
import os

import numpy as np
import pytest

import emu
import vkfft_oracle as orc


class env:
    def __init__(self, **kw):
        self.kw = {k: str(v) for k, v in kw.items()}

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kw}
        os.environ.update(self.kw)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def run(shape, batch, inv, x, **kw):
    kw.setdefault("B200FFT_FUSED4", 1)           # opt-in while the two-launch plan is the faster one on the device
    with env(**kw):
        d = emu.make_desc(shape, batch)
        rc, txt = emu.describe(d, inv)
        buf = x.copy()
        rc2, npass = emu.exec_plan(d, inv, buf)
    assert rc == 0 and rc2 == 0
    return buf, txt


# (log2 N, batch, extra environment).  The emulation runs one CTA at a time, so the launch is played by ONE group of one
# CTA that walks every sequence in phase order (pass B of sequence j-1, then pass A of sequence j, alternating between the two
# scratch slots); odd and even numbers of sequences, one sequence, every fused pair up to 2^17
CASES = [
    (15, 5, dict()),
    (15, 6, dict(B200FFT_FUSED_GROUP=4)),
    (15, 1, dict()),
    (15, 2, dict()),
    (16, 3, dict()),
    (16, 4, dict(B200FFT_FUSED_GROUP=64)),
    (17, 2, dict()),
]


@pytest.mark.parametrize("logn,batch,kw", CASES)
@pytest.mark.parametrize("inv", [-1, 1])
def test_fused_matches_oracle_and_two_launch_plan(logn, batch, kw, inv):
    n = 1 << logn
    x = orc.random_input((batch, n), np.complex64, seed=logn * 100 + batch)
    fused, txt = run((n,), batch, inv, x, **kw)
    assert "fused with the next launch" in txt, txt
    plain, txt2 = run((n,), batch, inv, x, B200FFT_NO_FUSED4=1)
    assert "fused" not in txt2
    assert orc.error_metrics(fused, orc.c2c(x, 1, inv == 1))["l2_rel"] < 8e-7
    # same stage code on the same data: the two plans agree bit for bit when the two-launch plan picks the same split AND
    # the same radix schedules as the fused pair (its default kernels are re-ranked from GPU timings now and then)
    import re
    f1 = [l.split(" n=")[1].split()[0] for l in txt.strip().split("\n")]
    f2 = [l.split(" n=")[1].split()[0] for l in txt2.strip().split("\n")]
    fused_radices = [r.replace(", ", "x") for r in re.findall(r"B2_R\(([^)]*)\)", txt)]
    plain_radices = re.findall(r"\[([0-9x]+)\]", txt2)
    if f1 == f2 and fused_radices == plain_radices:
        assert np.array_equal(fused.view(np.float32), plain.view(np.float32))
    else:
        assert orc.error_metrics(fused, plain)["l2_rel"] < 8e-7


def test_fused_normalized_inverse_round_trip():
    n, batch = 1 << 15, 4
    x = orc.random_input((batch, n), np.complex64, seed=7)
    d = emu.make_desc((n,), batch, normalize=1)
    buf = x.copy()
    with env(B200FFT_FUSED4=1):
        assert "fused" in emu.describe(d, -1)[1]
        assert emu.exec_plan(d, -1, buf)[0] == 0
        assert emu.exec_plan(d, 1, buf)[0] == 0
    assert orc.error_metrics(buf, x)["l2_rel"] < 8e-7


def test_fused_inside_a_2d_plan():
    """the long axis of a 2-D transform: sequences are the rows, the other axis runs as ordinary strided launches"""
    nx, ny, batch = 1 << 15, 4, 2
    x = orc.random_input((batch, ny, nx), np.complex64, seed=11)
    d = emu.make_desc((nx, ny), batch)
    buf = x.copy()
    with env(B200FFT_FUSED4=1):
        rc, txt = emu.describe(d, -1)
        assert rc == 0 and "fused with the next launch" in txt
        assert emu.exec_plan(d, -1, buf)[0] == 0
    assert orc.error_metrics(buf, orc.c2c(x, 2, False))["l2_rel"] < 8e-7


def test_unfused_when_disabled_or_unsupported():
    d = emu.make_desc((1 << 16,), 2)
    assert "fused" not in emu.describe(d, -1)[1]                     # opt-in
    with env(B200FFT_FUSED4=1, B200FFT_NO_FUSED4=1):
        assert "fused" not in emu.describe(d, -1)[1]
    # FP64 has no fused kernels (yet): plain two launches
    d64 = emu.make_desc((1 << 16,), 2, prec=1)
    with env(B200FFT_FUSED4=1):
        assert "fused" not in emu.describe(d64, -1)[1]
