"""GPU: the reference's own test-suite binary, built unmodified against include/vkFFT.h and linked to libb200fft.so
(oracle/_ref/VkFFT_TestSuite_b200, `make -C oracle testsuite`), running its convolution samples and its sample_0 benchmark --
the reference's user-facing programs running on this engine through the drop-in header."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "VkFFT_TestSuite_b200")
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.exists(EXE), reason="oracle/_ref/VkFFT_TestSuite_b200 not built (needs /root/reference at build time)")]


def _run(sample):
    return subprocess.run([EXE, "-vkfft", str(sample)], capture_output=True, text=True, timeout=600)


def test_sample_52_batched_r2c_convolution_prints_the_expected_constants():
    """sample_52: all-ones 32x32 input with 2 features, kernel spectrum = (kernel*2 + feature + 1) at every frequency,
    normalised inverse -> every printed value of block (kernel, feature) is that constant"""
    out = _run(52)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    blocks = re.split(r"coordinate: \d+", out.stdout)[1:]
    assert len(blocks) >= 4
    for idx, blk in enumerate(blocks[:4]):
        vals = [float(v) for v in re.findall(r"-?\d+\.\d{6}", blk)]
        assert len(vals) >= 32 * 32
        want = float(idx + 1)
        assert max(abs(v - want) for v in vals[:32 * 32]) < 1e-4, (idx, vals[:8])


@pytest.mark.parametrize("sample", [50, 51])
def test_identity_kernel_convolution_samples_run(sample):
    out = _run(sample)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]


def test_sample_0_benchmark_reports_a_score():
    out = _run(0)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "Benchmark score VkFFT" in out.stdout
