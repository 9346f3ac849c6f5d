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


def test_sample_2_half_precision_benchmark_reports_a_score():
    """sample_2: the reference's half-precision benchmark (halfPrecision = 1, N = 8 ... 2^26, 512 MiB buffers) on this engine; the
    half-storage kernels are instantiated at plan time, so without libnvrtc the program stops with the unsupported-length code"""
    from vkfft_b200 import _lib
    if not _lib.load().b2_jit_available():
        pytest.skip("half-storage kernels are instantiated at plan time: libnvrtc is not loadable here")
    out = _run(2)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    m = re.search(r"Benchmark score VkFFT: (\d+)", out.stdout)
    assert m and int(m.group(1)) > 100000, out.stdout[-1000:]
    assert len(re.findall(r"VkFFT System: \d+ \d+x\d+", out.stdout)) == 24
