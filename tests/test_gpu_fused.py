"""GPU (-m gpu): the fused Four-Step launch (csrc/fused4.cuh, opt-in with B200FFT_FUSED4=1) against the oracle and against the
two-launch plan.

One persistent launch runs both passes of N = n1 * n2 with the intermediate in an L2-resident scratch: groups of CTAs walk
their sequences in phases, tiles arrive by TMA (tensor-map copies for the strided pass), two counters per group order the
passes.  Both passes execute the stage code of the stand-alone kernels, so whenever the two plans use the same split and
radix schedules the results must agree bit for bit -- a missed dependency or a stale tile shows up as a difference."""
import os
import re

import numpy as np
import pytest

import vkfft_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import torch
    assert torch.cuda.is_available(), "these tests need a GPU"
    return torch


def _plan_and_run(torch, x, n, batch, inverse, fused, **cfgkw):
    import vkfft_b200 as vk
    old = os.environ.get("B200FFT_FUSED4")
    if fused:
        os.environ["B200FFT_FUSED4"] = "1"
    else:
        os.environ.pop("B200FFT_FUSED4", None)
    try:
        app = vk.VkFFTApplication()
        rc = vk.initializeVkFFT(app, vk.VkFFTConfiguration(FFTdim=1, size=[n], numberBatches=batch, device=0, **cfgkw))
    finally:
        if old is None:
            os.environ.pop("B200FFT_FUSED4", None)
        else:
            os.environ["B200FFT_FUSED4"] = old
    assert rc == 0, vk.getVkFFTErrorString(rc)
    txt = vk.planInfo(app)["inverse" if inverse == 1 else "forward"]
    t = torch.from_numpy(x).cuda()
    try:
        assert vk.VkFFTAppend(app, inverse, vk.VkFFTLaunchParams(buffer=t)) == 0
        torch.cuda.synchronize()
        return t.cpu().numpy(), txt
    finally:
        vk.deleteVkFFT(app)


@pytest.mark.parametrize("logn", [15, 16, 17, 18, 19, 20, 21])
@pytest.mark.parametrize("inverse", [-1, 1])
def test_fused_four_step_matches_oracle_and_two_launch_plan(gpu, logn, inverse):
    n = 1 << logn
    batch = max(6, (1 << 25) // n) + 1            # odd: the last phase of some groups has no partner sequence
    x = orc.random_input((batch, n), np.complex64, seed=logn)
    fused, txt = _plan_and_run(gpu, x, n, batch, inverse, True)
    assert "fused with the next launch" in txt, txt
    assert orc.error_metrics(fused, orc.c2c(x, 1, inverse == 1))["l2_rel"] < 1e-6
    plain, txt2 = _plan_and_run(gpu, x, n, batch, inverse, False)
    assert "fused" not in txt2
    same_split = [l.split(" n=")[1].split()[0] for l in txt.strip().split("\n")] == [l.split(" n=")[1].split()[0] for l in txt2.strip().split("\n")]
    same_radices = [r.replace(", ", "x") for r in re.findall(r"B2_R\(([^)]*)\)", txt)] == re.findall(r"\[([0-9x]+)\]", txt2)
    if same_split and same_radices:
        assert np.array_equal(fused.view(np.float32), plain.view(np.float32))
    else:
        assert orc.error_metrics(fused, plain)["l2_rel"] < 1e-6


def test_fused_repeated_executions_and_normalised_round_trip(gpu):
    """the control block is reset by every launch: run the same plan many times back to back, forward and inverse"""
    import vkfft_b200 as vk
    torch = gpu
    n, batch = 1 << 16, 257
    x = orc.random_input((batch, n), np.complex64, seed=3)
    os.environ["B200FFT_FUSED4"] = "1"
    try:
        app = vk.VkFFTApplication()
        assert vk.initializeVkFFT(app, vk.VkFFTConfiguration(FFTdim=1, size=[n], numberBatches=batch, device=0, normalize=1)) == 0
    finally:
        os.environ.pop("B200FFT_FUSED4", None)
    assert "fused" in vk.planInfo(app)["forward"]
    t = torch.from_numpy(x).cuda()
    lp = vk.VkFFTLaunchParams(buffer=t)
    for _ in range(5):
        assert vk.VkFFTAppend(app, -1, lp) == 0
        assert vk.VkFFTAppend(app, 1, lp) == 0
    torch.cuda.synchronize()
    vk.deleteVkFFT(app)
    assert orc.error_metrics(t.cpu().numpy(), x)["l2_rel"] < 3e-6          # ten transforms


def test_fused_falls_back_to_two_launches_on_unaligned_buffers(gpu):
    """TMA needs 16-byte aligned sources: with an 8-byte bufferOffset the same plan runs its two stand-alone launches"""
    import vkfft_b200 as vk
    torch = gpu
    n, batch = 1 << 15, 16
    x = orc.random_input((batch, n), np.complex64, seed=5)
    raw = torch.zeros(batch * n + 1, dtype=torch.complex64, device="cuda")
    raw[1:] = torch.from_numpy(x.reshape(-1)).cuda()
    os.environ["B200FFT_FUSED4"] = "1"
    try:
        app = vk.VkFFTApplication()
        assert vk.initializeVkFFT(app, vk.VkFFTConfiguration(FFTdim=1, size=[n], numberBatches=batch, device=0, specifyOffsetsAtLaunch=1)) == 0
    finally:
        os.environ.pop("B200FFT_FUSED4", None)
    assert "fused" in vk.planInfo(app)["forward"]
    assert vk.VkFFTAppend(app, -1, vk.VkFFTLaunchParams(buffer=raw, bufferOffset=8)) == 0
    torch.cuda.synchronize()
    got = raw[1:].cpu().numpy().reshape(batch, n)
    vk.deleteVkFFT(app)
    assert orc.error_metrics(got, orc.c2c(x, 1))["l2_rel"] < 1e-6
