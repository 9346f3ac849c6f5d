"""CPU: the distributed Four-Step plans (desc.dist_world > 1) on the kernel-body emulation.

Two host arrays stand for the two peer windows (sequence + temp, slab g = elements [g*N/R, (g+1)*N/R)).  Every rank's
plan is built, and the launches are played in orders the plan's barriers allow; the result must be the FFT of the
whole sequence (numpy double).  This pins the slicing / base-offset / phase-coordinate algebra without a GPU; the
2-GPU run of the same plans over NVLink is tools/dist_fused_check.py.
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
import emu  # noqa: E402


def _play(n, world, inverse, env, order, prec=0, normalize=0):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        cdt = np.complex64 if prec == 0 else np.complex128
        rng = np.random.default_rng(n + world)
        x = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(cdt)
        buf, tmp = x.copy(), np.zeros(n, dtype=cdt)
        descs = [emu.make_desc((n,), 1, prec, user_temp_buffer=1, dist_world=world, dist_rank=r, normalize=normalize)
                 for r in range(world)]
        rc, npass, sync = emu.exec_plan_pass(descs[0], inverse, buf, tmp, -1)
        assert rc == 0, rc
        assert sync[0], "a rank may only start once every slab holds its input"
        # segments of launches between barriers
        segs, cur = [], []
        for p in range(npass):
            if sync[p] and cur:
                segs.append(cur)
                cur = []
            cur.append(p)
        segs.append(cur)
        for seg in segs:
            ranks = list(range(world)) if order == "up" else list(range(world - 1, -1, -1))
            for r in ranks:                 # no barrier inside a segment: a rank may run through it alone
                for p in seg:
                    rc, _, _ = emu.exec_plan_pass(descs[r], inverse, buf, tmp, p)
                    assert rc == 0, rc
        ref = (np.fft.ifft(x.astype(np.complex128)) * n) if inverse == 1 else np.fft.fft(x.astype(np.complex128))
        if normalize and inverse == 1:
            ref = ref / n
        err = np.linalg.norm(buf - ref) / np.linalg.norm(ref)
        return err, npass, len(segs)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


CASES = [
    # n, world, env, launches, barrier-separated segments
    (1 << 15, 2, {}, 2, 2),
    (1 << 15, 4, {"B200FFT_FOUR_STEP_SPLIT": "64,512"}, 2, 2),
    (1 << 15, 8, {"B200FFT_FOUR_STEP_SPLIT": "32,16,64"}, 3, 2),
    (1 << 12, 2, {"B200FFT_MAX_SINGLE_PASS": "1024", "B200FFT_FOUR_STEP_SPLIT": "16,16,16"}, 3, 2),
    (3 * 5 * 7 * 8 * 9 * 4, 2, {"B200FFT_FOUR_STEP_SPLIT": "120,252"}, 2, 2),
]


@pytest.mark.parametrize("n,world,env,launches,segments", CASES)
@pytest.mark.parametrize("inverse", [-1, 1])
def test_distributed_four_step_all_ranks(n, world, env, launches, segments, inverse):
    for order in ("up", "down"):
        err, npass, nseg = _play(n, world, inverse, env, order)
        assert npass == launches and nseg == segments
        assert err < 2e-6, err


def test_distributed_double_and_normalize():
    err, _, _ = _play(1 << 13, 2, 1, {"B200FFT_MAX_SINGLE_PASS": "1024"}, "up", prec=1, normalize=1)
    assert err < 1e-13, err


def test_distributed_plan_rejects_what_it_cannot_shard():
    buf = np.zeros(4096, dtype=np.complex64)
    d = emu.make_desc((10125,), 1, 0, user_temp_buffer=1, dist_world=2, dist_rank=0)
    rc, _, _ = emu.exec_plan_pass(d, -1, buf, buf.copy(), -1)
    assert rc == 3002          # odd length: two ranks cannot split any Four-Step factor evenly
    d = emu.make_desc((509,), 1, 0, user_temp_buffer=1, dist_world=2, dist_rank=0)
    assert emu.exec_plan_pass(d, -1, buf, buf.copy(), -1)[0] == 3002   # Bluestein lengths are not distributed
    d = emu.make_desc((1 << 15,), 1, 0, dist_world=2, dist_rank=0)
    assert emu.exec_plan_pass(d, -1, buf, buf.copy(), -1)[0] == 2006   # windows must be supplied (temp included)
    d = emu.make_desc((1 << 15,), 2, 0, user_temp_buffer=1, dist_world=2, dist_rank=0)
    assert emu.exec_plan_pass(d, -1, buf, buf.copy(), -1)[0] == 3002   # batches are sharded whole, not distributed
    d = emu.make_desc((1 << 15,), 1, 0, user_temp_buffer=1, dist_world=2, dist_rank=2)
    assert emu.exec_plan_pass(d, -1, buf, buf.copy(), -1)[0] == 1002


# ---- distributed N-D transforms: slabs along the last dimension ---------------------------------------
def _play_nd(shape_xyz, world, inverse, order, env=None, normalize=0):
    """every rank's launches of the slab plan on one host array standing for the peer window (+ a temp window)"""
    env = env or {}
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        total = int(np.prod(shape_xyz))
        rng = np.random.default_rng(total + world)
        x = (rng.uniform(-1, 1, total) + 1j * rng.uniform(-1, 1, total)).astype(np.complex64)
        buf, tmp = x.copy(), np.zeros(total, dtype=np.complex64)
        descs = [emu.make_desc(shape_xyz, 1, 0, user_temp_buffer=1, dist_world=world, dist_rank=r, normalize=normalize) for r in range(world)]
        rc, npass, sync = emu.exec_plan_pass(descs[0], inverse, buf, tmp, -1)
        assert rc == 0, rc
        segs, cur = [], []
        for p in range(npass):
            if sync[p] and cur:
                segs.append(cur)
                cur = []
            cur.append(p)
        segs.append(cur)
        for seg in segs:
            for r in (range(world) if order == "up" else range(world - 1, -1, -1)):
                for p in seg:
                    rc, _, _ = emu.exec_plan_pass(descs[r], inverse, buf, tmp, p)
                    assert rc == 0, rc
        nd = x.astype(np.complex128).reshape(tuple(reversed(shape_xyz)))
        ref = np.fft.ifftn(nd) * total if inverse == 1 else np.fft.fftn(nd)
        if normalize and inverse == 1:
            ref = ref / total
        return np.linalg.norm(buf.reshape(ref.shape) - ref) / np.linalg.norm(ref), npass, len(segs)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("shape,world", [((64, 32), 2), ((128, 64), 4), ((32, 16, 8), 2), ((64, 32, 16), 4), ((48, 40, 12), 2),
                                         ((256, 8), 8)])
@pytest.mark.parametrize("inverse", [-1, 1])
def test_distributed_nd_slab_plans(shape, world, inverse):
    for order in ("up", "down"):
        err, npass, nseg = _play_nd(shape, world, inverse, order)
        assert npass == len(shape) and nseg == 2      # local axes | barrier | the axis across the slabs (inverse: the other way round)
        assert err < 2e-6, err


def test_distributed_nd_long_last_axis_runs_as_strided_four_step():
    """the axis across the slabs is too long for one strided launch: two launches along the stride, still one barrier segment"""
    for inverse in (-1, 1):
        err, npass, nseg = _play_nd((32, 4096), 2, inverse, "up", env={"B200FFT_MAX_SINGLE_PASS": "1024"})
        assert npass == 3 and nseg == 2
        assert err < 2e-6, err


def test_distributed_nd_rejects_what_it_cannot_slice():
    d = emu.make_desc((64, 30), 1, 0, user_temp_buffer=1, dist_world=4, dist_rank=0)      # 30 rows over 4 ranks
    assert emu.exec_plan_pass(d, -1, np.zeros(64 * 30, np.complex64), np.zeros(64 * 30, np.complex64), -1)[0] == 3002
