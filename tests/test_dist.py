"""CPU, world_size 2 over gloo: the host-side logic of the N>1 path (batch sharding, max-over-ranks timing)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_batches_partitions_exactly():
    sys.path.insert(0, ROOT)
    from vkfft_b200.dist import shard_batches
    for nb in (1, 2, 7, 8, 65536, 100003):
        for world in (1, 2, 3, 4, 8):
            slabs = [shard_batches(nb, world, r) for r in range(world)]
            pos = 0
            for start, count in slabs:
                assert start == pos
                pos += count
            assert pos == nb
            assert max(c for _, c in slabs) - min(c for _, c in slabs) <= 1


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import vkfft_b200 as vk
        from vkfft_b200.dist import max_over_ranks, shard_configuration
        cfg = vk.VkFFTConfiguration(FFTdim=1, size=[4096], numberBatches=65537, device=0)
        mine, first, off = shard_configuration(cfg, world, rank)
        t = torch.tensor([mine.numberBatches, first], dtype=torch.int64)
        gathered = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        slowest = max_over_ranks(1.0 + rank, dist)
        q.put((rank, [g.tolist() for g in gathered], off, slowest, mine.device))
    finally:
        dist.destroy_process_group()


def test_two_rank_batch_sharding_over_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, g0, off0, slow0, dev0), (r1, g1, off1, slow1, dev1) = res
    assert g0 == g1 == [[32769, 0], [32768, 32769]]
    assert off0 == 0 and off1 == 32769 * 4096 * 8
    assert slow0 == slow1 == 2.0
    assert (dev0, dev1) == (0, 1)


def _dist_fft_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    import numpy as np
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from vkfft_b200.dist import DistributedFFT1D
        n1, n2 = 16, 8
        n = n1 * n2
        rng = np.random.default_rng(0)
        full = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)
        slab = torch.from_numpy(full[rank * n // world:(rank + 1) * n // world].copy())
        out = {}
        for inverse in (False, True):
            def cols(b, inv=inverse):
                a = b.numpy()
                return torch.from_numpy((np.fft.ifft(a, axis=0) * a.shape[0] if inv else np.fft.fft(a, axis=0)).astype(np.complex64))

            def rows(b, inv=inverse):
                a = b.numpy()
                return torch.from_numpy((np.fft.ifft(a, axis=1) * a.shape[1] if inv else np.fft.fft(a, axis=1)).astype(np.complex64))

            f = DistributedFFT1D(n1, n2, dist, inverse=inverse, local_cols=cols, local_rows=rows)
            y = f(slab)
            ref = np.fft.ifft(full.astype(np.complex128)) * n if inverse else np.fft.fft(full.astype(np.complex128))
            mine = ref[rank * n // world:(rank + 1) * n // world]
            out[inverse] = float(np.linalg.norm(y.numpy() - mine) / np.linalg.norm(mine))
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_distributed_four_step_index_algebra_over_gloo():
    """the sharded Four-Step (three all-to-alls, natural order in and out) with numpy standing in for the two local
    transforms: verifies packing, phases and slab ownership on 2 ranks"""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dist_fft_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, errs in res:
        assert errs[False] < 1e-6 and errs[True] < 1e-6, (rank, errs)


def _fd_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from vkfft_b200.window import exchange_fds
        r, w = os.pipe()
        msg = f"from{rank}".encode()
        os.write(w, msg * (world - 1))                 # one copy for each reader of this pipe
        got = exchange_fds(dist, [r, w])
        seen = {}
        for peer, (pr, pw) in got.items():
            seen[peer] = os.read(pr, len(msg)).decode()   # the peer's pipe, reachable through the passed descriptor
            os.close(pr); os.close(pw)
        q.put((rank, seen))
    finally:
        dist.destroy_process_group()


def test_peer_window_descriptor_exchange_over_unix_sockets():
    """host plumbing of the peer windows: every rank receives working copies of every other rank's descriptors"""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world = 3
    procs = [ctx.Process(target=_fd_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(world):
        assert res[r] == {p: f"from{p}" for p in range(world) if p != r}
