"""GPU: the fused distributed Four-Step (peer windows + device-side barriers) with two processes.

Uses two GPUs when the box has them; on a one-GPU box both ranks map their slabs on cuda:0 -- the descriptor passing,
the flat mapping, the barriers and the sliced launches are exactly the same code, only the "peer" traffic stays on one
device.  torch.distributed (gloo) is only the host-side rendezvous here."""
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from vkfft_b200.dist import FusedDistributedFFT1D
        dev = rank % torch.cuda.device_count()
        torch.cuda.set_device(dev)
        g = torch.Generator(device="cpu").manual_seed(11)
        full = torch.view_as_complex(torch.empty(n, 2, dtype=torch.float32).uniform_(-1, 1, generator=g))
        lo, hi = rank * n // world, (rank + 1) * n // world
        f = FusedDistributedFFT1D(n, dist, dev, normalize=True)
        f.local.copy_(full[lo:hi])
        torch.cuda.synchronize()
        dist.barrier()
        f(inverse=False)
        f.check()
        ref = torch.fft.fft(full.to(torch.complex128))[lo:hi]
        got = f.local.cpu().to(torch.complex128)
        err = ((got - ref).abs().norm() / ref.abs().norm()).item()
        f(inverse=True)
        f.check()
        back = ((f.local.cpu() - full[lo:hi]).abs().norm() / full[lo:hi].abs().norm()).item()
        f.close()
        q.put((rank, err, back, None, dev))
    except Exception as e:  # noqa: BLE001
        q.put((rank, None, None, repr(e), -1))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("logn", [19, 23])   # 2 launches / 3 launches
def test_fused_distributed_fft_two_ranks(logn):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 1 << logn, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=120)
    for rank, err, back, exc, dev in res:
        assert exc is None, exc
        assert err < 1e-6, (rank, err)        # FP32 tolerance of north_star
        assert back < 1e-6, (rank, back)
    # on a box with two or more GPUs the two ranks sit on DIFFERENT devices: the peer traffic really crosses NVLink
    import torch
    if torch.cuda.device_count() >= 2:
        assert len({dev for *_, dev in res}) == 2, res


def _worker_nd(rank, world, port, shape_xyz, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from vkfft_b200.dist import FusedDistributedFFTND
        dev = rank % torch.cuda.device_count()
        torch.cuda.set_device(dev)
        g = torch.Generator(device="cpu").manual_seed(13)
        np_shape = tuple(reversed(shape_xyz))
        total = 1
        for v in shape_xyz:
            total *= v
        full = torch.view_as_complex(torch.empty(total, 2, dtype=torch.float32).uniform_(-1, 1, generator=g)).reshape(np_shape)
        sl = np_shape[0] // world
        f = FusedDistributedFFTND(shape_xyz, dist, dev, normalize=True)
        f.local.copy_(full[rank * sl:(rank + 1) * sl])
        torch.cuda.synchronize()
        dist.barrier()
        f(inverse=False)
        f.check()
        ref = torch.fft.fftn(full.to(torch.complex128))[rank * sl:(rank + 1) * sl]
        got = f.local.cpu().to(torch.complex128)
        err = ((got - ref).abs().norm() / ref.abs().norm()).item()
        f(inverse=True)
        f.check()
        mine = full[rank * sl:(rank + 1) * sl]
        back = ((f.local.cpu() - mine).abs().norm() / mine.abs().norm()).item()
        f.close()
        q.put((rank, err, back, None))
    except Exception as e:  # noqa: BLE001
        q.put((rank, None, None, repr(e)))
    finally:
        dist.destroy_process_group()


# slabs are mapped with the 2 MiB granularity of the virtual-memory API: shapes whose half is a multiple of 2 MiB
@pytest.mark.parametrize("shape_xyz", [(2048, 1024), (256, 128, 64), (64, 8192)])      # 2-D, 3-D, a Four-Step across the slabs
def test_fused_distributed_nd_two_ranks(shape_xyz):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_nd, args=(r, 2, port, shape_xyz, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=120)
    for rank, err, back, exc in res:
        assert exc is None, exc
        assert err < 1e-6, (rank, err)
        assert back < 1e-6, (rank, back)
