"""Peer windows (include/b200fft.h b200fft_window_*): host-side plumbing for one process per GPU.

The C side allocates this rank's slab with the CUDA virtual-memory API and maps every rank's slab back to back into one
address range; what is left for the host is handing the POSIX file descriptors of the allocations to the other
processes.  That goes over unix-domain sockets with SCM_RIGHTS (socket.send_fds); torch.distributed is only used to
agree on the socket names and as the host barrier.  The reference has no multi-GPU mode (README.md:26-28).
"""
import ctypes
import os
import socket
import uuid

from . import _lib


def exchange_fds(dist, my_fds, tag=None):
    """Give `my_fds` (list of ints) to every other rank and return {peer: [fds...]} with the descriptors received from
    each peer (valid in this process).  Collective over the default process group."""
    world, rank = dist.get_world_size(), dist.get_rank()
    if world == 1:
        return {}
    box = [tag or uuid.uuid4().hex[:12]]
    dist.broadcast_object_list(box, src=0)
    path = lambda r: f"/tmp/b200fft_{box[0]}_{r}.sock"
    srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    try:
        if os.path.exists(path(rank)):
            os.unlink(path(rank))
        srv.bind(path(rank))
        srv.listen(world)
        dist.barrier()                       # every rank listens before anyone connects
        outs = []
        for peer in range(world):
            if peer == rank:
                continue
            c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            c.connect(path(peer))
            socket.send_fds(c, [rank.to_bytes(4, "little")], list(my_fds))
            outs.append(c)
        got = {}
        for _ in range(world - 1):
            conn, _addr = srv.accept()
            msg, fds, _flags, _a = socket.recv_fds(conn, 4, len(my_fds))
            got[int.from_bytes(msg, "little")] = list(fds)
            conn.close()
        dist.barrier()                       # everything received: the senders may close
        for c in outs:
            c.close()
        return got
    finally:
        srv.close()
        try:
            os.unlink(path(rank))
        except OSError:
            pass


class _CudaView:
    def __init__(self, ptr, nbytes, typestr, itemsize):
        self.__cuda_array_interface__ = {"shape": (nbytes // itemsize,), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class PeerWindow:
    """One flat device address range over the slabs of all ranks: slab g at base + g*slab_bytes (NVLink for g != rank)."""

    def __init__(self, slab_bytes, dist, device):
        self.L = _lib.load()
        self.dist = dist
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.device = int(device)
        self.slab_bytes = int(slab_bytes)
        gran = self.L.b200fft_window_granularity(self.device)
        if gran == 0:
            raise RuntimeError("CUDA virtual memory management is not available on this device")
        if self.slab_bytes % gran:
            raise ValueError(f"slab_bytes must be a multiple of {gran}")
        self.handle = ctypes.c_void_p()
        rc = self.L.b200fft_window_create(self.device, self.world, self.rank, self.slab_bytes, ctypes.byref(self.handle))
        if rc != 0:
            raise RuntimeError("b200fft_window_create: " + self.L.b200fft_error_string(rc).decode())
        fds = (ctypes.c_int * 2)()
        rc = self.L.b200fft_window_export(self.handle, fds)
        if rc != 0:
            raise RuntimeError("b200fft_window_export: " + self.L.b200fft_error_string(rc).decode())
        mine = [fds[0], fds[1]]
        try:
            for peer, pf in sorted(exchange_fds(dist, mine).items()):
                arr = (ctypes.c_int * 2)(pf[0], pf[1])
                rc = self.L.b200fft_window_import(self.handle, peer, arr)
                for f in pf:
                    os.close(f)
                if rc != 0:
                    raise RuntimeError(f"b200fft_window_import(peer {peer}): " + self.L.b200fft_error_string(rc).decode())
        finally:
            for f in mine:
                os.close(f)
        dist.barrier()
        self.base = int(self.L.b200fft_window_base(self.handle))
        self.local_ptr = int(self.L.b200fft_window_local(self.handle))

    def tensor(self, torch, dtype, whole=False):
        """torch view of this rank's slab (or of the whole window: peer slabs are then read/written over NVLink)"""
        item = torch.empty((), dtype=dtype).element_size()
        ts = {torch.complex64: "<c8", torch.complex128: "<c16", torch.float32: "<f4", torch.float64: "<f8"}[dtype]
        ptr, nb = (self.base, self.slab_bytes * self.world) if whole else (self.local_ptr, self.slab_bytes)
        return torch.as_tensor(_CudaView(ptr, nb, ts, item), device=f"cuda:{self.device}")

    def barrier(self, stream=None):
        rc = self.L.b200fft_window_barrier(self.handle, stream)
        if rc != 0:
            raise RuntimeError("b200fft_window_barrier: " + self.L.b200fft_error_string(rc).decode())

    def status(self):
        return self.L.b200fft_window_status(self.handle)

    def close(self):
        if self.handle:
            self.dist.barrier()
            self.L.b200fft_window_destroy(self.handle)
            self.handle = None
