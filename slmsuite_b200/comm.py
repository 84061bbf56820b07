"""
Process-group plumbing for the multi-GPU paths -- no torch, no torch.distributed.

One process per GPU (``torchrun`` or any launcher that sets ``RANK`` / ``WORLD_SIZE`` / ``LOCAL_RANK`` /
``MASTER_ADDR`` / ``MASTER_PORT``).  The hologram loop never communicates; the job ends with ONE all-gather of the final
phases.  That collective is NCCL, loaded with ``dlopen`` inside ``libslmgs.so``
(``slmgs_comm_*`` / ``slmgs_allgather_phase``, include/slmgs.h); this module only does the rendezvous: rank 0 creates the
NCCL unique id and hands its 128 bytes to the other ranks over a TCP connection on ``MASTER_ADDR``.

The same TCP star doubles as a HOST communicator (``allgather_host`` / ``allreduce_host``): it carries the control
traffic, and it is the whole transport when the loaded library is the host-emulation build of the CPU test-suite
(which has no NCCL).
"""
import ctypes as C
import os
import pickle
import socket
import struct
import time

import numpy as np

from . import _lib

_PORT_OFFSET = 1717  # the rendezvous listens on MASTER_PORT + this (MASTER_PORT itself belongs to the launcher's store)


def _send(sock, obj):
    data = pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)
    sock.sendall(struct.pack("<Q", len(data)) + data)


def _recv(sock):
    need = struct.unpack("<Q", _recv_exact(sock, 8))[0]
    return pickle.loads(_recv_exact(sock, need))


def _recv_exact(sock, n):
    buf = bytearray(n)
    view = memoryview(buf)
    got = 0
    while got < n:
        k = sock.recv_into(view[got:], n - got)
        if k == 0:
            raise ConnectionError("peer closed the rendezvous connection")
        got += k
    return bytes(buf)


class Comm:
    """
    ``Comm(rank, world, addr, port, device=None)`` -- usually obtained from :func:`init`.

    ``rank`` / ``world``          position in the job
    ``allgather_phase(holo, n_total)``  the final all-gather of a sharded batch: host array ``(n_total, h, w)``
    ``allgather_host(array)``     list of every rank's array (host, through rank 0)
    ``allreduce_host(array)``     element-wise sum over ranks (host, through rank 0)
    ``barrier()``
    """

    def __init__(self, rank, world, addr="127.0.0.1", port=29500, device=None, timeout=120.0):
        self.rank, self.world = int(rank), int(world)
        self.device = int(device) if device is not None else self.rank
        self._peers = []   # rank 0: sockets of ranks 1..world-1 (index r-1)
        self._sock = None  # other ranks: socket to rank 0
        self._nccl = None
        self.last_allgather_ms = None
        if self.world > 1:
            self._connect(addr, int(port) + _PORT_OFFSET, timeout)

    # ---- rendezvous ----------------------------------------------------------------------------------------------
    def _connect(self, addr, port, timeout):
        if self.rank == 0:
            srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
            srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
            srv.bind((addr if addr not in ("localhost",) else "127.0.0.1", port))
            srv.listen(self.world)
            srv.settimeout(timeout)
            peers = {}
            while len(peers) < self.world - 1:
                conn, _ = srv.accept()
                conn.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
                r = _recv(conn)
                peers[int(r)] = conn
            srv.close()
            self._peers = [peers[r] for r in range(1, self.world)]
        else:
            deadline = time.time() + timeout
            while True:
                try:
                    s = socket.create_connection((addr, port), timeout=5.0)
                    break
                except OSError:
                    if time.time() > deadline:
                        raise
                    time.sleep(0.05)
            s.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
            s.settimeout(None)
            _send(s, self.rank)
            self._sock = s

    def close(self):
        if self._nccl is not None:
            _lib.lib().slmgs_comm_destroy(self._nccl)
            self._nccl = None
        for s in self._peers + ([self._sock] if self._sock else []):
            try:
                s.close()
            except OSError:
                pass
        self._peers, self._sock = [], None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- host collectives (through rank 0) -----------------------------------------------------------------------
    def broadcast_host(self, obj):
        if self.world == 1:
            return obj
        if self.rank == 0:
            for s in self._peers:
                _send(s, obj)
            return obj
        return _recv(self._sock)

    def allgather_host(self, array):
        """Every rank's ``array`` (shapes may differ), as a list indexed by rank."""
        if self.world == 1:
            return [np.asarray(array)]
        if self.rank == 0:
            parts = [np.asarray(array)] + [_recv(s) for s in self._peers]
            for s in self._peers:
                _send(s, parts)
            return parts
        _send(self._sock, np.asarray(array))
        return _recv(self._sock)

    def allreduce_host(self, array):
        parts = self.allgather_host(array)
        out = np.array(parts[0], copy=True)
        for p in parts[1:]:
            out += p
        return out

    def barrier(self):
        self.allgather_host(np.zeros(1, dtype=np.int8))

    # ---- device collectives (NCCL inside libslmgs.so) ---------------------------------------------------------------
    def on_device(self):
        """True when the loaded library is the CUDA build (NCCL collectives); False under the host emulation."""
        return _lib.library_path() in (None, _lib.DEFAULT_LIBRARY)

    def _nccl_comm(self):
        if self._nccl is None:
            lib = _lib.lib()
            uid = C.create_string_buffer(128)
            if self.rank == 0:
                if lib.slmgs_comm_unique_id(uid):
                    raise _lib.SlmgsError(lib.slmgs_comm_last_error().decode())
            uid_bytes = self.broadcast_host(bytes(uid.raw))
            comm = C.c_void_p()
            if lib.slmgs_comm_create(C.byref(comm), uid_bytes, self.rank, self.world, self.device):
                raise _lib.SlmgsError(lib.slmgs_comm_last_error().decode())
            self._nccl = comm
        return self._nccl

    def allgather_phase(self, holo, n_total=None, per_rank=None, shape=None, download=True):
        """
        The one collective of a sharded batch: all ranks receive the final phases of all holograms, ``(n_total, h, w)``
        on the host.  ``holo`` is this rank's ``Hologram`` / ``HologramBatch`` (or None when the rank owns nothing;
        then ``shape = (h, w)`` is needed).  Shards are contiguous blocks of ``per_rank`` holograms
        (``batch.shard_bounds``); a short or empty last shard is padded.  ``download=False`` leaves the gathered phases
        on the device (returns None; ``last_allgather_ms`` holds the device time of the collective either way).
        """
        n_local = 0 if holo is None else holo._batch_size()
        h, w = (tuple(holo.slm_shape) if holo is not None else tuple(shape))
        if per_rank is None:
            per_rank = n_local if n_total is None else -(-int(n_total) // self.world)
        if n_total is None:
            n_total = per_rank * self.world
        if self.world == 1:
            return np.asarray(holo.phase).reshape((-1, h, w))[:n_total]
        if not self.on_device():
            local = np.zeros((per_rank, h, w), dtype=np.float32)
            if n_local:
                local[:n_local] = np.asarray(holo.phase).reshape((n_local, h, w))
            return np.concatenate(self.allgather_host(local), axis=0)[:n_total]
        lib = _lib.lib()
        out = np.empty((self.world * per_rank, h, w), dtype=np.float32) if download else None
        ms = C.c_float()
        status = lib.slmgs_allgather_phase(None if holo is None else holo._ctx, self._nccl_comm(), n_local, per_rank, h * w,
                                           _lib.fptr(out) if download else None, None, C.byref(ms))
        if status:
            raise _lib.SlmgsError(lib.slmgs_comm_last_error().decode())
        self.last_allgather_ms = float(ms.value)
        return out[:n_total] if download else None

    # the two collectives of a pixel-sharded compressed spot hologram (compressed.ShardedCompressedSpotHologram)
    def allreduce_f64(self, ptr, count, on_device, device, stream_ptr=None):
        """In-place sum over ranks of ``count`` float64 values at ``ptr`` (device: NCCL ordered on ``stream_ptr``, no host
        synchronisation; host emulation: through rank 0)."""
        if self.world == 1:
            return
        if on_device:
            lib = _lib.lib()
            if lib.slmgs_comm_allreduce_f64(self._nccl_comm(), C.c_void_p(int(ptr)), int(count), C.c_void_p(int(stream_ptr or 0))):
                raise _lib.SlmgsError(lib.slmgs_comm_last_error().decode())
        else:
            a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(count,))
            a[:] = self.allreduce_host(a)

    def allgather_rows(self, local, rows_per_rank, on_device, device):
        """Concatenate the row slabs of every rank (slabs may differ in height)."""
        parts = self.allgather_host(np.asarray(local))
        return np.concatenate([p[:r] for p, r in zip(parts, rows_per_rank)], axis=0)


_default = None


def init(rank=None, world=None, addr=None, port=None, device=None):
    """Communicator from the launcher's environment (``RANK``, ``WORLD_SIZE``, ``LOCAL_RANK``, ``MASTER_ADDR``,
    ``MASTER_PORT``); a single process gets a trivial one."""
    rank = int(os.environ.get("RANK", "0")) if rank is None else rank
    world = int(os.environ.get("WORLD_SIZE", "1")) if world is None else world
    addr = os.environ.get("MASTER_ADDR", "127.0.0.1") if addr is None else addr
    port = int(os.environ.get("SLMGS_COMM_PORT", os.environ.get("MASTER_PORT", "29500"))) if port is None else port
    device = int(os.environ.get("LOCAL_RANK", str(rank))) if device is None else device
    return Comm(rank, world, addr, port, device)


def default():
    """The process-wide communicator (created from the environment on first use)."""
    global _default
    if _default is None:
        _default = init()
    return _default


def set_default(comm):
    global _default
    _default = comm
