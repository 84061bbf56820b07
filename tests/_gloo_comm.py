"""Test plumbing: the communicator interface of slmsuite_b200.comm over torch.distributed (gloo) -- the world_size-2
CPU tests drive the sharded host logic through it (the "device" buffers of the emulation library are host memory).
The product's own communicator (slmsuite_b200/comm.py) needs no torch."""
import ctypes as C

import numpy as np


class GlooComm:
    def __init__(self, group=None):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.group = torch, dist, group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = 0

    def allgather_host(self, array):
        torch = self.torch
        a = np.ascontiguousarray(array)
        shapes = [None] * self.world
        self.dist.all_gather_object(shapes, (a.shape, a.dtype.str), group=self.group)
        n = max(int(np.prod(s)) for s, _ in shapes)
        pad = np.zeros(n, dtype=a.dtype)
        pad[:a.size] = a.ravel()
        t = torch.from_numpy(pad)
        out = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t, group=self.group)
        return [o.numpy()[:int(np.prod(s))].reshape(s) for o, (s, _) in zip(out, shapes)]

    def allgather_phase(self, holo, n_total=None, per_rank=None, shape=None):
        n_local = 0 if holo is None else holo._batch_size()
        h, w = (tuple(holo.slm_shape) if holo is not None else tuple(shape))
        per = per_rank if per_rank is not None else -(-int(n_total) // self.world)
        local = np.zeros((per, h, w), dtype=np.float32)
        if n_local:
            local[:n_local] = np.asarray(holo.phase).reshape((n_local, h, w))
        return np.concatenate(self.allgather_host(local), axis=0)[:n_total]

    def allreduce_f64(self, ptr, count, on_device, device, stream_ptr=None):
        a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(count,))
        t = self.torch.from_numpy(a)
        self.dist.all_reduce(t, group=self.group)

    def allgather_rows(self, local, rows_per_rank, on_device, device):
        parts = self.allgather_host(np.asarray(local))
        return np.concatenate([p[:r] for p, r in zip(parts, rows_per_rank)], axis=0)
