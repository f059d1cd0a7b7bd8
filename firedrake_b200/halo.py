"""Halo exchange and the process-group bootstrap.

``Halo`` mirrors ``firedrake.halo.Halo`` (reference firedrake/halo.py:87-172):
``global_to_local_begin/end`` refreshes ghost copies from their owners
(PetscSF bcast, MPI.REPLACE) and ``local_to_global_begin/end`` sums ghost
contributions into the owners (PetscSF reduce, MPI.SUM), both in place on the
Dat's buffer -- here the DEVICE buffer, over NCCL (NVLink/NVSwitch).

One process per GPU.  The NCCL unique id is created on rank 0 and broadcast
with ``torch.distributed`` (plumbing only; the engine itself never imports
torch).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib

IntType = np.int32
READ_MODES = ("replace",)


def comm_init_from_env():
    """Initialise engine + NCCL communicator from the torchrun environment
    (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT).  Returns
    (rank, world, dist) where dist is torch.distributed (gloo) or None."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    L = _lib.init(int(os.environ.get("LOCAL_RANK", "0")))
    if world == 1:
        return rank, world, None
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        dist.init_process_group("gloo")
    buf = C.create_string_buffer(128)
    if rank == 0:
        _lib.check(L.fdb_comm_get_unique_id(buf), "fdb_comm_get_unique_id")
    t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
    dist.broadcast(t, 0)
    _lib.check(L.fdb_comm_init(rank, world, bytes(t.numpy().tobytes())), "fdb_comm_init")
    return rank, world, dist


class Halo:
    """``neighbours``: list of ``(rank, send_indices, recv_indices)`` -- owned
    dofs that are ghosts on ``rank`` / my ghost dofs owned by ``rank``, in the
    same canonical order on both sides."""

    def __init__(self, neighbours, max_cdim=1):
        self.neighbours = [(int(r), np.ascontiguousarray(s, dtype=IntType),
                            np.ascontiguousarray(q, dtype=IntType)) for r, s, q in neighbours]
        self.max_cdim = max_cdim
        self._handle = None

    @property
    def handle(self):
        if self._handle is None:
            L = _lib.lib()
            n = len(self.neighbours)
            ranks = np.array([r for r, _, _ in self.neighbours], dtype=np.int32)
            sc = np.array([len(s) for _, s, _ in self.neighbours], dtype=IntType)
            rc = np.array([len(q) for _, _, q in self.neighbours], dtype=IntType)
            si = np.concatenate([s for _, s, _ in self.neighbours] + [np.zeros(0, IntType)]).astype(IntType)
            ri = np.concatenate([q for _, _, q in self.neighbours] + [np.zeros(0, IntType)]).astype(IntType)
            h = C.c_void_p()
            _lib.check(L.fdb_halo_create(n, ranks.ctypes.data, sc.ctypes.data, si.ctypes.data,
                                         rc.ctypes.data, ri.ctypes.data, self.max_cdim, C.byref(h)),
                       "fdb_halo_create")
            self._handle = h
        return self._handle

    # firedrake/halo.py:124-138
    def global_to_local_begin(self, dat, insert_mode="replace"):
        assert insert_mode == "replace"
        _lib.check(_lib.lib().fdb_halo_global_to_local_begin(self.handle, dat.device_ptr, dat.cdim))

    def global_to_local_end(self, dat, insert_mode="replace"):
        _lib.check(_lib.lib().fdb_halo_global_to_local_end(self.handle, dat.device_ptr, dat.cdim))
        dat._device_written(halo_valid=True)

    # firedrake/halo.py:140-172
    def local_to_global_begin(self, dat, insert_mode="sum"):
        assert insert_mode == "sum"
        _lib.check(_lib.lib().fdb_halo_local_to_global_begin(self.handle, dat.device_ptr, dat.cdim))

    def local_to_global_end(self, dat, insert_mode="sum"):
        _lib.check(_lib.lib().fdb_halo_local_to_global_end(self.handle, dat.device_ptr, dat.cdim))
        dat._device_written(halo_valid=False)

    def __del__(self):
        try:
            if self._handle is not None and _lib._initialised is not None:
                _lib._lib.fdb_halo_destroy(self._handle)
        except Exception:
            pass
