"""The path's only collective -- the all-GPU sum of the summed cost -- over peer memory.

Host side of include/mrnnt_b200/peer_reduce.cuh: every rank owns a few bytes of device memory (its "board"),
mapped into all peers with CUDA IPC; the gradient kernel of each rank stores its cost sum into every board over
NVLink when the costs are final and reads the world's sums from its own board at its end.  No collective library
kernel runs in the step.  ``torch.distributed`` is used once, at set-up, to hand the 64-byte IPC handles round.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional

import torch

from . import _lib


class PeerBoards:
    """The boards of one process group, as seen from this rank.  Collective: every rank of `group` constructs one.

    ``local=[ptr0, ptr1, ...]`` (tests, one process driving several handles): plain device pointers instead of IPC.
    """

    def __init__(self, group=None, device: Optional[torch.device] = None):
        import torch.distributed as dist
        self._lib = _lib.load()
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.epoch = 0            # exchanges these boards have carried (the same on every rank)
        self._own = ctypes.c_void_p()
        self._opened: List[int] = []
        handle = (ctypes.c_ubyte * 64)()
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        with torch.cuda.device(dev):
            _lib.check(self._lib.mrnnt_peer_board_create(self.world, ctypes.byref(self._own), handle),
                       "mrnnt_peer_board_create")
            handles: List[Optional[bytes]] = [None] * self.world
            dist.all_gather_object(handles, bytes(handle), group=group)
            ptrs = []
            for r in range(self.world):
                if r == self.rank:
                    ptrs.append(self._own.value)
                    continue
                p = ctypes.c_void_p()
                buf = (ctypes.c_ubyte * 64).from_buffer_copy(handles[r])
                _lib.check(self._lib.mrnnt_peer_board_open(buf, ctypes.byref(p)), "mrnnt_peer_board_open")
                self._opened.append(p.value)
                ptrs.append(p.value)
            self.ptrs = ptrs
            dist.barrier(group=group)   # nobody publishes into a board that is not mapped yet

    @classmethod
    def local(cls, world: int, device: torch.device) -> List["PeerBoards"]:
        """`world` ranks inside ONE process on one device (tests): the boards are plain device allocations."""
        lib = _lib.load()
        owns = []
        with torch.cuda.device(device):
            for _ in range(world):
                p = ctypes.c_void_p()
                _lib.check(lib.mrnnt_peer_board_create(world, ctypes.byref(p), None), "mrnnt_peer_board_create")
                owns.append(p)
        out = []
        for r in range(world):
            b = cls.__new__(cls)
            b._lib, b.rank, b.world, b.epoch = lib, r, world, 0
            b._own, b._opened, b.ptrs = owns[r], [], [o.value for o in owns]
            out.append(b)
        return out

    def c_array(self):
        return (ctypes.c_void_p * self.world)(*self.ptrs)

    def close(self) -> None:
        for p in self._opened:
            self._lib.mrnnt_peer_board_close(ctypes.c_void_p(p))
        self._opened = []
        if self._own is not None and self._own.value:
            self._lib.mrnnt_peer_board_destroy(self._own)
            self._own = None
