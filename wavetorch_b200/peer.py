"""Peer-memory plumbing for the one collective of the batch-sharded path.

`PeerGradReducer` owns a symmetric exchange buffer (torch symmetric memory: every rank's allocation is mapped into every
process of the box) and calls `wt_peer_allreduce` (csrc/wt_peer.cu): one kernel that stores this rank's gradient into all
peers over NVLink, exchanges epoch flags and sums the world's contributions in rank order.  PyTorch supplies the memory
mapping and the rendezvous only; the data movement and the reduction are the kernel's.
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib


class PeerGradReducer:
    """Sum-all-reduce of float32 vectors of up to `capacity` elements across the ranks of `group` (one box, <= 16 GPUs)."""

    def __init__(self, capacity, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if self.world > 16:
            raise RuntimeError("wavetorch_b200: peer all-reduce supports at most 16 ranks on one box")
        self.device = torch.device(device)
        self.capacity = int(capacity)
        gather_floats = 2 * self.world * self.capacity
        self.flags_offset = ((gather_floats * 4 + 255) // 256) * 256
        total_floats = self.flags_offset // 4 + 64
        self.buf = symm_mem.empty(total_floats, dtype=torch.float32, device=self.device)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, self.group)
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        if len(ptrs) != self.world or any(p == 0 for p in ptrs):
            raise RuntimeError("wavetorch_b200: symmetric memory rendezvous did not map every peer")
        self.peer_base = (ctypes.c_uint64 * self.world)(*ptrs)
        self.state = torch.zeros(2, dtype=torch.int32, device=self.device)
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)          # every rank's buffer is zeroed before anybody's first store can land

    def all_reduce(self, flat, scale=1.0):
        """Returns the sum over ranks of `scale * flat` (float32, contiguous, <= capacity elements)."""
        lib = _lib.load()
        assert flat.dtype == torch.float32 and flat.is_contiguous() and flat.device == self.device
        out = torch.empty_like(flat)
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        with torch.cuda.device(self.device):
            st = lib.wt_peer_allreduce(self.world, self.rank, flat.numel(), self.capacity, float(scale), _lib.ptr(flat),
                                       _lib.ptr(out), self.peer_base, self.flags_offset, _lib.ptr(self.state), idx,
                                       _lib.stream_ptr(self.device))
        _lib.check(st, "wt_peer_allreduce")
        _lib.count_launches(1)
        return out
