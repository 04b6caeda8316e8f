"""Multi-GPU plumbing: voices shard across ranks (one process per GPU), each rank renders its
slice with its own plan, and the rank-local stereo mix bus is summed onto rank 0 with a
``torch.distributed`` reduce (NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests).

knaster has no counterpart (single process, single audio thread: README.md:25); the reduce is
the distributed form of the graph-out Add chain (graph.rs:850-864).  There is no other data-path
collective: voices are independent.

Two ways to sum the bus:

* ``PeerBus`` (GPUs of one NVLink/NVSwitch box): the engine's own bus-reduction kernel stores each
  rank's bus straight into rank 0's memory and rank 0 folds the slots as the launches arrive
  (include/knaster_gpu.h, kgpu_plan_set_peer_bus).  torch only provides the peer mapping
  (``torch.distributed._symmetric_memory``); no collective runs in the data path.
* ``reduce_bus``: a ``torch.distributed`` reduce of the rank-local bus (NCCL; gloo on CPU) -- the
  fallback when peer memory is unavailable, and what the CPU tests exercise.
"""
from __future__ import annotations

from typing import Tuple


def shard_range(rank: int, world: int, total_voices: int) -> Tuple[int, int]:
    """Contiguous voice slice [begin, end) of `rank`: [r*V/G, (r+1)*V/G)."""
    return (rank * total_voices) // world, ((rank + 1) * total_voices) // world


def reduce_bus(bus, dst: int = 0, chunks: int = 1):
    """Sum the rank-local bus tensor [n_blocks, channels, block] onto rank `dst`, in `chunks`
    messages along the block axis (the message is tiny -- 512 B per block -- so the cost is
    launch latency, not bandwidth).  Enqueued on the current stream; no host sync."""
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return bus
    n = bus.shape[0]
    chunks = max(1, min(chunks, n))
    for c in range(chunks):
        b0, b1 = (c * n) // chunks, ((c + 1) * n) // chunks
        if b1 > b0:
            dist.reduce(bus[b0:b1], dst=dst)
    return bus


class PeerBus:
    """The multi-GPU mix bus over peer memory.  Allocate once per processor (every rank, same
    arguments), then ``proc.render*`` sums all ranks' buses into RANK 0's output buffer.

    torch symmetric memory gives every rank a device pointer to rank 0's allocation; the buffer
    layout and the kernels that use it belong to the engine (kgpu_plan_set_peer_bus)."""

    def __init__(self, proc, n_blocks: int, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        from . import _ffi

        group = group or dist.group.WORLD
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        lib = _ffi.lib()
        floats = n_blocks * proc.block_size() * proc.outputs()
        nbytes = int(lib.kgpu_peer_bus_bytes(self.world, floats))
        self.buf = symm_mem.empty(nbytes // 4, dtype=torch.float32, device=torch.device("cuda", torch.cuda.current_device()))
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, group.group_name)
        torch.cuda.synchronize()
        dist.barrier(group)  # every rank's zero-fill is done before anyone's first store
        self.proc = proc
        proc.set_peer_bus(self.rank, self.world, int(self.hdl.buffer_ptrs[0]), nbytes)

    def timed_out(self) -> bool:
        return self.proc.peer_bus_timed_out()

    def close(self) -> None:
        self.proc.set_peer_bus(0, 0, 0, 0)
