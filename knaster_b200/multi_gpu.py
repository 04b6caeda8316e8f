"""Multi-GPU plumbing: voices shard across ranks (one process per GPU), each rank renders its
slice with its own plan, and the rank-local stereo mix bus is summed onto rank 0 with a
``torch.distributed`` reduce (NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests).

knaster has no counterpart (single process, single audio thread: README.md:25); the reduce is
the distributed form of the graph-out Add chain (graph.rs:850-864).  There is no other data-path
collective: voices are independent.
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
