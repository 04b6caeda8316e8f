"""World-size-2 gloo test of the multi-GPU host logic on CPU: the voice shards partition the
bank, and reducing the per-rank buses reproduces the full-bank bus.  The per-rank render is done
by the oracle here (no GPU in this test); the GPU ranks run the same sharding + reduce code."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from knaster_b200 import banks
from knaster_b200.graph import Graph
from knaster_b200.multi_gpu import reduce_bus, shard_range
from oracle.oracle import OracleProcessor

TOTAL, SECONDS, BLOCKS = 24, 0.25, 187


def render_shard(rank, world):
    b, e = shard_range(rank, world, TOTAL)
    g = Graph(0, 2, 64, 48000)
    banks.subtractive_bank(g, e - b, SECONDS, n_notes=2, voice_offset=b, total_voices=TOTAL)
    out, _ = OracleProcessor(g, ring_buffer_size=1 << 20).render(BLOCKS)
    return out


def worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    bus = torch.from_numpy(render_shard(rank, world))
    reduce_bus(bus, dst=0, chunks=5)
    if rank == 0:
        q.put(bus.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


def test_shards_partition_the_bank():
    for world in (1, 2, 3, 8):
        r = [shard_range(k, world, 131072) for k in range(world)]
        assert r[0][0] == 0 and r[-1][1] == 131072
        assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
    # the same seed + voice_offset/total_voices gives every shard its slice of ONE bank
    from knaster_b200 import ugens as U

    def saw_args(graph):
        return [n.ugen.args for n in graph.nodes if n.ugen.kind == U.KIND_POLYBLEP]

    full = Graph(0, 2, 64, 48000)
    banks.subtractive_bank(full, TOTAL, SECONDS, n_notes=2)
    n_full = len(full.take_events())
    n, saws = 0, []
    for k in range(2):
        b, e = shard_range(k, 2, TOTAL)
        g = Graph(0, 2, 64, 48000)
        banks.subtractive_bank(g, e - b, SECONDS, n_notes=2, voice_offset=b, total_voices=TOTAL)
        n += len(g.take_events())
        saws += saw_args(g)
    assert n == n_full and saws == saw_args(full)


def test_two_rank_gloo_reduce_matches_full_bank():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = render_shard(0, 1)
    assert np.abs(ref).max() > 1e-3
    assert np.abs(got - ref).max() <= 1e-5
