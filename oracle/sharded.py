"""The CPU oracle at BASELINE size: a voice bank rendered by the threaded -O3 oracle build, voices sharded
over the host cores (independent replicas summed at the end), with optional per-voice taps.

TEST INFRASTRUCTURE (like everything under oracle/): imported by tests/, by bench.py's `parity` /
`cpu_baseline` / `--impl reference` legs and by nothing in the product package.

knaster renders a graph on one audio thread (README.md:25, NOTES.md:84-87); sharding is only how the checker gets
through 16 384 voices x 10 s in seconds.  Voices are independent nodes of one flat graph
(knaster/examples/many_sines.rs:52-63), so a shard's taps are bit-identical to the same voices' taps in a
one-graph render; the bus is the f32 sum of the shard buses (a different summation tree than knaster's left
fold, graph.rs:850-864 -- compared at <= 1e-5 on normalised buses).
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from knaster_b200.graph import Graph

from .oracle import OracleProcessor

# build(graph, n_voices, voice_offset, total_voices) -> one tap node id per voice of the shard
BankBuilder = Callable[[Graph, int, int, int], List[int]]


from knaster_b200.banks import bank_builder  # noqa: E402,F401  (re-exported: the workloads by name)


class ShardedOracle:
    """n_voices of `build` split into contiguous shards, one OracleProcessor each.

    voice_offset / total_voices place the bank inside a larger one (a GPU rank's slice of configs[4]).
    tap_voices: bank-local voice indices whose pre-mix signal is recorded."""

    def __init__(self, build: BankBuilder, n_voices: int, tap_voices: Sequence[int] = (), threads: Optional[int] = None,
                 voice_offset: int = 0, total_voices: int = 0, outputs: int = 2, block_size: int = 64, sample_rate: int = 48000,
                 fast: bool = True):
        self.threads = max(1, min(threads or (os.cpu_count() or 1), n_voices))
        self.n_voices, self.outputs, self.block_size = n_voices, outputs, block_size
        total = total_voices or n_voices
        per = (n_voices + self.threads - 1) // self.threads
        self.shards = []
        self.tap_index: List[Tuple[int, int]] = []   # per requested tap: (shard, tap row inside the shard)
        taps_of = {}
        for t, v in enumerate(tap_voices):
            taps_of.setdefault(int(v) // per, []).append((t, int(v) % per))
        self.tap_index = [(-1, -1)] * len(tap_voices)
        for c in range(self.threads):
            nv = min(per, n_voices - c * per)
            if nv <= 0:
                break
            g = Graph(0, outputs, block_size, sample_rate)
            ids = build(g, nv, voice_offset + c * per, total)
            ev = g.take_events()
            proc = OracleProcessor(g, ring_buffer_size=1 << 24, fast=fast)
            for row, (t, local) in enumerate(taps_of.get(c, [])):
                proc.add_tap(ids[local], 0)
                self.tap_index[t] = (c, row)
            self.shards.append((g, ev, proc))

    def render(self, n_blocks: int, events_filter=None):
        """(bus [n_blocks, outputs, block], taps [n_taps, frames]); every shard's queued events are fed first.
        events_filter(ev) -> ev lets a caller window / shift the schedule (bench.py's repeating steps)."""

        def one(shard):
            g, ev, proc = shard
            e = ev if events_filter is None else events_filter(ev)
            g.pending_event_arrays = [e.copy()] if len(e) else []
            return proc.render(n_blocks)

        with ThreadPoolExecutor(len(self.shards)) as pool:
            res = list(pool.map(one, self.shards))
        bus = np.zeros((n_blocks, self.outputs, self.block_size), dtype=np.float32)
        for out, _ in res:   # shard buses folded left to right in f32
            bus += out
        taps = np.zeros((len(self.tap_index), n_blocks * self.block_size), dtype=np.float32)
        for t, (c, row) in enumerate(self.tap_index):
            taps[t] = res[c][1][row]
        return bus, taps


def sample_voices(n_voices: int, n_taps: int, seed: int = 7) -> List[int]:
    """n_taps distinct voices spread over the bank: first, last, and a seeded draw in between."""
    n_taps = min(n_taps, n_voices)
    r = np.random.Generator(np.random.PCG64(seed))
    picks = set([0, n_voices - 1][:n_taps])
    while len(picks) < n_taps:
        picks.add(int(r.integers(0, n_voices)))
    return sorted(picks)
