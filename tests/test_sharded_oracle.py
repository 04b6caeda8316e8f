"""The sharded oracle (oracle/sharded.py) against the one-graph oracle: taps bit-identical, bus within the
summation-order budget.  CPU only."""
import numpy as np

from knaster_b200.graph import Graph
from oracle.oracle import OracleProcessor
from oracle.sharded import ShardedOracle, bank_builder, sample_voices


def test_sharded_oracle_equals_one_graph_oracle():
    for workload, nv in (("subtractive", 23), ("additive", 17), ("fm", 9), ("subtractive_seg", 11)):
        build = bank_builder(workload, 0.25)
        voices = sample_voices(nv, 5)
        g = Graph(0, 2, 64, 48000)
        ids = build(g, nv, 0, nv)
        one = OracleProcessor(g, ring_buffer_size=1 << 22)
        for v in voices:
            one.add_tap(ids[v], 0)
        ref, ref_taps = one.render(188)
        bus, taps = ShardedOracle(build, nv, tap_voices=voices, threads=4).render(188)
        assert np.array_equal(taps, ref_taps), workload
        assert np.abs(bus - ref).max() <= 1e-6, workload
        assert np.abs(ref_taps).max() > 0


def test_sharded_oracle_rank_slice_matches_the_whole_bank():
    # a GPU rank's slice of a larger bank: voice_offset / total_voices select the same random stream
    build = bank_builder("subtractive", 0.2)
    whole = ShardedOracle(build, 12, tap_voices=[6, 11], threads=3)
    part = ShardedOracle(build, 6, tap_voices=[0, 5], threads=2, voice_offset=6, total_voices=12)
    _, tw = whole.render(150)
    _, tp = part.render(150)
    assert np.array_equal(tw, tp)
