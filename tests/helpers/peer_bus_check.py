"""Run under torchrun (one rank per GPU): every rank renders its shard of a subtractive bank, the
buses are summed (a) over peer memory by the engine's own kernels and (b) with an NCCL reduce of the
rank-local buses; rank 0 compares both with a single-GPU render of the whole bank.  Prints
PEER_BUS_CHECK_OK on success (used by tests/test_multi_gpu_gpu.py)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from knaster_b200 import banks  # noqa: E402
from knaster_b200.multi_gpu import PeerBus, reduce_bus, shard_range  # noqa: E402
from knaster_b200.processor import AudioProcessor, AudioProcessorOptions  # noqa: E402

TOTAL, SECONDS, N_BLOCKS, CALLS = 256, 1.0, 250, 3  # three render calls of 250 blocks: slot reuse + back-pressure


def render_shard(rank, world, device, mode):
    lo, hi = shard_range(rank, world, TOTAL)
    graph, proc = AudioProcessor.new(0, 2, AudioProcessorOptions(device=device))
    banks.subtractive_bank(graph, hi - lo, SECONDS, n_notes=4, voice_offset=lo, total_voices=TOTAL)
    proc.set_blocks_per_launch(64)  # several launches per call
    peer = PeerBus(proc, N_BLOCKS) if mode == "peer" else None
    bus = torch.zeros((N_BLOCKS, 2, 64), dtype=torch.float32, device="cuda")
    outs = []
    for _ in range(CALLS):
        proc.render_device(N_BLOCKS, bus.data_ptr(), torch.cuda.current_stream().cuda_stream)
        if peer is None:
            reduce_bus(bus, dst=0, chunks=4)
        torch.cuda.synchronize()
        outs.append(bus.cpu().numpy().copy())
    if peer is not None:
        assert not peer.timed_out(), "a rank timed out waiting for peer data"
    return np.concatenate(outs)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_stream(torch.cuda.Stream())
    peer = render_shard(rank, world, local, "peer")
    nccl = render_shard(rank, world, local, "nccl")
    if rank == 0:
        graph, proc = AudioProcessor.new(0, 2, AudioProcessorOptions(device=local))
        banks.subtractive_bank(graph, TOTAL, SECONDS, n_notes=4)
        whole = np.concatenate([proc.render(N_BLOCKS) for _ in range(CALLS)])
        peak = float(np.abs(whole).max())
        e_peer, e_nccl = float(np.abs(peer - whole).max()), float(np.abs(nccl - whole).max())
        print(f"peak {peak:.4f}  |peer - single| {e_peer:.3e}  |nccl - single| {e_nccl:.3e}")
        assert peak > 1e-3
        assert e_peer <= 1e-6 and e_nccl <= 1e-6  # same voices, different summation trees
        print("PEER_BUS_CHECK_OK")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
