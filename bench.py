#!/usr/bin/env python
"""bench.py -- voice-samples/sec of the batched render path on N B200s (BASELINE.json metric).

A "step" renders `--seconds` (default 10 s = 7500 blocks of 64 frames @ 48 kHz) of the
subtractive polysynth bank -- 16384 voices per GPU of PolyBlep saw -> SvfFilter lowpass ->
EnvAsr -> VCA with sample-accurate note events (BASELINE.json configs[2]; with N GPUs the
voices are sharded 16384/GPU and the stereo mix bus is reduced with NCCL: configs[4]).

  value     device-resident: events already compiled + uploaded; times kernels (+ NCCL reduce)
  e2e       through the C ABI with HOST buffers: kgpu_plan_push_events(host events) +
            kgpu_render(host_out): control simulation, H2D of events, kernels, D2H of audio
  --impl reference   the CPU oracle (C++ restatement of knaster's render path; knaster is Rust
            and cannot be built in this image) on all host cores, bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR, BLOCK = 48000, 64
W_FLOPS = {"subtractive": 40.0, "subtractive_seg": 42.0, "additive": 6.0, "fm": 15.0}  # algorithmic ops / voice-sample (Envelope: 7 f64 + cvt in place of EnvAsr's 5 + scale)
N_SM, FP32_LANES = 148, 128


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "MEASURED_PEAKS.json"
    except OSError:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device=0):
        self.rows, self.proc, self.device = [], None, device

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower() == "active":
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def shift_events(ev, frames):
    """Copy of an EVENT_DTYPE array with every absolute time moved `frames` later."""
    T = 282_240_000
    out = ev.copy()
    fr = out["seconds"].astype(np.uint64) * SR + (out["subsec"].astype(np.uint64) * SR) // T + np.uint64(frames)
    out["seconds"] = (fr // SR).astype(np.uint32)
    out["subsec"] = ((fr % SR) * T // SR).astype(np.uint32)
    return out


def build_bank(graph, workload, voices, seconds, rank, world):
    from knaster_b200 import banks

    total = voices * world
    if workload in ("subtractive", "subtractive_seg"):
        seed = 2002 if world == 1 else 4004  # configs[2] / configs[4]
        banks.subtractive_bank(graph, voices, seconds, seed=seed, voice_offset=rank * voices, total_voices=total,
                               envelope="asr" if workload == "subtractive" else "segments")
    elif workload == "additive":
        banks.additive_bank(graph, voices, seconds, voice_offset=rank * voices, total_voices=total)
    elif workload == "fm":
        banks.fm_bank(graph, voices, voice_offset=rank * voices, total_voices=total)
    else:
        raise SystemExit(f"unknown workload {workload}")
    return graph.take_events()


def run_reference(args, rank, world):
    """--impl reference: the oracle (kind "port") on all host cores, bounded sample per step."""
    if rank != 0:
        return
    from concurrent.futures import ThreadPoolExecutor

    from knaster_b200.graph import Graph
    from oracle.oracle import OracleProcessor

    cores = os.cpu_count() or 1
    sample_seconds = min(args.seconds, args.ref_seconds)
    voices = args.voices  # one GPU's share of the bank
    n_blocks = int(round(sample_seconds * SR)) // BLOCK
    per = (voices + cores - 1) // cores
    shards = []
    for c in range(cores):
        nv = min(per, voices - c * per)
        if nv <= 0:
            break
        g = Graph(0, 2, BLOCK, SR)
        from knaster_b200 import banks

        seed = 2002
        if args.workload in ("subtractive", "subtractive_seg"):
            banks.subtractive_bank(g, nv, args.seconds, seed=seed, voice_offset=c * per, total_voices=voices,
                                   envelope="asr" if args.workload == "subtractive" else "segments")
        elif args.workload == "additive":
            banks.additive_bank(g, nv, args.seconds, voice_offset=c * per, total_voices=voices)
        else:
            banks.fm_bank(g, nv, voice_offset=c * per, total_voices=voices)
        shards.append((g, g.take_events(), OracleProcessor(g, ring_buffer_size=1 << 24, fast=True)))

    period = int(round(args.seconds * SR))
    window = n_blocks * BLOCK
    T = 282_240_000

    def one(shard, step):
        # each step renders the NEXT `sample_seconds` of the bank (the oracle's clock runs on);
        # the 10 s event schedule repeats with period `--seconds`, exactly like the GPU arm's steps
        g, ev, proc = shard
        pos = step * window
        e = shift_events(ev, (pos // period) * period)
        fr = e["seconds"].astype(np.uint64) * SR + (e["subsec"].astype(np.uint64) * SR) // T
        e = e[(fr >= pos) & (fr < pos + window)]
        g.pending_event_arrays = [e] if len(e) else []
        out, _ = proc.render(n_blocks)
        return out

    times = []
    with ThreadPoolExecutor(cores) as pool:
        for step in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            outs = list(pool.map(lambda s: one(s, step), shards))
            bus = np.sum(np.stack(outs), axis=0)  # voice-shard replicas summed at the end
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                times.append(dt)
            del bus
    ms = 1e3 * float(np.mean(times))
    value = voices * n_blocks * BLOCK / (ms / 1e3)
    sample = f"{voices} voices x {sample_seconds:g} s of audio per step ({voices * n_blocks * BLOCK:.3g} voice-samples), voice-sharded over {len(shards)} threads"
    line = {
        "impl": "reference", "metric": "voice-samples/sec (f32, 48 kHz)", "value": value, "unit": "voice-samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": "voice-samples/s", "cores": len(shards), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "voice-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "C++ restatement of knaster's CPU render path (oracle/); knaster itself is Rust and cannot be built here. "
                "knaster renders on ONE audio thread; the all-core figure is an upper bound it does not offer.",
    }
    emit_json(line)


def workload_config(args, world):
    names = {"subtractive": "subtractive polysynth: saw -> SvfFilter lowpass -> EnvAsr -> VCA, sample-accurate note events",
             "subtractive_seg": "subtractive polysynth, Envelope variant: saw -> SvfFilter lowpass -> Envelope (A/D/R segments) -> VCA, sample-accurate note events",
             "additive": "additive bank: SinWt partials with per-partial amp smoothing",
             "fm": "FM bank: SinNumeric -> SinNumeric audio-rate freq"}
    return {"workload": names[args.workload], "voices_per_gpu": args.voices, "total_voices": args.voices * world,
            "seconds_per_step": args.seconds, "blocks_per_step": int(round(args.seconds * SR)) // BLOCK, "block_size": BLOCK,
            "sample_rate": SR, "notes_per_voice_per_step": 8 if args.workload.startswith("subtractive") else 0,
            "l2": "per-voice state + events stream once per launch; working set changes every launch (no L2 reuse to flush)",
            "reduce": "none (1 GPU)" if world == 1 else "see config.bus"}


def cpu_baseline(args):
    """Oracle, single thread (the faithful figure: knaster renders on one audio thread), bounded sample."""
    from knaster_b200 import banks
    from knaster_b200.graph import Graph
    from oracle.oracle import OracleProcessor

    voices = min(args.voices, args.cpu_voices)
    secs = min(args.seconds, args.cpu_seconds)
    g = Graph(0, 2, BLOCK, SR)
    if args.workload.startswith("subtractive"):
        banks.subtractive_bank(g, voices, secs, total_voices=args.voices, envelope="asr" if args.workload == "subtractive" else "segments")
    elif args.workload == "additive":
        banks.additive_bank(g, voices, secs, total_voices=args.voices)
    else:
        banks.fm_bank(g, voices, total_voices=args.voices)
    p = OracleProcessor(g, ring_buffer_size=1 << 24, fast=True)
    n_blocks = int(round(secs * SR)) // BLOCK
    t0 = time.perf_counter()
    p.render(n_blocks)
    dt = time.perf_counter() - t0
    return {"value": voices * n_blocks * BLOCK / dt, "unit": "voice-samples/s", "cores": 1, "kind": "port",
            "sample": f"first {voices} voices x first {secs:g} s of the same bank ({dt:.1f} s of CPU); C++ restatement of knaster's "
                      f"one-audio-thread render path (oracle/), host has {os.cpu_count()} cores"}


_JSON_FD = None


def emit_json(line) -> None:
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version line
    to stdout when NCCL_DEBUG asks for it), so main() points fd 1 at stderr for the duration of the run
    and the result goes to the saved original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="subtractive", choices=["subtractive", "subtractive_seg", "additive", "fm"])
    ap.add_argument("--voices", type=int, default=16384, help="voices per GPU")
    ap.add_argument("--seconds", type=float, default=10.0, help="audio seconds per step")
    ap.add_argument("--chunks", type=int, default=10, help="NCCL reduce chunks per step (N>1, --bus nccl)")
    ap.add_argument("--bus", default="peer", choices=["peer", "nccl"],
                    help="N>1: sum the mix bus over peer memory inside the engine's own kernels (falls back to nccl if torch "
                         "symmetric memory is unavailable) or with an NCCL reduce of the rank-local buses")
    ap.add_argument("--blocks-per-launch", type=int, default=0)
    ap.add_argument("--force-interpreter", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-voices", type=int, default=4096)
    ap.add_argument("--cpu-seconds", type=float, default=4.0)
    ap.add_argument("--ref-seconds", type=float, default=1.0, help="--impl reference: audio seconds per step")
    args = ap.parse_args()
    sys.stdout.flush()
    _JSON_FD = os.dup(1)   # see emit_json
    os.dup2(2, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from knaster_b200.multi_gpu import reduce_bus
    from knaster_b200.processor import AudioProcessor, AudioProcessorOptions

    n_blocks = int(round(args.seconds * SR)) // BLOCK
    step_frames = n_blocks * BLOCK
    graph, proc = AudioProcessor.new(0, 2, AudioProcessorOptions(device=local_rank, force_interpreter=args.force_interpreter))
    t0 = time.perf_counter()
    ev0 = build_bank(graph, args.workload, args.voices, args.seconds, rank, world)
    build_s = time.perf_counter() - t0
    info = proc.info()
    if args.blocks_per_launch:
        proc.set_blocks_per_launch(args.blocks_per_launch)
    # one process per GPU shares the box's host cores: each plan's event-pipeline workers get their share
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    host_threads = max(1, min(16, (os.cpu_count() or 1) // max(1, local_world) - 1))
    proc.set_host_threads(host_threads)
    step_no = [0]

    def push_step_events():
        graph.pending_event_arrays = [shift_events(ev0, step_no[0] * step_frames)] if len(ev0) else []
        step_no[0] += 1

    # a non-default torch stream: its cudaStream_t is handed to the C ABI, so the engine's kernels,
    # the NCCL reduce and the torch.cuda.Event timers all live on the same stream
    torch.cuda.set_stream(torch.cuda.Stream())
    bus = torch.empty((n_blocks, 2, BLOCK), dtype=torch.float32, device="cuda")
    chunks = max(1, min(args.chunks, n_blocks)) if world > 1 else 1
    bounds = [round(i * n_blocks / chunks) for i in range(chunks + 1)]

    bus_mode, peer_bus = ("none (1 GPU)" if world == 1 else "nccl"), None
    if world > 1 and args.bus == "peer":
        try:
            from knaster_b200.multi_gpu import PeerBus

            peer_bus = PeerBus(proc, n_blocks)
            ok = torch.ones(1, device="cuda")
        except Exception as e:  # no symmetric memory / no peer access on this box
            sys.stderr.write(f"[bench] peer bus unavailable ({type(e).__name__}: {e}); using the NCCL reduce\n")
            ok = torch.zeros(1, device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)  # all ranks or none
        if ok.item() < 1:
            if peer_bus is not None:
                peer_bus.close()
            peer_bus = None
        else:
            bus_mode = "peer"

    def device_step():
        """kernels + mix-bus sum onto rank 0 (peer-memory stores inside the engine, or an NCCL reduce of
        the rank-local buses) with inputs already in HBM"""
        cur = torch.cuda.current_stream()
        proc.render_device(n_blocks, bus.data_ptr(), cur.cuda_stream)
        if peer_bus is None:
            reduce_bus(bus, dst=0, chunks=chunks)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident inputs.  Each step's events are compiled + uploaded (prepare)
    # outside the timed region, then K steps of kernels are timed back to back.
    launches = 0
    for _ in range(args.warmup):
        push_step_events()
        proc.prepare(n_blocks)
        device_step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    step_ms, kern_ms, red_ms, kern_launches = [], 0.0, 0.0, 0
    for _ in range(args.steps):
        push_step_events()
        proc.prepare(n_blocks)               # host work + H2D, outside the timed region
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        device_step()
        e1.record()
        barrier()
        step_ms.append(e0.elapsed_time(e1))
        k, n = proc.last_kernel_ms(0)
        r, n2 = proc.last_kernel_ms(1)
        kern_ms += k; red_ms += r; kern_launches += n; launches += n + n2
    clocks = sampler.stop()
    t_local = torch.tensor([sum(step_ms)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_local, op=dist.ReduceOp.MAX)
    total_ms = float(t_local.item())
    ms_per_step = total_ms / args.steps
    vs_per_step = args.voices * world * step_frames
    value = vs_per_step / (ms_per_step / 1e3)

    # ---- e2e: through the C ABI with host buffers (rank-local; N>1: plus the reduce)
    host_out = np.empty((n_blocks, 2, BLOCK), dtype=np.float32)
    e2e_times, h2d = [], 0
    for i in range(1 + min(args.steps, 3)):
        push_step_events()
        barrier()
        t0 = time.perf_counter()
        if world == 1:
            proc.render(n_blocks, host_out)
        else:
            device_step()
            torch.cuda.synchronize()
            if rank == 0:
                host_out[:] = bus.cpu().numpy()
        dt = time.perf_counter() - t0
        h2d = proc.last_upload_bytes()
        if i > 0:
            e2e_times.append(dt)
    t_e2e = torch.tensor([float(np.mean(e2e_times))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = vs_per_step / float(t_e2e.item())

    if rank == 0:
        peaks, peak_src = measured_peaks()
        sm_max = float(peaks.get("sm_max_mhz", 1965.0))
        W = W_FLOPS[args.workload]
        peak_fp32 = N_SM * FP32_LANES * sm_max * 1e6 / 1e9  # G un-fused FP32 instr/s
        # dominant kernel: the voice-bank render kernel; per-launch average from CUDA events around every launch
        avg_launch_ms = kern_ms / max(1, kern_launches)
        vs_per_launch = args.voices * step_frames * args.steps / max(1, kern_launches)
        achieved = vs_per_launch * W / (avg_launch_ms / 1e3) / 1e9
        state_bytes = info["state_bytes"]
        frames_per_launch = step_frames * args.steps / max(1, kern_launches)
        rows = (args.voices + 31) // 32
        alg_bytes = 2 * state_bytes + (h2d / max(1, kern_launches / args.steps)) + rows * frames_per_launch * 4
        hbm_gbs = alg_bytes / (avg_launch_ms / 1e3) / 1e9
        traffic = None
        try:  # DRAM bytes of the dominant kernel from the committed ncu capture, scaled to this run's launch size
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f)[info["kernels"][0]]
            if tj["voices"] == args.voices:
                traffic = tj["traffic_bytes_per_launch"] * frames_per_launch / tj["frames_per_launch"]
        except (OSError, KeyError, ValueError):
            pass
        line = {
            "metric": "voice-samples/sec (f32, 48 kHz)", "value": value, "unit": "voice-samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args, world), bus=(
                "peer memory: every rank's reduce_bus kernel stores its bus into rank 0 over NVLink, rank 0 folds the slots per launch"
                if bus_mode == "peer" else f"NCCL reduce(sum) of the rank-local stereo bus to rank 0, {chunks} chunks per step"
                if bus_mode == "nccl" else bus_mode)),
            "kernels": info["kernels"],
            "roofline": {
                "bound": "fp32", "achieved": achieved, "peak": peak_fp32, "unit": "GFLOP/s (un-fused FP32 instr)",
                "frac": achieved / peak_fp32, "traffic": traffic,
                "kernel": info["kernels"][0], "avg_launch_ms": avg_launch_ms, "launches": kern_launches,
                "flops_per_voice_sample": W,
                "peak_source": f"{N_SM} SMs x {FP32_LANES} FP32 lanes x sm_max_mhz {sm_max:g} ({peak_src}); no tensor/HBM bound: "
                               "nothing here is a dense contraction and intermediates never touch HBM",
                "hbm": {"achieved": hbm_gbs, "peak": float(peaks.get("hbm_gbs", 6650.0)), "unit": "GB/s",
                        "frac": hbm_gbs / float(peaks.get("hbm_gbs", 6650.0)), "algorithmic_bytes_per_launch": alg_bytes},
                "kernel_share_of_step": kern_ms / max(1e-9, sum(step_ms)), "reduce_bus_ms_per_step": red_ms / args.steps,
            },
            "e2e": {"value": e2e_value, "unit": "voice-samples/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(host_out.nbytes), "ms_per_step": 1e3 * float(t_e2e.item())},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "plan": {k: info[k] for k in ("n_groups", "n_voices", "n_mix_nodes", "n_fused_groups", "state_bytes")},
            "build_graph_s": build_s,
            "host": {"cores": os.cpu_count(), "event_pipeline_threads_per_gpu": host_threads},
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(args)
        emit_json(line)
    if peer_bus is not None and peer_bus.timed_out():
        sys.stderr.write("[bench] WARNING: a rank timed out waiting for peer-bus data\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
