#!/usr/bin/env python3
"""Write the manifest `knaster_ref_dump` (real knaster, Rust) renders: the per-voice constructor arguments
and the parameter events of one bench configuration, taken from the SAME graph builder (knaster_b200/
banks.py) the GPU engine and the oracle use.

    python tools/knaster_ref_dump/export_manifest.py --config subtractive_asr --voices 64 --seconds 1 --out m.txt
    cargo run --release --manifest-path tools/knaster_ref_dump/Cargo.toml -- m.txt ref.f32      # on a Rust box
    python tools/knaster_ref_dump/compare_ref_dump.py m.txt ref.f32                              # back here
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
SR, BLOCK, T = 48000, 64, 282_240_000

from knaster_b200 import banks, ugens as U  # noqa: E402
from knaster_b200.graph import Graph  # noqa: E402


def build(config, voices, seconds, graph=None):
    g = graph if graph is not None else Graph(0, 2, BLOCK, SR)
    if config == "subtractive_asr":
        banks.subtractive_bank(g, voices, seconds)
    elif config == "subtractive_seg":
        banks.subtractive_bank(g, voices, seconds, envelope="segments")
    elif config == "additive":
        banks.additive_bank(g, voices, seconds)
    elif config == "fm":
        banks.fm_bank(g, voices)
    elif config == "readme_sine":
        banks.readme_sine(g)
    else:
        raise SystemExit(f"unknown config {config}")
    return g, g.take_events()


def wr_value(ugen, kind):
    return next(w.value for w in ugen.wrappers if w.kind == kind)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True, choices=["subtractive_asr", "subtractive_seg", "additive", "fm", "readme_sine"])
    ap.add_argument("--voices", type=int, default=64)
    ap.add_argument("--seconds", type=float, default=1.0)
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    g, ev = build(a.config, a.voices, a.seconds)
    n_blocks = int(round(a.seconds * SR)) // BLOCK
    user = [(i, n.ugen) for i, n in enumerate(g.nodes) if not n.auto_math_node]   # nodes the builder pushed, in order
    lines = ["knaster_ref_dump_manifest 1", f"config {a.config}", f"sr {SR} block {BLOCK} blocks {n_blocks} voices {a.voices}"]
    owner = {}  # node id -> (voice, target)
    names = {}  # node id -> parameter names
    def own(node_id, ugen, voice, target):
        owner[node_id] = (voice, target)
        names[node_id] = ugen.param_descriptions()
    if a.config in ("subtractive_asr", "subtractive_seg"):
        per = [(i, u) for i, u in user if u.kind in (U.KIND_POLYBLEP, U.KIND_SVF, U.KIND_ENV_ASR, U.KIND_ENVELOPE)]
        for v in range(a.voices):
            (si, saw), (fi, svf), (ei, env) = per[3 * v: 3 * v + 3]
            own(si, saw, v, "saw"); own(fi, svf, v, "svf"); own(ei, env, v, "env")
            gain = wr_value(env, U.WR_MUL)
            if a.config == "subtractive_asr":
                vals = [saw.args[0], svf.args[0], svf.args[1], env.args[0], env.args[1], gain]
            else:
                (att, _), (dec, sus), (rel, _) = env.segments
                vals = [saw.args[0], svf.args[0], svf.args[1], att, dec, sus, rel, gain]
            lines.append(f"voice {v} " + " ".join(repr(float(x)) for x in vals))
    elif a.config == "additive":
        per = [(i, u) for i, u in user if u.kind == U.KIND_SIN_WT]
        for v, (i, osc) in enumerate(per):
            own(i, osc, v, "osc")
            lines.append(f"voice {v} {float(osc.args[0])!r} {float(wr_value(osc, U.WR_MUL))!r}")
    elif a.config == "fm":
        sins = [(i, u) for i, u in user if u.kind == U.KIND_SIN_NUMERIC]
        consts = [u for i, u in user if u.kind == U.KIND_CONSTANT]
        for v in range(a.voices):
            mod, car = sins[2 * v][1], sins[2 * v + 1][1]
            idx, fc, amp = (c.args[0] for c in consts[3 * v: 3 * v + 3])
            assert fc == car.args[0]
            lines.append(f"voice {v} {float(mod.args[0])!r} {float(car.args[0])!r} {float(idx)!r} {float(amp)!r}")
    else:
        lines.append("voice 0")
    frames = ev["seconds"].astype(np.uint64) * SR + (ev["subsec"].astype(np.uint64) * SR) // T
    for e, fr in zip(ev, frames):
        voice, target = owner[int(e["node"])]
        name = names[int(e["node"])][int(e["param"])]
        if e["smoothing_kind"] == 2:
            kind, val = "s", float(e["smooth_seconds"])
        else:
            kind, val = {1: "f", 2: "t", 3: "i"}[int(e["value_kind"])], float(e["value"])
        lines.append(f"event {voice} {target} {name} {kind} {val!r} {int(fr)}")
    with open(a.out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print(f"{a.out}: {a.voices} voices, {len(ev)} events, {n_blocks} blocks")


if __name__ == "__main__":
    main()
