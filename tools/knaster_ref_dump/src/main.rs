//! knaster_ref_dump: render a manifest written by `export_manifest.py` with real knaster and write the
//! stereo bus as raw little-endian f32, `[block][channel][frame]` (knaster's `RawContiguousBlock`
//! layout per block) -- the layout `kgpu_render` and the oracle use, so `compare_ref_dump.py` can diff
//! the three directly.
//!
//! usage: knaster_ref_dump <manifest.txt> <out.f32> [--solo <voice>]
//!
//! Not compiled in the repository's image (no Rust toolchain there); written against the reference's
//! public API as used by its own examples and tests:
//!   AudioProcessor::<f32>::new::<U0, U2>(AudioProcessorOptions { .. })   knaster_graph/src/processor.rs:69
//!   graph.edit(|g| { g.push(..); (a >> b) * c; .out([0, 0]).to_graph_out(); handle.param("name") })
//!                                                  knaster_graph/examples/scheduling_test.rs:40-46
//!   Parameter::{set_at, trig_at, smooth_at}        knaster_graph/src/graph_edit.rs:1738,1862,1803
//!   UGenWrapperCoreExt::{wr_mul, smooth_params, ar_params, precise_timing::<N>}
//!                                                  knaster_core_dsp/src/wrappers_core.rs:26-56
//! Two properties of the reference shape the driver loop below: scheduled events
//! that wait longer than ~1 s on the audio side are dropped (graph_gen.rs:123) and the event ring holds
//! `ring_buffer_size` entries -- so events are fed at most half a second ahead of the render clock and
//! the ring is sized for the busiest half second.
use std::fs::File;
use std::io::{BufRead, BufReader, BufWriter, Write};

use knaster::envelopes::{EnvAsr, Envelope, EnvelopeSegment};
use knaster::osc::{SinNumeric, SinWt};
use knaster::polyblep::{PolyBlep, Waveform};
use knaster::svf::{SvfFilter, SvfFilterType};
use knaster::typenum::{U0, U2};
use knaster::wrappers_core::UGenWrapperCoreExt;
use knaster_graph::graph_edit::Parameter;
use knaster_graph::processor::{AudioProcessor, AudioProcessorOptions};
use knaster_graph::{Block, ParameterSmoothing, Seconds};

struct Event {
    voice: usize,
    target: String, // saw | svf | env | osc
    param: String,
    kind: char, // f float, t trigger, i integer, s smoothing (Linear seconds)
    value: f64,
    frame: u64,
}

struct Manifest {
    config: String,
    sr: u32,
    block: usize,
    blocks: usize,
    voices: Vec<Vec<f64>>,
    events: Vec<Event>,
}

fn read_manifest(path: &str) -> Manifest {
    let mut m = Manifest { config: String::new(), sr: 48000, block: 64, blocks: 0, voices: vec![], events: vec![] };
    for line in BufReader::new(File::open(path).expect("manifest")).lines() {
        let line = line.unwrap();
        let t: Vec<&str> = line.split_whitespace().collect();
        match t.first().copied() {
            Some("config") => m.config = t[1].to_string(),
            Some("sr") => {
                m.sr = t[1].parse().unwrap();
                m.block = t[3].parse().unwrap();
                m.blocks = t[5].parse().unwrap();
            }
            Some("voice") => m.voices.push(t[2..].iter().map(|x| x.parse().unwrap()).collect()),
            Some("event") => m.events.push(Event {
                voice: t[1].parse().unwrap(),
                target: t[2].to_string(),
                param: t[3].to_string(),
                kind: t[4].chars().next().unwrap(),
                value: t[5].parse().unwrap(),
                frame: t[6].parse().unwrap(),
            }),
            _ => {}
        }
    }
    m.events.sort_by_key(|e| e.frame); // stable: arrival order inside a frame is the manifest's
    m
}

/// One `Parameter` per (voice, target, name): built inside `graph.edit`, used outside it.
struct VoiceParams {
    named: Vec<(String, String, Parameter)>,
}
impl VoiceParams {
    fn get(&mut self, target: &str, name: &str) -> &mut Parameter {
        &mut self.named.iter_mut().find(|(t, n, _)| t == target && n == name).expect("unknown parameter").2
    }
}

fn main() {
    let args: Vec<String> = std::env::args().collect();
    let m = read_manifest(&args[1]);
    let solo: Option<usize> = args.iter().position(|a| a == "--solo").map(|i| args[i + 1].parse().unwrap());
    let half_second = (m.sr / 2) as u64;
    let mut busiest = 0usize; // events per half second, for the ring size
    {
        let (mut lo, mut hi) = (0usize, 0usize);
        while hi < m.events.len() {
            while m.events[hi].frame - m.events[lo].frame > half_second { lo += 1; }
            busiest = busiest.max(hi - lo + 1);
            hi += 1;
        }
    }
    let (mut graph, mut proc_, _log) = AudioProcessor::<f32>::new::<U0, U2>(AudioProcessorOptions {
        block_size: m.block,
        sample_rate: m.sr,
        ring_buffer_size: (2 * busiest + 1024).max(4 * m.voices.len()),
        ..Default::default()
    });
    let cfg = m.config.clone();
    let voices = m.voices.clone();
    let mut params: Vec<VoiceParams> = graph.edit(|g| {
        let mut all = Vec::with_capacity(voices.len());
        for (i, v) in voices.iter().enumerate() {
            if solo.is_some_and(|s| s != i) { all.push(VoiceParams { named: vec![] }); continue; }
            let mut named = Vec::new();
            match cfg.as_str() {
                // voice: f amp   -- SinWt(f).wr_mul(amp).smooth_params() -> stereo
                "additive" => {
                    let osc = g.push(SinWt::new(v[0] as f32).wr_mul(v[1] as f32).smooth_params());
                    osc.out([0, 0]).to_graph_out();
                    named.push(("osc".into(), "wr_mul".into(), osc.param("wr_mul")));
                    named.push(("osc".into(), "freq".into(), osc.param("freq")));
                }
                // voice: f0 fc q att rel gain
                "subtractive_asr" => {
                    let saw = g.push(PolyBlep::new(Waveform::Sawtooth, v[0] as f32).precise_timing::<8>());
                    let svf = g.push(SvfFilter::new(SvfFilterType::Low, v[1] as f32, v[2] as f32, 0.0).precise_timing::<8>());
                    let env = g.push(EnvAsr::new(v[3] as f32, v[4] as f32).wr_mul(v[5] as f32).precise_timing::<8>());
                    ((saw >> svf) * env).out([0, 0]).to_graph_out();
                    named.push(("saw".into(), "freq".into(), saw.param("freq")));
                    named.push(("svf".into(), "cutoff_freq".into(), svf.param("cutoff_freq")));
                    named.push(("env".into(), "t_restart".into(), env.param("t_restart")));
                    named.push(("env".into(), "t_release".into(), env.param("t_release")));
                }
                // voice: f0 fc q att decay sustain rel gain
                "subtractive_seg" => {
                    let saw = g.push(PolyBlep::new(Waveform::Sawtooth, v[0] as f32).precise_timing::<8>());
                    let svf = g.push(SvfFilter::new(SvfFilterType::Low, v[1] as f32, v[2] as f32, 0.0).precise_timing::<8>());
                    let segs = vec![EnvelopeSegment::new(v[3], 1.0), EnvelopeSegment::new(v[4], v[5]), EnvelopeSegment::new(v[6], 0.0)];
                    let env = g.push(Envelope::new(0.0, segs).wr_mul(v[7] as f32).precise_timing::<8>());
                    ((saw >> svf) * env).out([0, 0]).to_graph_out();
                    named.push(("saw".into(), "freq".into(), saw.param("freq")));
                    named.push(("svf".into(), "cutoff_freq".into(), svf.param("cutoff_freq")));
                    named.push(("env".into(), "t_restart".into(), env.param("t_restart")));
                    named.push(("env".into(), "t_stop".into(), env.param("t_stop")));
                    named.push(("env".into(), "jump_to_segment".into(), env.param("jump_to_segment")));
                }
                // voice: fm fc idx amp   -- car.link("freq", mod * idx + fc); car * amp
                "fm" => {
                    let modu = g.push(SinNumeric::new(v[0] as f32));
                    let car = g.push(SinNumeric::new(v[1] as f32).ar_params());
                    car.link("freq", modu * (v[2] as f32) + (v[1] as f32));
                    (car * (v[3] as f32)).out([0, 0]).to_graph_out();
                }
                // README.md:35-47
                "readme_sine" => {
                    let sine = g.push(SinWt::new(440.0));
                    (sine * 0.2).out([0, 0]).to_graph_out();
                }
                other => panic!("unknown config {other}"),
            }
            all.push(VoiceParams { named });
        }
        all
    });

    let mut out = BufWriter::new(File::create(&args[2]).expect("output file"));
    let mut next = 0usize;
    for b in 0..m.blocks {
        let clock = (b * m.block) as u64;
        while next < m.events.len() && m.events[next].frame < clock + half_second {
            let e = &m.events[next];
            next += 1;
            if solo.is_some_and(|s| s != e.voice) { continue; }
            let at = Seconds::from_samples(e.frame, m.sr as u64);
            let p = params[e.voice].get(&e.target, &e.param);
            match e.kind {
                'f' => p.set_at(e.value, at).unwrap(),
                'i' => p.set_at(e.value as usize, at).unwrap(),
                't' => p.trig_at(at).unwrap(),
                's' => p.smooth_at(ParameterSmoothing::Linear(e.value as f32), at).unwrap(),
                k => panic!("unknown event kind {k}"),
            }
        }
        proc_.run_without_inputs();
        let block = proc_.output_block();
        for ch in 0..2 {
            for fr in 0..m.block {
                out.write_all(&block.read(ch, fr).to_le_bytes()).unwrap();
            }
        }
    }
    out.flush().unwrap();
    eprintln!("{}: {} voices x {} blocks written to {}", m.config, m.voices.len(), m.blocks, args[2]);
}
