// NOT COMPILED IN THIS REPOSITORY'S IMAGE (no cargo/rustc).  The text of INTEGRATION.md as files:
// the binding a knaster maintainer would add next to knaster_graph.  tests/test_host_plan.py checks that
// src/ffi.rs declares every entry point of include/knaster_gpu.h.
//! knaster_gpu: `GpuProcessor`, a drop-in for `knaster_graph::processor::AudioProcessor`'s non-realtime
//! surface (`run_without_inputs`, `output_block`, plus a batched `render`) backed by libknaster_gpu.so.
//! See INTEGRATION.md for the one hook it needs inside knaster (`UGen::gpu_desc`) and for the event path.
pub mod ffi;

use std::ffi::CStr;
use std::os::raw::c_int;
use std::sync::{Arc, Mutex};

use knaster_core::{ParameterSmoothing, ParameterValue, Rate};
use knaster_graph::graph::{Graph, GraphError, NodeKey};
use knaster_graph::scheduling::{SchedulingEvent, Time};
use slotmap::SecondaryMap;

/// source index of an edge: a node, or the graph's own inputs (-2 = KGPU_GRAPH in include/knaster_gpu.h)
fn src(index: &SecondaryMap<NodeKey, i32>, source: knaster_graph::graph::NodeOrGraph) -> i32 {
    match source {
        knaster_graph::graph::NodeOrGraph::Node(k) => index[k],
        knaster_graph::graph::NodeOrGraph::Graph => -2,
    }
}

/// The event path (knaster_graph/src/handle.rs:38-73): `SchedulingChannelSender::send` pushes a `SchedulingEvent` into an
/// `rtrb` ring that GraphGen drains at the top of every block (graph_gen.rs:143-166).  `GpuEventSink` is the second sink
/// INTEGRATION.md asks `SchedulingChannelSender` to carry: `send` converts the event to a `kgpu_event` (scheduling.rs:29-36
/// field by field) and appends it to a vector that `GpuProcessor::run* / render` hands to `kgpu_plan_push_events` before
/// rendering -- arrival order is kept, nothing is dropped for a full ring.  Any number of control threads
/// may hold a clone, exactly like `Arc<Mutex<rtrb::Producer>>` today.
#[derive(Clone)]
pub struct GpuEventSink { queue: Arc<Mutex<Vec<ffi::kgpu_event>>>, index: Arc<SecondaryMap<NodeKey, i32>> }

impl GpuEventSink {
    /// `SchedulingChannelSender::send` for the GPU engine.
    pub fn send(&self, ev: SchedulingEvent) -> Result<(), GraphError> {
        if ev.token.is_some() { return Err(GraphError::UnsupportedOnGpu); }   // SchedulingToken::activate is todo!() (scheduling.rs:175-178)
        let node = *self.index.get(ev.node_key).ok_or(GraphError::NodeNotFound)? as u32;
        let (value_kind, value) = match ev.value {                            // parameters/types.rs:25-37
            None => (0, 0.0),
            Some(ParameterValue::Float(v)) => (1, v),
            Some(ParameterValue::Trigger) => (2, 0.0),
            Some(ParameterValue::Integer(i)) => (3, i.0 as f64),
            Some(ParameterValue::Bool(b)) => (4, if b { 1.0 } else { 0.0 }),
            Some(ParameterValue::Smoothing(..)) => return Err(GraphError::UnsupportedOnGpu), // travels in `smoothing`
        };
        let (smoothing_kind, smooth_seconds) = match ev.smoothing {           // parameters/types.rs:108-119
            None => (0, 0.0),
            Some(ParameterSmoothing::None) => (1, 0.0),
            Some(ParameterSmoothing::Linear(s)) => (2, s),
        };
        let (time_kind, seconds, subsec) = match ev.time {                    // scheduling.rs:73-92, time.rs:25-28
            None => (0, 0, 0),
            Some(Time::At(s)) => (1, s.seconds(), s.subsecond_tesimals()),
            Some(Time::After(s)) => (2, s.seconds(), s.subsecond_tesimals()),
        };
        let e = ffi::kgpu_event { node, param: ev.parameter as u32, value_kind, smoothing_kind, value, smooth_seconds,
                                  smooth_rate: Rate::BlockRate as u32, time_kind, seconds, subsec, _pad: 0 };
        self.queue.lock().unwrap_or_else(|p| p.into_inner()).push(e);
        Ok(())
    }
}

pub struct GpuProcessor { plan: *mut ffi::kgpu_plan, block: usize, outputs: usize,
                          queue: Arc<Mutex<Vec<ffi::kgpu_event>>>, index: Arc<SecondaryMap<NodeKey, i32>> }

impl GpuProcessor {
    /// Call after `graph.edit(..)` / `commit_changes()`; the graph must stay static afterwards.
    pub fn new(graph: &Graph<f32>, device: i32) -> Result<Self, GraphError> {
        let order = graph.node_keys();                       // any order: indices are remapped below
        let index: SecondaryMap<NodeKey, i32> = order.iter().enumerate().map(|(i, k)| (*k, i as i32)).collect();
        let mut keep = Vec::new();                           // wrapper/segment arrays must outlive the call
        let nodes: Vec<ffi::kgpu_node_desc> = order.iter().map(|k| graph.node(*k).ugen_gpu_desc(&mut keep)).collect::<Option<_>>()
            .ok_or(GraphError::UnsupportedOnGpu)?;
        let mut edges = Vec::new();
        for (sink, ins) in graph.node_input_edges() {        // graph.rs:176
            for (ch, e) in ins.iter().enumerate() { if let Some(e) = e {
                edges.push(ffi::kgpu_edge { source_node: src(&index, e.source), source_channel: e.channel_in_source as u32,
                                            sink_node: index[sink], sink_channel: ch as u32 });
            }}
        }
        for (ch, e) in graph.output_edges().iter().enumerate() { if let Some(e) = e {   // graph.rs:182
            edges.push(ffi::kgpu_edge { source_node: src(&index, e.source), source_channel: e.channel_in_source as u32,
                                        sink_node: -2 /* KGPU_GRAPH */, sink_channel: ch as u32 });
        }}
        let pedges: Vec<_> = graph.node_parameter_edges().flat_map(|(sink, es)| es.iter().map(move |pe|   // graph.rs:179
            ffi::kgpu_param_edge { source_node: index[pe.source], source_channel: pe.channel_in_source as u32,
                                   sink_node: index[sink], param_index: pe.parameter_index as u32 })).collect();
        let desc = ffi::kgpu_graph_desc { abi_version: 1, sample_rate: graph.sample_rate(), block_size: graph.block_size() as u32,
            n_inputs: graph.inputs() as u32, n_outputs: graph.outputs() as u32, device,
            n_nodes: nodes.len() as u32, n_edges: edges.len() as u32, n_param_edges: pedges.len() as u32, flags: 0,
            nodes: nodes.as_ptr(), edges: edges.as_ptr(), param_edges: pedges.as_ptr() };
        let mut plan = std::ptr::null_mut();
        check(unsafe { ffi::kgpu_plan_create(&desc, &mut plan) })?;
        Ok(Self { plan, block: graph.block_size(), outputs: graph.outputs() as usize,
                  queue: Arc::new(Mutex::new(Vec::new())), index: Arc::new(index) })
    }
    /// The sender control threads use instead of (or beside) the rtrb producer; see `GpuEventSink`.
    pub fn event_sink(&self) -> GpuEventSink { GpuEventSink { queue: self.queue.clone(), index: self.index.clone() } }
    /// What GraphGen does at the top of a block (graph_gen.rs:143-166): everything sent so far reaches the engine, in send order.
    fn drain_events(&mut self) -> Result<(), GraphError> {
        let batch = std::mem::take(&mut *self.queue.lock().unwrap_or_else(|p| p.into_inner()));
        if batch.is_empty() { return Ok(()); }
        check(unsafe { ffi::kgpu_plan_push_events(self.plan, batch.as_ptr(), batch.len()) })
    }
    /// AudioProcessor::run_without_inputs (processor.rs:142-148)
    pub fn run_without_inputs(&mut self) -> Result<(), GraphError> {
        self.drain_events()?;
        check(unsafe { ffi::kgpu_render_block(self.plan) })
    }
    /// AudioProcessor::run (processor.rs:119-141): one block with one slice per graph input
    pub fn run(&mut self, inputs: &[&[f32]]) -> Result<(), GraphError> {
        self.drain_events()?;
        let flat: Vec<f32> = inputs.iter().flat_map(|c| c[..self.block].iter().copied()).collect();   // [inputs][block]
        check(unsafe { ffi::kgpu_render_inputs(self.plan, 1, flat.as_ptr(), std::ptr::null_mut()) })
    }
    /// AudioProcessor::output_block (processor.rs:182-184): [outputs][block] f32
    pub fn output_block(&mut self) -> &[f32] { unsafe { std::slice::from_raw_parts(ffi::kgpu_output_block(self.plan), self.block * self.outputs) } }
    /// Non-realtime render of n blocks: [n_blocks][outputs][block]
    pub fn render(&mut self, n_blocks: u64, out: &mut [f32]) -> Result<(), GraphError> {
        assert_eq!(out.len() as u64, n_blocks * (self.block * self.outputs) as u64);
        self.drain_events()?;
        check(unsafe { ffi::kgpu_render(self.plan, n_blocks, out.as_mut_ptr()) })
    }
}
impl Drop for GpuProcessor { fn drop(&mut self) { unsafe { ffi::kgpu_plan_destroy(self.plan) } } }
fn check(rc: c_int) -> Result<(), GraphError> {
    if rc == 0 { Ok(()) } else { Err(GraphError::Gpu(rc, unsafe { CStr::from_ptr(ffi::kgpu_last_error()) }.to_string_lossy().into())) }
}
