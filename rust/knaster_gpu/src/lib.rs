// NOT COMPILED IN THIS REPOSITORY'S IMAGE (no cargo/rustc).  The text of INTEGRATION.md as files:
// the binding a knaster maintainer would add next to knaster_graph.  tests/test_host_plan.py checks that
// src/ffi.rs declares every entry point of include/knaster_gpu.h.
//! knaster_gpu: `GpuProcessor`, a drop-in for `knaster_graph::processor::AudioProcessor`'s non-realtime
//! surface (`run_without_inputs`, `output_block`, plus a batched `render`) backed by libknaster_gpu.so.
//! See INTEGRATION.md for the one hook it needs inside knaster (`UGen::gpu_desc`) and for the event path.
pub mod ffi;

use std::ffi::CStr;
use std::os::raw::c_int;

use knaster_graph::graph::{Graph, GraphError, NodeKey};
use slotmap::SecondaryMap;

/// source index of an edge: a node, or the graph's own inputs (-2 = KGPU_GRAPH in include/knaster_gpu.h)
fn src(index: &SecondaryMap<NodeKey, i32>, source: knaster_graph::graph::NodeOrGraph) -> i32 {
    match source {
        knaster_graph::graph::NodeOrGraph::Node(k) => index[k],
        knaster_graph::graph::NodeOrGraph::Graph => -2,
    }
}

pub struct GpuProcessor { plan: *mut ffi::kgpu_plan, block: usize, outputs: usize }

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
        Ok(Self { plan, block: graph.block_size(), outputs: graph.outputs() as usize })
    }
    /// AudioProcessor::run_without_inputs (processor.rs:142-148)
    pub fn run_without_inputs(&mut self) -> Result<(), GraphError> { check(unsafe { ffi::kgpu_render_block(self.plan) }) }
    /// AudioProcessor::output_block (processor.rs:182-184): [outputs][block] f32
    pub fn output_block(&mut self) -> &[f32] { unsafe { std::slice::from_raw_parts(ffi::kgpu_output_block(self.plan), self.block * self.outputs) } }
    /// Non-realtime render of n blocks: [n_blocks][outputs][block]
    pub fn render(&mut self, n_blocks: u64, out: &mut [f32]) -> Result<(), GraphError> {
        assert_eq!(out.len() as u64, n_blocks * (self.block * self.outputs) as u64);
        check(unsafe { ffi::kgpu_render(self.plan, n_blocks, out.as_mut_ptr()) })
    }
}
impl Drop for GpuProcessor { fn drop(&mut self) { unsafe { ffi::kgpu_plan_destroy(self.plan) } } }
fn check(rc: c_int) -> Result<(), GraphError> {
    if rc == 0 { Ok(()) } else { Err(GraphError::Gpu(rc, unsafe { CStr::from_ptr(ffi::kgpu_last_error()) }.to_string_lossy().into())) }
}
