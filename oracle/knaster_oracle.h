/*
 * knaster_oracle.h -- C API of the CPU oracle.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.
 * The product (knaster_b200/, include/knaster_gpu.h) never links or calls it.
 *
 * The oracle is a C++ restatement of knaster's f32 CPU render path
 * (AudioProcessor -> GraphGen -> Task -> UGen/wrappers), following the Rust
 * source line by line; every function in knaster_oracle.cpp cites the
 * reference file:line it follows.  knaster itself is Rust and cannot be built
 * in this image (no cargo/rustc), so there is no oracle/_ref.
 *
 * Parity pinning: the reference's own known-answer tests for this path are
 * reproduced in tests/test_oracle_golden.py (WrPreciseTiming golden vector,
 * wrapper arithmetic, MathUGen arithmetic/channel layout, Seconds<->samples,
 * graph tests).  The DSP bodies of SinWt / SinNumeric / PolyBlep / SvfFilter /
 * OnePole / EnvAsr / Envelope / WrSmoothParams / WrArParams have NO golden
 * vectors in the reference: for those, "parity unpinned" -- parity rests on
 * this restatement's fidelity to the cited source lines.
 */
#ifndef KNASTER_ORACLE_H
#define KNASTER_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ko_graph ko_graph;

/* UGen kinds (numerically identical to kgpu_ugen_kind, defined independently) */
enum {
    KO_SIN_WT = 1,
    KO_SIN_NUMERIC = 2,
    KO_POLYBLEP = 3,
    KO_SVF = 4,
    KO_ONEPOLE_LPF = 5,
    KO_ONEPOLE_HPF = 6,
    KO_ENV_ASR = 7,
    KO_ENV_AR = 8,
    KO_ENVELOPE = 9,
    KO_MATH = 10,
    KO_CONSTANT = 11,
    KO_TEST_NUM = 12,           /* reference test fixture TestNumUGen */
    KO_TEST_IN_PLUS_PARAM = 13, /* reference test fixture TestInPlusParamUGen */
    KO_MATH1 = 14,              /* Math1UGen, mode: 0 Ceil 1 Sqrt 2 Floor 3 Trunc 4 Fract 5 Exp */
    KO_PHASOR = 15,
    KO_WHITE_NOISE = 16, /* noise.rs; args[0] = seed */
    KO_PINK_NOISE = 17,
    KO_BROWN_NOISE = 18,
    KO_RANDOM_LIN = 19,  /* args[0] = freq, args[1] = seed */
    KO_PAN2 = 20,        /* pan.rs; args[0] = pan */
};
/* MathUGen ops */
enum { KO_OP_ADD = 0, KO_OP_SUB = 1, KO_OP_MUL = 2, KO_OP_DIV = 3, KO_OP_POW = 4 };
/* wrapper kinds */
enum {
    KO_WR_MUL = 1,
    KO_WR_ADD = 2,
    KO_WR_SUB = 3,
    KO_WR_VSUB = 4,
    KO_WR_DIV = 5,
    KO_WR_VDIV = 6,
    KO_WR_POWF = 7,
    KO_WR_POWI = 8,
    KO_WR_SMOOTH_PARAMS = 9,
    KO_WR_PRECISE_TIMING = 10,
    KO_WR_AR_PARAMS = 11,
};

typedef struct {
    uint32_t kind;     /* KO_WR_* */
    uint32_t capacity; /* WrPreciseTiming<N>: N */
    double value;      /* WrMul/WrAdd/... value (f32 in knaster), WrPowi exponent */
} ko_wrapper_desc;

typedef struct {
    uint32_t kind;     /* KO_* ugen kind */
    uint32_t mode;     /* PolyBlep waveform / Svf filter type / Math op */
    uint32_t channels; /* MathUGen channel count N (inputs 2N, outputs N) */
    uint32_t flags;    /* bit0: Envelope looping */
    double args[4];    /* ctor args, see knaster_oracle.cpp make_ugen() */
    uint32_t n_wrappers;
    uint32_t n_segments;
    const ko_wrapper_desc *wrappers; /* innermost first */
    const double *segments;          /* Envelope: (duration, value) pairs */
} ko_node_desc;

typedef struct {
    uint32_t node;
    uint32_t param;
    uint32_t value_kind;     /* 0 none, 1 float, 2 trigger, 3 integer, 4 bool */
    uint32_t smoothing_kind; /* 0 no smoothing field, 1 ParameterSmoothing::None, 2 Linear */
    double value;
    float smooth_seconds;
    uint32_t smooth_rate;    /* 0 BlockRate, 1 AudioRate */
    uint32_t time_kind;      /* 0 None, 1 Time::at (absolute), 2 Time::after (relative) */
    uint32_t seconds;        /* Seconds.seconds */
    uint32_t subsec;         /* Seconds.subsecond_tesimals */
    uint32_t _pad;
} ko_event;

ko_graph *ko_graph_create(uint32_t sample_rate, uint32_t block_size, uint32_t n_inputs,
                          uint32_t n_outputs, uint32_t ring_buffer_size);
void ko_graph_destroy(ko_graph *g);
const char *ko_last_error(void);

/* push a node (Graph::push_internal + Node::init), returns node index or <0 */
int ko_push(ko_graph *g, const ko_node_desc *desc);
/* query channel/parameter counts of a pushed node */
int ko_node_inputs(ko_graph *g, int node);
int ko_node_outputs(ko_graph *g, int node);
int ko_node_parameters(ko_graph *g, int node);
/* lowered edges (the front end resolves additive connects into Add nodes).
 * source_node: >=0 node, -2 graph input, -1 clears the edge. */
int ko_set_input_edge(ko_graph *g, int sink_node, uint32_t sink_channel, int source_node,
                      uint32_t source_channel);
int ko_set_output_edge(ko_graph *g, uint32_t out_channel, int source_node,
                       uint32_t source_channel);
int ko_set_param_edge(ko_graph *g, int sink_node, uint32_t param_index, int source_node,
                      uint32_t source_channel);
/* Graph::commit_changes: node order, buffers, tasks, AR parameter buffers */
int ko_commit(ko_graph *g);

/* faithful path: push one event into the scheduling ring NOW */
int ko_send_event(ko_graph *g, const ko_event *ev);
/* AudioProcessor::run / run_without_inputs; inputs = n_inputs pointers of block_size */
int ko_run_block(ko_graph *g, const float *const *inputs);
/* AudioProcessor::output_block: [n_outputs][block_size] */
const float *ko_output_block(ko_graph *g);
uint64_t ko_frame_clock(ko_graph *g);
/* number of rt_log! warnings raised so far (dropped events, unreachable delays...) */
uint64_t ko_log_count(ko_graph *g);

/* record the output channel of a node every block ("tap"), for pre-mix parity */
int ko_add_tap(ko_graph *g, int node, uint32_t channel);
/* render n_blocks; events (any order) are fed just in time: an event enters
 * the ring right before the block that contains its absolute due frame (or
 * before the next block if it is already late / has no time), preserving array
 * order inside a block.  out: [n_blocks][n_outputs][block] or NULL.
 * taps_out: [n_taps][n_blocks*block] or NULL. */
int ko_render(ko_graph *g, uint64_t n_blocks, float *out, const ko_event *events,
              size_t n_events, float *taps_out);

/* knaster_primitives/src/time.rs helpers, exported for the known-answer tests */
void ko_seconds_from_samples(uint64_t samples, uint64_t sample_rate, uint32_t *seconds,
                             uint32_t *subsec);
uint64_t ko_seconds_to_samples(uint32_t seconds, uint32_t subsec, uint64_t sample_rate);
void ko_seconds_from_secs_f64(double s, uint32_t *seconds, uint32_t *subsec);
uint64_t ko_seconds_to_tesimals(uint32_t seconds, uint32_t subsec);
void ko_seconds_from_tesimals(uint64_t t, uint32_t *seconds, uint32_t *subsec);
void ko_seconds_add(uint32_t s0, uint32_t t0, uint32_t s1, uint32_t t1, uint32_t *seconds,
                    uint32_t *subsec);

/* standalone wrapper harness used by the reference's wrappers_core.rs tests:
 * build a ugen (no graph), set delays / params, process one frame or block */
typedef struct ko_ugen ko_ugen;
ko_ugen *ko_ugen_create(const ko_node_desc *desc, uint32_t sample_rate, uint32_t block_size);
void ko_ugen_destroy(ko_ugen *u);
int ko_ugen_set_delay(ko_ugen *u, uint32_t param, uint32_t delay);
int ko_ugen_param(ko_ugen *u, uint32_t param, const ko_event *value);
/* in: [n_in][frames], out: [n_out][frames] (block) */
int ko_ugen_process_block(ko_ugen *u, const float *in, float *out, uint32_t frames);
/* in: [n_in], out: [n_out] (one frame) */
int ko_ugen_process(ko_ugen *u, const float *in, float *out);

#ifdef __cplusplus
}
#endif
#endif
