"""knaster_b200: a B200-native batched render engine for knaster's audio graph.

The package holds only what the hot path needs:

* ``ugens`` / ``graph``   host-side mirror of knaster's node descriptions and of its
                          Graph / GraphEdit / Parameter builder surface;
* ``processor``           ``AudioProcessor``: the render surface (processor.rs:47-197),
                          backed by the CUDA engine through the C ABI in
                          ``include/knaster_gpu.h`` (``csrc/`` holds the kernels);
* ``multi_gpu``           voice sharding across ranks + NCCL mix-bus reduce.

There is no CPU fallback: creating an ``AudioProcessor`` fails loudly when the
CUDA library is missing.
"""
from .graph import (Graph, GraphEdit, GraphError, Parameter, ParameterError, ParameterSmoothing, PTrigger,
                    SchedulingEvent, Seconds, SH, Time)
from .ugens import (BrownNoise, Constant, EnvAr, EnvAsr, Envelope, EnvelopeSegment, Math1Op, Math1UGen, MathOp, MathUGen,
                    OnePoleHpf, OnePoleLpf, Pan2, Phasor, PinkNoise, PolyBlep, RandomLin, SinNumeric, SinWt, SvfFilter,
                    SvfFilterType, TestInPlusParamUGen, TestNumUGen, UGen, Waveform, WhiteNoise,
                    next_randomness_seed, reset_randomness_seed)

__all__ = [n for n in dir() if not n.startswith("_")]


def __getattr__(name):  # lazy: importing the package must not require the CUDA library
    if name in ("AudioProcessor", "AudioProcessorOptions"):
        from . import processor

        return getattr(processor, name)
    raise AttributeError(name)
