"""Output sinks for rendered audio: the step after the hot path.

``save_to_disk`` mirrors ``Buffer::save_to_disk`` (knaster_core_dsp/src/dsp/buffer.rs:315-331): a
16-bit PCM WAVE file, samples interleaved by frame, each ``(x * i16::MAX) as i16`` -- Rust's ``as``
truncates toward zero and saturates (NaN -> 0)."""
from __future__ import annotations

import struct

import numpy as np


def to_pcm16(audio: np.ndarray) -> np.ndarray:
    """[n_blocks, channels, block] f32 (the engine's RawContiguousBlock layout per block,
    block.rs:449-456) -> interleaved int16 [frames * channels]."""
    a = np.asarray(audio, dtype=np.float32)
    if a.ndim != 3:
        raise ValueError("expected [n_blocks, channels, block_size]")
    inter = np.ascontiguousarray(a.transpose(0, 2, 1)).reshape(-1)          # frame-major, channel-minor
    scaled = inter * np.float32(32767.0)                                    # F::from(i16::MAX)
    scaled = np.where(np.isnan(scaled), np.float32(0.0), scaled)
    return np.clip(np.trunc(scaled), -32768.0, 32767.0).astype(np.int16)    # `as i16`


def save_to_disk(audio: np.ndarray, path: str, sample_rate: int) -> None:
    """Write `audio` as a 16-bit PCM WAVE file (hound::WavSpec {bits_per_sample: 16, Int})."""
    a = np.asarray(audio)
    channels = a.shape[1]
    pcm = to_pcm16(a)
    data = pcm.astype("<i2").tobytes()
    block_align = channels * 2
    header = b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE"
    header += b"fmt " + struct.pack("<IHHIIHH", 16, 1, channels, sample_rate, sample_rate * block_align, block_align, 16)
    header += b"data" + struct.pack("<I", len(data))
    with open(path, "wb") as f:
        f.write(header)
        f.write(data)


def save_stems(taps: np.ndarray, path_pattern: str, sample_rate: int) -> list:
    """Per-voice stem dumps: `taps` is what ``AudioProcessor.read_taps()`` returns, [n_taps, frames] f32 (one
    pre-mix signal per ``add_tap``); tap i goes to ``path_pattern.format(i)`` as a mono 16-bit WAVE file in
    the same sample format as ``save_to_disk``.  Returns the paths written."""
    t = np.asarray(taps, dtype=np.float32)
    if t.ndim != 2:
        raise ValueError("expected [n_taps, frames]")
    paths = []
    for i in range(t.shape[0]):
        path = path_pattern.format(i)
        save_to_disk(t[i].reshape(1, 1, -1), path, sample_rate)
        paths.append(path)
    return paths


class WavStream:
    """Streaming form of ``save_to_disk``: append rendered audio ([n_blocks, channels, block] f32) call by call
    -- e.g. one ``AudioProcessor.render(n)`` at a time for renders that do not fit in memory -- and patch the
    RIFF sizes on ``close()``.  Same sample format (16-bit PCM, ``(x * i16::MAX) as i16``)."""

    def __init__(self, path: str, channels: int, sample_rate: int):
        self.channels, self.sample_rate, self._bytes = int(channels), int(sample_rate), 0
        self._f = open(path, "wb")
        self._f.write(self._header(0))

    def _header(self, n: int) -> bytes:
        block_align = self.channels * 2
        return (b"RIFF" + struct.pack("<I", 36 + n) + b"WAVE" + b"fmt "
                + struct.pack("<IHHIIHH", 16, 1, self.channels, self.sample_rate, self.sample_rate * block_align, block_align, 16)
                + b"data" + struct.pack("<I", n))

    def write(self, audio: np.ndarray) -> None:
        a = np.asarray(audio)
        if a.ndim != 3 or a.shape[1] != self.channels:
            raise ValueError(f"expected [n_blocks, {self.channels}, block_size]")
        data = to_pcm16(a).astype("<i2").tobytes()
        self._f.write(data)
        self._bytes += len(data)

    def close(self) -> None:
        if self._f:
            self._f.seek(0)
            self._f.write(self._header(self._bytes))
            self._f.close()
            self._f = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
