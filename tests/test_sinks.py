"""Output sink: 16-bit WAVE files as Buffer::save_to_disk writes them (buffer.rs:315-331)."""
import wave

import numpy as np

from knaster_b200 import sinks


def test_pcm16_conversion_matches_rust_as_i16():
    x = np.array([0.0, 1.0, -1.0, 0.5, -0.5, 1.5, -1.5, 3.05e-5, -3.05e-5, np.nan, 0.99999], dtype=np.float32)
    audio = x.reshape(1, 1, -1)
    want = []
    for v in x:
        s = np.float32(v) * np.float32(32767.0)
        want.append(0 if np.isnan(s) else int(max(-32768, min(32767, np.trunc(s)))))  # truncate toward zero, saturate
    assert sinks.to_pcm16(audio).tolist() == want
    assert want[:7] == [0, 32767, -32767, 16383, -16383, 32767, -32768]


def test_wave_file_round_trip(tmp_path):
    rng = np.random.Generator(np.random.PCG64(5))
    audio = rng.uniform(-1.0, 1.0, (7, 2, 64)).astype(np.float32)   # [blocks][channels][frames]
    path = tmp_path / "out.wav"
    sinks.save_to_disk(audio, str(path), 48000)
    with wave.open(str(path), "rb") as w:
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (2, 2, 48000, 7 * 64)
        pcm = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").reshape(-1, 2)
    left = audio[:, 0, :].reshape(-1)
    right = audio[:, 1, :].reshape(-1)
    assert np.array_equal(pcm[:, 0], np.trunc(left * np.float32(32767.0)).astype(np.int16))
    assert np.array_equal(pcm[:, 1], np.trunc(right * np.float32(32767.0)).astype(np.int16))


def test_stem_dumps_one_mono_file_per_tap(tmp_path):
    rng = np.random.Generator(np.random.PCG64(6))
    taps = rng.uniform(-1.0, 1.0, (3, 500)).astype(np.float32)
    paths = sinks.save_stems(taps, str(tmp_path / "voice_{:03d}.wav"), 48000)
    assert [p.split("/")[-1] for p in paths] == ["voice_000.wav", "voice_001.wav", "voice_002.wav"]
    for i, p in enumerate(paths):
        with wave.open(p, "rb") as w:
            assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (1, 2, 48000, 500)
            pcm = np.frombuffer(w.readframes(500), dtype="<i2")
        assert np.array_equal(pcm, sinks.to_pcm16(taps[i].reshape(1, 1, -1)))


def test_streamed_wave_equals_the_one_shot_file(tmp_path):
    rng = np.random.Generator(np.random.PCG64(7))
    audio = rng.uniform(-1.0, 1.0, (9, 2, 64)).astype(np.float32)
    sinks.save_to_disk(audio, str(tmp_path / "a.wav"), 44100)
    with sinks.WavStream(str(tmp_path / "b.wav"), 2, 44100) as w:
        w.write(audio[:4])
        w.write(audio[4:5])
        w.write(audio[5:])
    assert (tmp_path / "a.wav").read_bytes() == (tmp_path / "b.wav").read_bytes()
