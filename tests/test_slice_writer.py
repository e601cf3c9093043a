"""Host logic: pass-through slice writer, mirroring tests/test_processing.py:72-96 of the reference."""
from __future__ import annotations

import struct
import wave

import numpy as np
import pytest

from iq_to_audio_b200.input_formats import get_format
from iq_to_audio_b200.pipeline import IQSliceWriter, SampleRateProbe, encode_iq_frames


def test_sample_rate_probe_value_prefers_wave():
    assert SampleRateProbe(ffprobe=None, header=None, wave=48_000.0).value == 48_000.0
    assert SampleRateProbe(ffprobe=44_100.0, header=96_000.0).value == 96_000.0


def test_iq_slice_writer_raw_roundtrip(tmp_path):
    spec = get_format("raw", "pcm_s16le")
    path = tmp_path / "slice.cs16"
    writer = IQSliceWriter(path, 48_000.0, spec)
    samples = np.array([1.0 + 0.0j, -0.5 + 0.5j], dtype=np.complex64)
    writer.write(samples)
    writer.close()
    assert path.read_bytes() == encode_iq_frames(samples, "pcm_s16le")
    # the reference's rule, spelled out (processing.py:533-535): clip to [-1, 0.999969], * 32767, truncate
    assert np.frombuffer(path.read_bytes(), "<i2").tolist() == [32765, 0, -16383, 16383]
    assert writer.peak == pytest.approx(1.0)


@pytest.mark.parametrize("codec,dtype,want", [("pcm_u8", np.uint8, [255, 128, 64, 191]),
                                              ("pcm_f32le", "<f4", [1.0, 0.0, -0.5, 0.5])])
def test_raw_codecs(tmp_path, codec, dtype, want):
    path = tmp_path / "slice.bin"
    w = IQSliceWriter(path, 48_000.0, get_format("raw", codec))
    w.write(np.array([1.0 + 0.0j, -0.5 + 0.5j], dtype=np.complex64))
    w.write(np.empty(0, np.complex64))
    w.close()
    assert np.frombuffer(path.read_bytes(), dtype).tolist() == want            # round((x + 1) * 127.5) for u8


def test_iq_slice_writer_wav_roundtrip(tmp_path):
    spec = get_format("wav", "pcm_s16le")
    path = tmp_path / "slice.wav"
    writer = IQSliceWriter(path, 32_000.0, spec)
    samples = np.array([0.25 + 0.1j, -0.75 - 0.2j, 0.5 + 0j], dtype=np.complex64)
    writer.write(samples)
    writer.close()
    with wave.open(str(path), "rb") as w:
        assert (w.getframerate(), w.getnchannels(), w.getsampwidth(), w.getnframes()) == (32_000, 2, 2, samples.size)
        data = np.frombuffer(w.readframes(3), "<i2").reshape(-1, 2) / 32768.0
    np.testing.assert_allclose(data[:, 0], samples.real, atol=1e-3)
    np.testing.assert_allclose(data[:, 1], samples.imag, atol=1e-3)


def test_wav_float_slices_have_a_fact_chunk(tmp_path):
    path = tmp_path / "slice_f32.wav"
    w = IQSliceWriter(path, 96_153.8, get_format("wav", "pcm_f32le"))
    x = (np.arange(10) / 10.0 + 0.5j).astype(np.complex64)
    w.write(x[:4]); w.write(x[4:]); w.close()
    raw = path.read_bytes()
    assert raw[:4] == b"RIFF" and struct.unpack("<I", raw[4:8])[0] == len(raw) - 8
    assert raw[12:16] == b"fmt " and struct.unpack("<HHI", raw[20:28]) == (3, 2, 96_154)
    assert raw[36:40] == b"fact" and struct.unpack("<I", raw[44:48])[0] == 10
    assert raw[48:52] == b"data" and struct.unpack("<I", raw[52:56])[0] == 80
    np.testing.assert_array_equal(np.frombuffer(raw[56:], "<f4").view(np.complex64), x)
    with pytest.raises(ValueError, match="Unsupported"):
        encode_iq_frames(x, "pcm_s24le")
