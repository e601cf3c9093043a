"""Host logic: input format detection / override parsing, mirroring the reference's tests/test_input_formats.py
(same cases and expectations; WAV files are written with the stdlib instead of soundfile)."""
from __future__ import annotations

import struct
import wave
from pathlib import Path

import numpy as np
import pytest

from iq_to_audio_b200.input_formats import (InputFormatSpec, deduce_container, detect_input_format, get_format,
                                            list_supported_formats, parse_user_format, probe_wav,
                                            resolve_input_format)


def _write_wave(path: Path, tag: int, bits: int, *, channels: int = 2, rate: int = 48_000, frames: int = 32,
                extensible: bool = False, junk: bool = False) -> None:
    block = channels * bits // 8
    data = bytes(frames * block)
    if extensible:
        fmt = struct.pack("<HHIIHHHHIH14s", 0xFFFE, channels, rate, rate * block, block, bits, 22, bits, 3, tag,
                          b"\x00\x00\x00\x00\x10\x00\x80\x00\x00\xaa\x00\x38\x9b\x71")
    else:
        fmt = struct.pack("<HHIIHH", tag, channels, rate, rate * block, block, bits)
    body = b"WAVE"
    if junk:
        body += b"JUNK" + struct.pack("<I", 5) + b"\0" * 6          # odd-sized chunk, padded
    body += b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"data" + struct.pack("<I", len(data)) + data
    path.write_bytes(b"RIFF" + struct.pack("<I", len(body)) + body)


def test_detect_input_format_wav_pcm16(tmp_path):
    path = tmp_path / "pcm16.wav"
    _write_wave(path, 1, 16)
    detection = detect_input_format(path)
    assert detection.ok and detection.spec is not None
    assert detection.spec.codec == "pcm_s16le" and detection.spec.container == "wav"
    assert detection.source == "wav:pcm_16" and detection.spec.bytes_per_frame == 4


def test_detect_input_format_wav_float(tmp_path):
    path = tmp_path / "float.wav"
    _write_wave(path, 3, 32)
    detection = detect_input_format(path)
    assert detection.ok and detection.spec.codec == "pcm_f32le"


def test_detect_input_format_wav_u8_and_extensible(tmp_path):
    _write_wave(tmp_path / "u8.wav", 1, 8)
    assert detect_input_format(tmp_path / "u8.wav").spec.codec == "pcm_u8"
    _write_wave(tmp_path / "ext.wav", 1, 16, extensible=True, junk=True)
    d = detect_input_format(tmp_path / "ext.wav")
    assert d.ok and d.spec.codec == "pcm_s16le"


def test_detect_input_format_wav_int32_rejected(tmp_path):
    path = tmp_path / "int32.wav"
    _write_wave(path, 1, 32)
    detection = detect_input_format(path)
    assert not detection.ok and detection.error is not None and "32-bit" in detection.error
    _write_wave(tmp_path / "d.wav", 3, 64)
    d = detect_input_format(tmp_path / "d.wav")
    assert not d.ok and "Unsupported WAV subtype DOUBLE" in d.error


def test_detect_input_format_raw_extension(tmp_path):
    path = tmp_path / "capture.cu8"
    path.write_bytes(b"")
    detection = detect_input_format(path)
    assert detection.ok and detection.spec.container == "raw" and detection.spec.codec == "pcm_u8"
    assert detection.source == "extension:.cu8" and detection.spec.requires_sample_rate
    assert detect_input_format(tmp_path / "x.iq").spec.codec == "pcm_s16le"
    assert "manual format" in detect_input_format(tmp_path / "x.raw").error
    assert "Unsupported input type" in detect_input_format(tmp_path / "x.mp3").error


@pytest.mark.parametrize("value,expected", [("wav-s16", ("wav", "pcm_s16le")), ("wav:u8", ("wav", "pcm_u8")),
                                            ("raw-cu8", ("raw", "pcm_u8")), ("cf32", ("raw", "pcm_f32le")),
                                            ("pcm_s16le", ("wav", "pcm_s16le")), (" RAW:CS16 ", ("raw", "pcm_s16le")),
                                            ("float32", ("wav", "pcm_f32le"))])
def test_parse_user_format_variants(value, expected):
    assert parse_user_format(value, default_container=None) == expected


def test_parse_user_format_default_container_and_errors():
    assert parse_user_format("s16", default_container="raw") == ("raw", "pcm_s16le")
    with pytest.raises(ValueError):
        parse_user_format("unknown-format", default_container=None)
    with pytest.raises(ValueError, match="non-auto"):
        parse_user_format("auto")
    with pytest.raises(ValueError, match="container"):
        parse_user_format("ogg:s16")


def test_detect_input_format_unreadable_header(tmp_path):
    path = tmp_path / "fallback.wav"
    path.write_bytes(b"not a real wav")
    detection = detect_input_format(path)
    assert not detection.ok and "Unable to read WAV header" in detection.error
    with pytest.raises(ValueError, match="Unable to read WAV header"):
        resolve_input_format(path, requested=None)
    spec, source = resolve_input_format(path, requested="wav-s16")          # the override still stands
    assert (spec.key, source) == ("wav:pcm_s16le", "manual")


def test_format_table_and_resolution(tmp_path):
    specs = list(list_supported_formats())
    assert len(specs) == 6 and all(isinstance(s, InputFormatSpec) for s in specs)
    assert {s.key for s in list_supported_formats("raw")} == {"raw:pcm_u8", "raw:pcm_s16le", "raw:pcm_f32le"}
    assert get_format("raw", "pcm_f32le").ffmpeg_input_format == "f32le" and get_format("wav", "pcm_u8").bytes_per_frame == 2
    with pytest.raises(ValueError, match="Unsupported input format"):
        get_format("wav", "pcm_s24le")
    assert deduce_container(Path("a.cs16")) == "raw" and deduce_container(Path("a.wav")) == "wav" and deduce_container(Path("a")) == "wav"
    _write_wave(tmp_path / "c.wav", 1, 16)
    assert resolve_input_format(tmp_path / "c.wav", requested="auto")[1] == "wav:pcm_16"
    assert resolve_input_format(tmp_path / "c.cf32", requested=None)[0].key == "raw:pcm_f32le"
    # the path's container wins over the 'c' prefix of the token (the reference passes it as default_container)
    assert resolve_input_format(tmp_path / "c.bin", requested="cu8", container_hint=None)[0].key == "wav:pcm_u8"
    assert resolve_input_format(tmp_path / "c.bin", requested="cu8", container_hint="raw")[0].key == "raw:pcm_u8"


def test_probe_wav_offsets_match_the_stdlib_reader(tmp_path):
    path = tmp_path / "sdr.wav"
    pcm = (np.arange(2 * 100) % 251).astype("<i2")
    with wave.open(str(path), "wb") as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(2_500_000)
        w.writeframes(pcm.tobytes())
    h = probe_wav(path)
    assert (h.format_tag, h.channels, h.sample_rate, h.bits, h.data_bytes, h.subtype) == (1, 2, 2_500_000, 16, 400, "PCM_16")
    assert path.read_bytes()[h.data_offset:h.data_offset + 400] == pcm.tobytes()
    # RF64 with the length in a ds64 chunk and 0xFFFFFFFF placeholders (captures beyond 4 GiB)
    fmt = struct.pack("<HHIIHH", 1, 2, 10_000_000, 40_000_000, 4, 16)
    ds64 = struct.pack("<QQQI", 0, 5_000_000_000, 1_250_000_000, 0)
    body = b"WAVE" + b"ds64" + struct.pack("<I", len(ds64)) + ds64 + b"fmt " + struct.pack("<I", 16) + fmt + b"data" + struct.pack("<I", 0xFFFFFFFF) + b"\x01\x02\x03\x04"
    (tmp_path / "big.rf64").write_bytes(b"RF64" + struct.pack("<I", 0xFFFFFFFF) + body)
    h = probe_wav(tmp_path / "big.rf64")
    assert h.data_bytes == 5_000_000_000 and h.sample_rate == 10_000_000
    assert detect_input_format(tmp_path / "big.rf64").spec.codec == "pcm_s16le"
    with pytest.raises(OSError):
        probe_wav(tmp_path / "none.wav")
    (tmp_path / "short.wav").write_bytes(b"RIFF\x00\x00\x00\x00WAVEfmt ")
    with pytest.raises(RuntimeError):
        probe_wav(tmp_path / "short.wav")
