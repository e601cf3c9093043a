"""Host logic: centre-frequency detection, mirroring the reference's tests/test_utils.py."""
from pathlib import Path

import pytest

from iq_to_audio_b200 import utils
from iq_to_audio_b200.utils import detect_center_frequency, parse_center_frequency


def test_parse_center_frequency_success():
    assert parse_center_frequency(Path("baseband_456834049Hz_14-17-29_09-10-2025.wav")) == 456834049.0


def test_parse_center_frequency_missing():
    assert parse_center_frequency(Path("capture.wav")) is None
    assert detect_center_frequency(Path("capture.wav")).source == "unavailable"


def test_detect_center_frequency_filename_sdrsharp():
    result = detect_center_frequency(Path("16-04-05_457026566Hz.wav"))
    assert result.value == pytest.approx(457026566.0) and result.source == "filename:sdrsharp"


def test_detect_center_frequency_prefers_metadata(monkeypatch, tmp_path):
    target = tmp_path / "capture_433MHz.wav"
    target.write_bytes(b"")
    monkeypatch.setattr(utils, "_metadata_tags", lambda _path: {"center_frequency": "462.5 MHz"})
    result = detect_center_frequency(target)
    assert result.value == pytest.approx(462_500_000.0) and result.source == "metadata:center_frequency"
    monkeypatch.setattr(utils, "_metadata_tags", lambda _path: {"Tuner_Freq_Note": "1,250,000 Hz", "title": "999 MHz"})
    result = detect_center_frequency(target)
    assert result.value == pytest.approx(1_250_000.0) and result.source == "metadata:tuner_freq_note"


def test_detect_center_frequency_ignores_non_frequency_metadata(monkeypatch, tmp_path):
    target = tmp_path / "baseband_456834049Hz.wav"
    target.write_bytes(b"")
    monkeypatch.setattr(utils, "_metadata_tags", lambda _path: {"file": str(target)})
    result = detect_center_frequency(target)
    assert result.value == pytest.approx(456_834_049.0) and result.source == "filename:sdrpp"


def test_detect_center_frequency_handles_uppercase_units():
    result = detect_center_frequency(Path("capture_460MHZ.wav"))
    assert result.value == pytest.approx(460_000_000.0) and result.source.startswith("filename")


def test_detect_center_frequency_multiple_candidates_picks_largest():
    result = detect_center_frequency(Path("notes_125.5MHz_462.612MHz.wav"))
    assert result.value == pytest.approx(462_612_000.0) and result.source.startswith("filename")


def test_small_tokens_and_the_benchmark_name():
    assert parse_center_frequency(Path("tone_440Hz.wav")) is None                      # below 1 kHz: not a tuner frequency
    assert parse_center_frequency(Path("bench_fc-400000000Hz.wav")) == 400_000_000.0     # benchmark.py capture name
    assert parse_center_frequency(Path("sweep_2.4GHz_48kHz.wav")) == 2.4e9
