"""Host logic: centre-frequency detection.  Same cases and expected answers as the reference's tests/test_utils.py
(file names of SDR++ / SDR# captures, tag priority, unit handling), table-driven."""
from pathlib import Path

import pytest

from iq_to_audio_b200 import utils

NAME_CASES = [
    # file name, Hz, source
    ("baseband_456834049Hz_14-17-29_09-10-2025.wav", 456_834_049.0, "filename:sdrpp"),
    ("16-04-05_457026566Hz.wav", 457_026_566.0, "filename:sdrsharp"),
    ("capture_460MHZ.wav", 460e6, "filename"),
    ("notes_125.5MHz_462.612MHz.wav", 462.612e6, "filename"),          # several candidates: the largest wins
    ("bench_fc-400000000Hz.wav", 400e6, "filename"),                   # benchmark.py's synthetic capture
    ("sweep_2.4GHz_48kHz.wav", 2.4e9, "filename"),
    ("capture.wav", None, "unavailable"),
    ("tone_440Hz.wav", None, "unavailable"),                           # below 1 kHz: not a tuner frequency
]


@pytest.mark.parametrize("name,hz,source", NAME_CASES)
def test_file_name_heuristics(name, hz, source):
    found = utils.detect_center_frequency(Path(name))
    assert found.source == source
    assert found.value == (pytest.approx(hz) if hz is not None else None)
    assert utils.parse_center_frequency(Path(name)) == found.value


TAG_CASES = [
    # tags, file name, Hz, source
    ({"center_frequency": "462.5 MHz"}, "capture_433MHz.wav", 462.5e6, "metadata:center_frequency"),
    ({"Tuner_Freq_Note": "1,250,000 Hz", "title": "999 MHz"}, "capture_433MHz.wav", 1.25e6, "metadata:tuner_freq_note"),
    ({"frequency": "0", "hz": "145_500_000"}, "capture.wav", 145.5e6, "metadata:hz"),
    ({"file": "/data/baseband_456834049Hz.wav"}, "baseband_456834049Hz.wav", 456_834_049.0, "filename:sdrpp"),
]


@pytest.mark.parametrize("tags,name,hz,source", TAG_CASES)
def test_tags_come_first_when_they_name_a_frequency(monkeypatch, tmp_path, tags, name, hz, source):
    target = tmp_path / name
    target.write_bytes(b"")
    monkeypatch.setattr(utils, "_metadata_tags", lambda _path: tags)
    found = utils.detect_center_frequency(target)
    assert (found.value, found.source) == (pytest.approx(hz), source)
