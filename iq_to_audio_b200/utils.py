"""Capture metadata helpers with the reference's names (``src/iq_to_audio/utils.py:17-56``, ``:154-200``,
``:267-300``): centre-frequency detection from tags and from SDR++ / SDR# style file names."""
from __future__ import annotations

import re
from dataclasses import dataclass
from pathlib import Path

_UNIT = {"": 1.0, "k": 1e3, "m": 1e6, "g": 1e9}
_NAME_FREQ = re.compile(r"(?i)(\d+(?:\.\d+)?)([kmg]?)hz")
_TEXT_FREQ = re.compile(r"([-+]?\d+(?:\.\d+)?)\s*([kKmMgG]?)\s*(?:[Hh][Zz])?")
_TAG_PRIORITY = ("center_frequency", "centerfrequency", "frequency", "tuner_frequency", "tunerfrequency",
                 "carrier_frequency", "rx_frequency", "hz")


@dataclass
class CenterFrequencyResult:
    value: float | None
    source: str = "unavailable"


def _metadata_tags(path: Path) -> dict[str, str]:
    """Free-form tags of the capture, lower-cased keys.  The reference collects them through libsndfile and
    ffprobe; neither is used here, and SDR++ / SDR# captures carry the frequency in the file name, so this is the
    hook a host with a tag reader fills in."""
    return {}


def _frequency_from_text(text: str | None) -> float | None:
    """'462.5 MHz', '456834049', '1_250_000 Hz' -> Hz (ref :267-291)"""
    if text is None:
        return None
    cleaned = text.strip().replace(",", "").replace("_", "")
    if not cleaned:
        return None
    try:
        plain = float(cleaned)
    except ValueError:
        plain = None
    if plain is not None and plain > 0:
        return plain
    hit = _TEXT_FREQ.search(cleaned)
    if hit is None:
        return None
    value = float(hit.group(1)) * _UNIT.get(hit.group(2).lower(), 1.0)
    return value if value > 0 else None


def _from_tags(path: Path) -> CenterFrequencyResult | None:
    tags = {k.lower(): v for k, v in _metadata_tags(path).items()}
    ordered = [k for k in _TAG_PRIORITY if k in tags]
    ordered += [k for k in tags if k not in _TAG_PRIORITY and ("freq" in k or "hz" in k)]
    for key in ordered:
        freq = _frequency_from_text(tags[key])
        if freq:
            return CenterFrequencyResult(freq, f"metadata:{key}")
    return None


def _from_name(path: Path) -> CenterFrequencyResult | None:
    """Largest '<number>[k|M|G]Hz' token of at least 1 kHz in the file name (ref :179-200)."""
    best = None
    for number, unit in _NAME_FREQ.findall(path.name):
        value = float(number) * _UNIT[unit.lower()]
        if value >= 1_000.0 and (best is None or value > best):
            best = value
    if best is None:
        return None
    stem = path.stem.lower()
    if stem.startswith("baseband_"):
        source = "filename:sdrpp"
    elif re.match(r"\d{2}-\d{2}-\d{2}_", stem):
        source = "filename:sdrsharp"
    else:
        source = "filename"
    return CenterFrequencyResult(best, source)


def detect_center_frequency(path: Path) -> CenterFrequencyResult:
    """Tags first, then the file name."""
    path = Path(path)
    return _from_tags(path) or _from_name(path) or CenterFrequencyResult(None, "unavailable")


def parse_center_frequency(path: Path) -> float | None:
    return detect_center_frequency(path).value
