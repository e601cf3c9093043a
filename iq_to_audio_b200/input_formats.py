"""Input encodings of a baseband capture: same names and answers as the reference's ``input_formats.py``
(``src/iq_to_audio/input_formats.py:17-338``), without libsndfile or ffprobe.

The reference asks ``soundfile.info`` (falling back to ``ffprobe``) for the WAV subtype; here the RIFF / RF64
header is read directly (``probe_wav``), which also yields what the raw-frame reader needs: the payload offset and
the sample rate.  Everything downstream of the answer is the GPU's: the frames go to the device exactly as stored.
"""
from __future__ import annotations

import os
import struct
from collections.abc import Iterable
from dataclasses import dataclass
from pathlib import Path

_WAV_SUFFIXES = (".wav", ".wave", ".wv", ".rf64")


@dataclass(slots=True, frozen=True)
class InputFormatSpec:
    container: str                    # "wav" | "raw"
    codec: str                        # pcm_u8 | pcm_s16le | pcm_f32le
    label: str
    bytes_per_frame: int              # one complex sample (I + Q) on disk
    ffmpeg_input_format: str | None   # kept for hosts that still hand the file to ffmpeg
    requires_sample_rate: bool

    @property
    def key(self) -> str:
        return f"{self.container}:{self.codec}"


@dataclass(slots=True)
class InputFormatDetection:
    spec: InputFormatSpec | None
    source: str
    message: str | None = None
    error: str | None = None

    @property
    def ok(self) -> bool:
        return self.spec is not None and self.error is None


def _table() -> dict[tuple[str, str], InputFormatSpec]:
    rows = (("pcm_u8", 2, "u8", "unsigned 8-bit", "u8"), ("pcm_s16le", 4, "s16le", "signed 16-bit", "s16"),
            ("pcm_f32le", 8, "f32le", None, "f32"))
    out = {}
    for codec, nbytes, raw_hint, pcm_name, short in rows:
        wav_label = f"WAV PCM {pcm_name}" if pcm_name else "WAV float32"
        out[("wav", codec)] = InputFormatSpec("wav", codec, wav_label, nbytes, None, False)
        out[("raw", codec)] = InputFormatSpec("raw", codec, f"RAW complex {short} (.c{short})", nbytes, raw_hint, True)
    return out


_FORMATS = _table()
_RAW_BY_SUFFIX = {".cu8": "pcm_u8", ".cs16": "pcm_s16le", ".cf32": "pcm_f32le", ".iq": "pcm_s16le"}
_CODEC_ALIASES = {"pcm_u8": ("u8", "cu8", "s8"), "pcm_s16le": ("s16", "cs16", "pcm16", "pcm_s16"),
                  "pcm_f32le": ("f32", "float32", "cf32")}
_ALIAS_TO_CODEC = {alias: codec for codec, names in _CODEC_ALIASES.items() for alias in names}


def list_supported_formats(container: str | None = None) -> Iterable[InputFormatSpec]:
    return (s for s in _FORMATS.values() if container in (None, s.container))


def get_format(container: str, codec: str) -> InputFormatSpec:
    spec = _FORMATS.get((container, codec))
    if spec is None:
        raise ValueError(f"Unsupported input format: {container}:{codec}")
    return spec


def parse_user_format(value: str, *, default_container: str | None = None) -> tuple[str, str]:
    """'raw:cu8', 'wav-s16', 'f32', 'pcm_s16le' ... -> (container, codec)   (ref :132-171)"""
    text = value.strip().lower()
    if text in ("", "auto"):
        raise ValueError("parse_user_format() expects a non-auto value.")
    container, token = None, text
    for sep in ":-":
        pieces = [p for p in text.split(sep) if p]
        if sep in text and len(pieces) == 2:
            container, token = pieces
            break
    container = container or default_container
    codec = _ALIAS_TO_CODEC.get(token, token.replace(".", ""))
    if ("wav", codec) not in _FORMATS:
        raise ValueError(f"Unsupported input codec override: {value}")
    if container is None:
        container = "raw" if token[:1] == "c" else "wav"        # cu8 / cs16 / cf32 name raw files
    if container not in ("wav", "raw"):
        raise ValueError(f"Unknown input container override: {container}")
    return container, codec


# ------------------------------------------------------------------------------------------
# RIFF / RF64 header
# ------------------------------------------------------------------------------------------
@dataclass(slots=True)
class WavHeader:
    format_tag: int
    channels: int
    sample_rate: int
    bits: int
    data_offset: int          # byte offset of the first frame
    data_bytes: int           # declared payload length (0xFFFFFFFF / 0 when the writer could not know it)

    @property
    def subtype(self) -> str:
        """libsndfile's name for the encoding (what the reference keys on, ref :101-108)."""
        if self.format_tag == 1:
            return {8: "PCM_U8", 16: "PCM_16", 24: "PCM_24", 32: "PCM_32"}.get(self.bits, f"PCM_{self.bits}")
        if self.format_tag == 3:
            return {32: "FLOAT", 64: "DOUBLE"}.get(self.bits, f"FLOAT_{self.bits}")
        return {6: "ALAW", 7: "ULAW"}.get(self.format_tag, f"FORMAT_{self.format_tag:#06x}")


def probe_wav(path: Path) -> WavHeader:
    """Walk the chunks up to 'data'.  Raises RuntimeError (like soundfile.info) when the file is not a readable
    RIFF/WAVE."""
    with Path(path).open("rb") as fh:
        head = fh.read(12)
        if len(head) < 12 or head[:4] not in (b"RIFF", b"RF64", b"BW64") or head[8:12] != b"WAVE":
            raise RuntimeError(f"{path}: not a RIFF/WAVE file")
        fmt = None
        ds64_data = None
        while True:
            hdr = fh.read(8)
            if len(hdr) < 8:
                raise RuntimeError(f"{path}: no data chunk")
            tag, size = hdr[:4], struct.unpack("<I", hdr[4:])[0]
            if tag == b"ds64":
                body = fh.read(size + (size & 1))
                if len(body) >= 16:
                    ds64_data = struct.unpack("<Q", body[8:16])[0]
            elif tag == b"fmt ":
                body = fh.read(size + (size & 1))
                if len(body) < 16:
                    raise RuntimeError(f"{path}: truncated fmt chunk")
                ftag, channels, rate, _, _, bits = struct.unpack("<HHIIHH", body[:16])
                if ftag == 0xFFFE and len(body) >= 26:              # WAVE_FORMAT_EXTENSIBLE: sub-format GUID
                    ftag = struct.unpack("<H", body[24:26])[0]
                fmt = (ftag, channels, rate, bits)
            elif tag == b"data":
                if fmt is None:
                    raise RuntimeError(f"{path}: data chunk before fmt chunk")
                declared = ds64_data if (size == 0xFFFFFFFF and ds64_data is not None) else size
                return WavHeader(fmt[0], fmt[1], fmt[2], fmt[3], fh.tell(), int(declared))
            else:
                fh.seek(size + (size & 1), os.SEEK_CUR)


_SUBTYPE_TO_CODEC = {"PCM_U8": "pcm_u8", "PCM_S8": "pcm_u8", "PCM_16": "pcm_s16le", "FLOAT": "pcm_f32le"}


def detect_input_format(path: Path) -> InputFormatDetection:
    """Encoding from the WAV header, or from the extension for headerless captures (ref :174-241)."""
    path = Path(path)
    suffix = path.suffix.lower()
    raw_codec = _RAW_BY_SUFFIX.get(suffix)
    if raw_codec is not None:
        spec = get_format("raw", raw_codec)
        return InputFormatDetection(spec, f"extension:{suffix}", message=f"Detected {spec.label} via extension.")
    if suffix == ".raw":
        return InputFormatDetection(None, "extension:.raw",
                                    error="Raw '.raw' files need a manual format selection (cu8/cs16/cf32).")
    if suffix not in _WAV_SUFFIXES:
        return InputFormatDetection(None, f"extension:{suffix or 'none'}",
                                    error="Unsupported input type. Provide a WAV/RAW IQ recording.")
    try:
        header = probe_wav(path)
    except (RuntimeError, OSError):
        return InputFormatDetection(None, "header", error="Unable to read WAV header; specify format manually.")
    subtype = header.subtype
    source = f"wav:{subtype.lower()}"
    codec = _SUBTYPE_TO_CODEC.get(subtype)
    if codec is not None:
        return InputFormatDetection(get_format("wav", codec), source, message=f"WAV subtype {subtype} detected.")
    if subtype in ("PCM_24", "PCM_32"):
        return InputFormatDetection(None, source, error="32-bit/24-bit PCM WAV inputs are not supported. "
                                                        "Export as 16-bit or float32.")
    return InputFormatDetection(None, source, error=f"Unsupported WAV subtype {subtype}. "
                                                    "Export as PCM 16-bit or float32.")


def deduce_container(path: Path) -> str:
    return "raw" if Path(path).suffix.lower() in _RAW_BY_SUFFIX else "wav"


def resolve_input_format(path: Path, *, requested: str | None,
                         container_hint: str | None = None) -> tuple[InputFormatSpec, str]:
    """Effective format: the user's override when given, detection otherwise (ref :316-338)."""
    container = container_hint or deduce_container(path)
    if requested and requested.strip().lower() != "auto":
        return get_format(*parse_user_format(requested, default_container=container)), "manual"
    found = detect_input_format(path)
    if found.spec is None:
        raise ValueError(found.error or "Unable to determine input format.")
    return found.spec, found.source
