"""Host pipeline for the B200 path: the reference's `ProcessingPipeline.run` loop body
(``src/iq_to_audio/processing.py:741-1211``) driving a `ChannelBank` instead of the per-target
numpy stage objects.

What is kept from the reference surface: `ProcessingConfig` field names, `ProcessingPipeline(config)
.run(progress_sink) -> ProcessingResult`, `.cancel()`, `ProcessingCancelled`, `IQReader`,
`AudioWriter`, the decimation / chunk / filter / mix-sign derivations, per-chunk cancel polling and
progress calls, preview truncation (`max_input_seconds`), removal of a cancelled run's output.
What is new: `ProcessingConfig.target_freqs` -- several targets run as ONE pass over the capture
(the reference runs them as sequential passes, ``cli.py:683-710``), and the capture is read as raw
PCM bytes (the sample-format conversion and IQ order fix happen on the GPU), so the decode-side
ffmpeg subprocess is not needed for WAV/raw inputs.  The encode side is unchanged: float32 audio is
piped to ffmpeg for the 48 kHz resample + PCM_16 encode when ffmpeg is available.

Pass-through (`demod_mode` none/pass/iq) writes the channelised IQ with `IQSliceWriter` in the
capture's own container and sample format (processing.py:542-596, :1114-1121).

Out of scope here: metadata sniffing via ffprobe/soundfile, stage plots, the GUI.
"""
from __future__ import annotations

import logging
import math
import os
import sys
import queue
import re
import shutil
import struct
import subprocess
import threading
import wave
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

from . import _lib
from .bank import ChannelBank, Target
from .input_formats import probe_wav, resolve_input_format
from .utils import detect_center_frequency, parse_center_frequency
from .processing import channel_decimation, choose_mix_signs, design_channel_filter, tune_chunk_size

LOG = logging.getLogger(__name__)

FFMPEG_HINT = ("ffmpeg executable not found. Install FFmpeg and ensure it is on PATH "
               "(or point IQ_TO_AUDIO_FFMPEG at it).")

_CODEC_BY_FORMAT = {"wav-s16": ("wav", "pcm_s16le"), "wav-u8": ("wav", "pcm_u8"), "wav-f32": ("wav", "pcm_f32le"),
                    "raw-cs16": ("raw", "pcm_s16le"), "raw-cu8": ("raw", "pcm_u8"), "raw-cf32": ("raw", "pcm_f32le"),
                    "pcm_s16le": (None, "pcm_s16le"), "pcm_u8": (None, "pcm_u8"), "pcm_f32le": (None, "pcm_f32le")}
_RAW_SUFFIX = {".cu8": "pcm_u8", ".cs16": "pcm_s16le", ".cf32": "pcm_f32le", ".iq": "pcm_s16le"}
_FRAME_BYTES = {"pcm_u8": 2, "pcm_s16le": 4, "pcm_f32le": 8}


@dataclass
class ProcessingConfig:
    """Same fields as the reference's ProcessingConfig (processing.py:38-62) plus `target_freqs`."""
    in_path: Path
    target_freq: float = 0.0
    bandwidth: float = 12_500.0
    center_freq: float | None = None
    center_freq_source: str | None = None
    demod_mode: str = "nfm"
    fs_ch_target: float = 96_000.0
    deemph_us: float = 300.0
    agc_enabled: bool = True
    output_path: Path | None = None
    dump_iq_path: Path | None = None
    chunk_size: int = 1_048_576
    filter_block: int = 65_536
    iq_order: str = "iq"
    probe_only: bool = False
    mix_sign_override: int | None = None
    plot_stages_path: Path | None = None
    fft_workers: int | None = None
    max_input_seconds: float | None = None
    input_container: str | None = None
    input_format: str | None = None
    input_format_source: str | None = None
    input_sample_rate: float | None = None
    target_freqs: list[float] | None = None      # extension: batch of targets in one pass
    device: int = 0


@dataclass
class SampleRateProbe:
    ffprobe: float | None = None
    header: float | None = None
    wave: float | None = None

    @property
    def value(self) -> float:
        for v in (self.wave, self.header, self.ffprobe):
            if v:
                return float(v)
        raise RuntimeError("sample rate unavailable")


@dataclass
class ProcessingResult:
    sample_rate_probe: SampleRateProbe
    center_freq: float
    target_freq: float
    freq_offset: float
    decimation: int
    fs_channel: float
    mix_sign: int
    audio_peak: float
    output_path: Path | None = None
    samples_processed: int = 0


class ProcessingCancelled(RuntimeError):  # noqa: N818
    """Raised when processing is aborted early by user request."""


# ------------------------------------------------------------------------------------------
# input: raw PCM frames straight from the container
# ------------------------------------------------------------------------------------------
@dataclass
class InputFormat:
    container: str      # "wav" | "raw"
    codec: str          # pcm_u8 | pcm_s16le | pcm_f32le
    sample_rate: float | None
    data_offset: int

    @property
    def bytes_per_frame(self) -> int:
        return _FRAME_BYTES[self.codec]


def resolve_input(path: Path, requested: str | None, container_hint: str | None,
                  sample_rate: float | None) -> InputFormat:
    """Effective encoding (input_formats.resolve_input_format: override or detection, same errors as the
    reference) plus what the raw-frame reader needs from a WAV header: payload offset and sample rate.  The data
    chunk's declared length is ignored and the payload is read to end of file -- what `-ignore_length 1` does for
    > 4 GiB SDR++ captures (ref: processing.py:149-150)."""
    spec, _source = resolve_input_format(Path(path), requested=requested, container_hint=container_hint)
    if spec.container == "raw":
        return InputFormat("raw", spec.codec, sample_rate, 0)
    try:
        header = probe_wav(Path(path))
    except RuntimeError as exc:
        raise ValueError(str(exc)) from exc
    if header.channels != 2:
        raise ValueError(f"{path}: expected a 2-channel (I/Q) capture, found {header.channels} channels")
    return InputFormat("wav", spec.codec, float(header.sample_rate), header.data_offset)


class IQReader:
    """Stream a baseband capture in chunks of `chunk_size` complex samples (ref: processing.py:84-279).

    `read_raw_block()` returns the PCM payload bytes of the next chunk (what the fused path eats);
    iterating yields complex64 blocks with the IQ order applied, like the reference, by running the
    unpack kernel with a zero-frequency oscillator."""

    #: buffers of the pinned read ring: the chunk being filled + the (at most) two ChannelBank.stream has in flight
    RING_DEPTH = 3
    READ_THREADS = 4
    PARALLEL_READ_MIN = 8 << 20

    def __init__(self, path: Path, chunk_size: int, iq_order: str, input_format: InputFormat, *,
                 sample_rate: float | None = None, pinned: bool = False, batch: int = 1, device: int = 0):
        if iq_order not in _lib.ORDER_IDS:
            raise ValueError(f"Unsupported iq_order '{iq_order}'")
        self.path = Path(path)
        self.chunk_size = int(chunk_size)
        self.iq_order = iq_order
        self.input_format = input_format
        self.sample_rate = sample_rate
        self.input_bytes_per_frame = input_format.bytes_per_frame
        self.batch = max(1, int(batch))          # reference chunks per read_raw_block(batched=True)
        self.pinned = bool(pinned)
        self.device = int(device)
        self._fh = None
        self._ring: list[np.ndarray] = []
        self._ring_ptrs: list[int] = []
        self._slot = 0
        self._pool = None
        self._pos = 0

    def __enter__(self) -> "IQReader":
        if self.input_format.container == "raw" and not (self.sample_rate and self.sample_rate > 0):
            raise ValueError("Raw IQ inputs require a sample rate override. Provide --input-sample-rate.")
        self._fh = self.path.open("rb", buffering=0)
        self._fh.seek(self.input_format.data_offset)
        self._pos = self.input_format.data_offset
        if self.pinned:
            # page-locked ring (iq2a_host_alloc): the file is read straight into the memory the H2D copy engine
            # takes it from -- no pageable staging copy inside the driver, the copy of chunk k+1 really overlaps
            # the kernels of chunk k (the reference reads through an ffmpeg pipe, processing.py:238-259)
            import ctypes as C
            lib = _lib.load()
            nbytes = self.chunk_size * self.batch * self.input_bytes_per_frame
            for _ in range(self.RING_DEPTH):
                p = C.c_void_p()
                _lib.check(lib.iq2a_host_alloc(C.byref(p), nbytes))
                self._ring_ptrs.append(p.value)
                self._ring.append(np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(nbytes,)))
            if hasattr(os, "preadv") and nbytes >= self.PARALLEL_READ_MIN:
                # one thread copies out of the page cache at ~6 GB/s (1.5 GS/s of int16 IQ), an eighth of what the
                # PCIe link takes: large blocks are read as slices by a few threads (preadv releases the GIL)
                from concurrent.futures import ThreadPoolExecutor
                self._pool = ThreadPoolExecutor(max_workers=self.READ_THREADS, thread_name_prefix="iq2a-read")
        return self

    def __exit__(self, *exc) -> None:
        if self._pool is not None:
            self._pool.shutdown(wait=True)
            self._pool = None
        if self._fh:
            self._fh.close()
            self._fh = None
        if self._ring_ptrs:
            lib = _lib.load()
            self._ring = []
            for p in self._ring_ptrs:
                lib.iq2a_host_free(p)
            self._ring_ptrs = []

    def read_raw_block(self, max_frames: int | None = None, *, batched: bool = False) -> np.ndarray | None:
        """Next chunk of PCM payload bytes (None at the end of the capture); `batched`: up to `batch` reference
        chunks at once.  With a pinned ring the returned view is valid until RING_DEPTH - 1 further reads."""
        if self._fh is None:
            raise RuntimeError("IQReader has not been entered.")
        frames = self.chunk_size * (self.batch if batched else 1)
        if max_frames is not None:
            frames = min(frames, max_frames)
        fb = self.input_bytes_per_frame
        if self._ring:
            buf = self._ring[self._slot][:frames * fb]
            self._slot = (self._slot + 1) % len(self._ring)
        else:
            buf = np.empty(frames * fb, dtype=np.uint8)
        if self._pool is not None and buf.size >= self.PARALLEL_READ_MIN:
            got = self._read_parallel(buf)
        else:
            got = self._fh.readinto(memoryview(buf))
            while 0 < got < buf.size:
                more = self._fh.readinto(memoryview(buf)[got:])
                if not more:
                    break
                got += more
            self._pos += max(got, 0)
        got -= got % fb                      # drop a trailing partial frame (ref: processing.py:253-256)
        return buf[:got] if got > 0 else None

    def _read_parallel(self, buf: np.ndarray) -> int:
        """Fill `buf` from the current file position with READ_THREADS positional reads; returns the bytes read
        (short only at the end of the file) and keeps the sequential file position in step."""
        fd = self._fh.fileno()
        view = memoryview(buf)
        k = self.READ_THREADS
        step = -(-buf.size // k)
        step += -step % 4096
        base = self._pos

        def one(lo: int) -> int:
            hi = min(buf.size, lo + step)
            done = 0
            while lo + done < hi:
                n = os.preadv(fd, [view[lo + done:hi]], base + lo + done)
                if n <= 0:
                    break
                done += n
            return done
        parts = list(self._pool.map(one, range(0, buf.size, step)))
        got = 0
        for i, n in enumerate(parts):            # contiguous prefix: a short slice can only be the file's last one
            got += n
            if n < min(step, buf.size - i * step):
                break
        self._pos = base + got
        self._fh.seek(self._pos)
        return got

    def __iter__(self):
        lib = _lib.load()
        codec_id = _lib.CODEC_IDS[self.input_format.codec]
        while True:
            raw = self.read_raw_block()
            if raw is None:
                break
            n = raw.size // self.input_bytes_per_frame
            out = np.empty(n, dtype=np.complex64)
            _lib.check(lib.iq2a_unpack_mix(raw.ctypes.data, n, codec_id, _lib.ORDER_IDS[self.iq_order], 0.0, 0.0,
                                           out.ctypes.data, self.device))
            yield out


# ------------------------------------------------------------------------------------------
# output: float32 audio -> 48 kHz PCM_16 WAV
# ------------------------------------------------------------------------------------------
def resolve_ffmpeg_executable() -> Path | None:
    override = os.environ.get("IQ_TO_AUDIO_FFMPEG")        # ref: utils.py:126-151
    if override:
        p = Path(override)
        return p if p.exists() else None
    found = shutil.which("ffmpeg")
    return Path(found) if found else None


class AudioWriter:
    """ref: processing.py:381-524.  `write()` takes decoder output (computes peak + clip like the
    reference); `write_clipped()` takes audio the GPU already clipped.  With ffmpeg present the
    float32 stream is piped to it exactly as the reference does (`-f f32le -ar round(fs) ... -acodec
    pcm_s16le -ar 48000`).  Without ffmpeg the writer falls back to a plain PCM_16 WAV at the channel
    rate (no resampling) and says so -- the 48 kHz libswresample-exact stage is the next row of the
    scope table, not part of this one."""

    def __init__(self, output_path: Path, input_rate: float, *, require_ffmpeg: bool = False, device: int = 0):
        self.output_path = Path(output_path)
        self.input_rate = float(input_rate)
        self.ffmpeg_rate = max(1, int(round(self.input_rate)))
        self.peak = 0.0
        self._closed = False
        self._error: BaseException | None = None
        self.proc = None
        self._wav = None
        self._res = None
        ffmpeg = None if os.environ.get("IQ_TO_AUDIO_B200_NATIVE_WAV") == "1" else resolve_ffmpeg_executable()
        if ffmpeg is None and require_ffmpeg:
            raise RuntimeError(FFMPEG_HINT)
        self.output_path.parent.mkdir(parents=True, exist_ok=True)
        if ffmpeg is not None:
            cmd = [str(ffmpeg), "-hide_banner", "-loglevel", "error", "-y", "-f", "f32le", "-ac", "1", "-ar",
                   str(self.ffmpeg_rate), "-i", "-", "-acodec", "pcm_s16le", "-ar", "48000", str(self.output_path)]
            try:
                self.proc = subprocess.Popen(cmd, stdin=subprocess.PIPE, stderr=subprocess.PIPE)
            except OSError as exc:
                raise RuntimeError(f"Failed to launch ffmpeg: {exc}") from exc
            # bounded: the GPU produces audio orders of magnitude faster than ffmpeg takes it, so an unbounded queue
            # would hold the whole backlog in host memory; a full queue back-pressures the producer instead
            self._queue: queue.Queue = queue.Queue(maxsize=64)
            self._thread = threading.Thread(target=self._drain, name="AudioWriter", daemon=True)
            self._thread.start()
        else:
            from .resample import Resampler48k
            LOG.info("Encoding natively: %d Hz float32 -> 48 kHz PCM_16 on the GPU.", self.ffmpeg_rate)
            self._res = Resampler48k(self.ffmpeg_rate, 1, device=device) if self.ffmpeg_rate > 48_000 else None
            if self._res is None and self.ffmpeg_rate != 48_000:
                raise RuntimeError(f"channel rate {self.ffmpeg_rate} Hz is below 48 kHz: ffmpeg is required to upsample")
            self._wav = wave.open(str(self.output_path), "wb")
            self._wav.setnchannels(1)
            self._wav.setsampwidth(2)
            self._wav.setframerate(48_000)

    def _drain(self) -> None:
        pipe = self.proc.stdin
        try:
            while True:
                item = self._queue.get()
                if item is None:
                    break
                pipe.write(item)
        except BaseException as exc:  # broken pipe etc.: surfaced on the next write/close
            self._error = exc

    def write(self, samples: np.ndarray) -> None:
        if samples.size == 0:
            return
        pk = float(np.max(np.abs(samples)))
        self.peak = max(self.peak, pk)
        self.write_clipped(np.clip(samples, -0.99, 0.99).astype(np.float32, copy=False))

    def write_clipped(self, safe: np.ndarray) -> None:
        if self._closed:
            raise RuntimeError("AudioWriter has already been closed.")
        if self._error:
            raise RuntimeError("ffmpeg writer failed") from self._error
        if safe.size == 0:
            return
        if self.proc is not None:
            payload = np.ascontiguousarray(safe, dtype=np.float32).tobytes()
            while True:
                try:
                    self._queue.put(payload, timeout=1.0)
                    break
                except queue.Full:
                    if self._error or not self._thread.is_alive():
                        raise RuntimeError("ffmpeg writer failed") from self._error
        elif self._res is not None:
            self._wav.writeframes(self._res.process(safe)[0].astype("<i2").tobytes())
        else:   # already 48 kHz: swr's flt -> s16 conversion only
            pcm = np.clip(np.rint(safe.astype(np.float32) * np.float32(32768.0)), -32768, 32767).astype("<i2")
            self._wav.writeframes(pcm.tobytes())

    def close(self) -> None:
        if self._closed:
            return
        self._closed = True
        if self.proc is not None:
            # the drain thread must have written everything before stdin closes: wait for it as long as it makes
            # progress (it ends on the sentinel, or on a pipe error that is surfaced below)
            while self._thread.is_alive():
                try:
                    self._queue.put(None, timeout=1.0)
                    break
                except queue.Full:
                    continue
            self._thread.join()
            try:
                self.proc.stdin.close()
            except OSError:
                pass
            try:
                self.proc.wait(timeout=60)
            except subprocess.TimeoutExpired:
                self.proc.terminate()
            if self._error:
                raise RuntimeError("ffmpeg writer failed") from self._error
        elif self._wav is not None:
            if self._res is not None:
                self._wav.writeframes(self._res.flush()[0].astype("<i2").tobytes())
                self._res.close()
            self._wav.close()


# ------------------------------------------------------------------------------------------
# pass-through output: channelised IQ in the capture's own container / sample format
# ------------------------------------------------------------------------------------------
def encode_iq_frames(samples: np.ndarray, codec: str) -> bytes:
    """complex64 -> interleaved PCM frames, the raw-file rule of the reference (processing.py:527-539):
    f32 as is; s16 = trunc(clip(x, -1, 0.999969) * 32767); u8 = round((clip(x, -1, 1) + 1) * 127.5)."""
    pairs = np.ascontiguousarray(samples, dtype=np.complex64).view(np.float32)
    if codec == "pcm_f32le":
        return pairs.astype("<f4", copy=False).tobytes()
    if codec == "pcm_s16le":
        return (np.clip(pairs, -1.0, 0.999969) * 32767.0).astype("<i2").tobytes()
    if codec == "pcm_u8":
        return np.round((np.clip(pairs, -1.0, 1.0) + 1.0) * 127.5).astype(np.uint8).tobytes()
    raise ValueError(f"Unsupported raw codec {codec}")


def _sndfile_frames(samples: np.ndarray, codec: str) -> bytes:
    """What libsndfile stores when float32 frames are written to a PCM_16 / PCM_U8 / FLOAT WAV with its default
    settings (normalisation on, clipping off): lrintf(x * 0x7FFF) resp. lrintf(x * 0x7F) + 128, wrapping on
    overflow.  (soundfile is not installable in the build image: this conversion is restated from libsndfile's
    pcm.c, not pinned against it.)"""
    pairs = np.ascontiguousarray(samples, dtype=np.complex64).view(np.float32)
    if codec == "pcm_f32le":
        return pairs.astype("<f4", copy=False).tobytes()
    if codec == "pcm_s16le":
        return (np.rint(pairs * np.float32(32767.0)).astype(np.int64) & 0xFFFF).astype("<u2").tobytes()
    if codec == "pcm_u8":
        return ((np.rint(pairs * np.float32(127.0)).astype(np.int64) + 128) & 0xFF).astype(np.uint8).tobytes()
    raise ValueError(f"Unsupported WAV codec for slices: {codec}")


class IQSliceWriter:
    """ref: processing.py:542-596.  Writes the decimated channel as 2-channel IQ, WAV when the capture was a WAV
    (same sample format), headerless frames otherwise; tracks the peak magnitude."""

    _WAV_FORMATS = {"pcm_u8": (1, 8), "pcm_s16le": (1, 16), "pcm_f32le": (3, 32)}

    def __init__(self, output_path: Path, sample_rate: float, spec: "InputFormat"):
        self.output_path = Path(output_path)
        self.sample_rate = float(sample_rate)
        self.spec = spec
        self.peak = 0.0
        self._bytes = 0
        if spec.container == "wav" and spec.codec not in self._WAV_FORMATS:
            raise ValueError(f"Unsupported WAV codec for slices: {spec.codec}")
        self._fd = self.output_path.open("wb")
        if spec.container == "wav":
            self._fd.write(self._header(0))

    def _header(self, data_bytes: int) -> bytes:
        tag, bits = self._WAV_FORMATS[self.spec.codec]
        rate = max(1, int(round(self.sample_rate)))
        block = 2 * bits // 8
        fmt = struct.pack("<HHIIHH", tag, 2, rate, rate * block, block, bits)
        body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt
        if tag == 3:
            body += b"fact" + struct.pack("<II", 4, min(data_bytes // block, 0xFFFFFFFF))
        body += b"data" + struct.pack("<I", min(data_bytes, 0xFFFFFFFF))
        return b"RIFF" + struct.pack("<I", min(len(body) + data_bytes, 0xFFFFFFFF)) + body

    def write(self, samples: np.ndarray) -> None:
        if samples.size == 0:
            return
        peak = float(np.max(np.abs(samples)))
        if peak > self.peak:
            self.peak = peak
        if self.spec.container == "wav":
            payload = _sndfile_frames(samples, self.spec.codec)
        else:
            payload = encode_iq_frames(samples, self.spec.codec)
        self._fd.write(payload)
        self._bytes += len(payload)

    def close(self) -> None:
        if self._fd is None:
            return
        if self.spec.container == "wav":
            self._fd.seek(0)
            self._fd.write(self._header(self._bytes))
        self._fd.close()
        self._fd = None


# ------------------------------------------------------------------------------------------
def center_frequency_from_filename(path: Path) -> float | None:
    """SDR++ names captures `baseband_<freq>Hz_<time>.wav`; the benchmark uses `..._fc-<freq>Hz.wav`
    (utils.detect_center_frequency: tags first, then the largest '<n>[k|M|G]Hz' token of the name)."""
    return parse_center_frequency(Path(path))


class _Progress:
    """Duck-typed bridge to the reference's ProgressSink (progress.py:26-53): phases ingest /
    channel / demod / encode advanced once per chunk, status text, cancel callback."""

    def __init__(self, sink, totals: dict[str, float]):
        self.sink = sink
        self.done = {k: 0.0 for k in totals}
        self.totals = totals
        if sink is not None and hasattr(sink, "start"):
            try:
                from types import SimpleNamespace
                self.phases = {k: SimpleNamespace(key=k, label=k, total=v, unit="samples", completed=0.0)
                               for k, v in totals.items()}
                sink.start(list(self.phases.values()), overall_total=float(sum(totals.values())))
            except TypeError:
                self.phases = {}

    def advance(self, key: str, amount: float) -> None:
        if self.sink is None or key not in self.done:
            return
        self.done[key] += amount
        ph = getattr(self, "phases", {}).get(key)
        if ph is not None and hasattr(self.sink, "advance"):
            ph.completed = self.done[key]
            self.sink.advance(ph, amount, overall_completed=float(sum(self.done.values())),
                              overall_total=float(sum(self.totals.values())))

    def status(self, msg: str) -> None:
        if self.sink is not None and hasattr(self.sink, "status"):
            self.sink.status(msg)

    def close(self) -> None:
        if self.sink is not None and hasattr(self.sink, "close"):
            self.sink.close()


class ProcessingPipeline:
    def __init__(self, config: ProcessingConfig):
        self.config = config
        self._cancelled = False

    def cancel(self) -> None:
        self._cancelled = True

    def _check_cancel(self, where: str) -> None:
        if self._cancelled:
            raise ProcessingCancelled(f"Processing cancelled during {where}.")

    def _is_pass_through_mode(self) -> bool:                      # reference: processing.py:693-695
        return (self.config.demod_mode or "").lower() in {"none", "pass", "iq"}

    # reference: processing.py:1214-1233
    def _default_output_path(self, freq: float, fmt: "InputFormat | None" = None) -> Path:
        in_path = Path(self.config.in_path)
        if self._is_pass_through_mode():
            suffix = in_path.suffix
            if fmt is not None and fmt.container == "wav":
                ext = suffix if suffix.lower() in {".wav", ".wave", ".wv", ".rf64"} else ".wav"
            elif fmt is not None and fmt.container == "raw":
                ext = suffix or {"pcm_u8": ".cu8", "pcm_s16le": ".cs16", "pcm_f32le": ".cf32"}.get(fmt.codec, ".raw")
            else:
                ext = suffix or ".wav"
            return in_path.with_name(f"slice_{int(freq)}{ext}")
        return in_path.with_name(f"audio_{int(freq)}_48k.wav")

    @staticmethod
    def _annotate(base: Path | None, freq: float, total: int) -> Path | None:       # cli.py:523-527
        if base is None or total <= 1:
            return None if base is None else Path(base)
        return Path(base).with_name(f"{Path(base).stem}_{int(round(freq))}{Path(base).suffix}")

    def _output_for(self, freq: float, total: int, fmt: "InputFormat | None" = None) -> Path:
        base = self.config.output_path
        if base is None:
            return self._default_output_path(freq, fmt)
        if total <= 1:
            return Path(base)
        return Path(base).with_name(f"{Path(base).stem}_{int(round(freq))}{Path(base).suffix}")   # cli.py:523-527

    def run(self, progress_sink=None) -> ProcessingResult:
        return self.run_many(progress_sink)[0]

    def run_many(self, progress_sink=None) -> list[ProcessingResult]:
        cfg = self.config
        if hasattr(progress_sink, "set_cancel_callback"):
            progress_sink.set_cancel_callback(self.cancel)
        manual_rate = cfg.input_sample_rate
        if manual_rate is not None and manual_rate <= 0:
            raise ValueError("Input sample rate override must be positive.")
        fmt = resolve_input(Path(cfg.in_path), cfg.input_format, cfg.input_container, manual_rate)
        cfg.input_container = cfg.input_container or fmt.container
        cfg.input_format = cfg.input_format or fmt.codec
        if fmt.container == "raw" and manual_rate is None:
            raise ValueError("Raw IQ inputs require --input-sample-rate (CLI) or a manual entry in the GUI.")
        sample_rate = float(manual_rate) if manual_rate is not None else float(fmt.sample_rate)
        probe = SampleRateProbe(wave=sample_rate)

        targets = list(cfg.target_freqs) if cfg.target_freqs else [cfg.target_freq]
        if any(f <= 0 for f in targets) and not cfg.probe_only:
            raise ValueError("Target frequency must be positive. Provide --ft or use --interactive.")
        if cfg.bandwidth <= 0:
            raise ValueError("Bandwidth must be positive.")
        center = cfg.center_freq
        if center is None:
            found = detect_center_frequency(Path(cfg.in_path))
            if found.value is None:
                raise ValueError("Center frequency not supplied and could not be determined from metadata or "
                                 "filename. Use --fc to provide it explicitly.")
            center = found.value
            cfg.center_freq, cfg.center_freq_source = center, found.source

        targets = [f if f > 0 else center for f in targets]      # probe-only without --ft: the centre itself (:880-881)
        decimation, fs_channel = channel_decimation(sample_rate, cfg.fs_ch_target)       # :885-890
        chunk = tune_chunk_size(sample_rate, cfg.chunk_size)                            # :929
        max_samples = None
        if cfg.max_input_seconds is not None and cfg.max_input_seconds > 0:
            max_samples = max(1, int(math.floor(cfg.max_input_seconds * sample_rate)))  # :842-845
        taps = design_channel_filter(sample_rate, cfg.bandwidth, decimation)             # :987
        LOG.info("Input sample rate %.2f Hz; decimation %d -> %.2f Hz; %d-tap channel filter; chunk %d.",
                 sample_rate, decimation, fs_channel, len(taps), chunk)

        try:
            payload = max(Path(cfg.in_path).stat().st_size - fmt.data_offset, 0)
        except OSError:
            payload = 0
        total_in = payload / fmt.bytes_per_frame
        if max_samples is not None:
            total_in = min(total_in, max_samples) if total_in > 0 else max_samples
        prog = _Progress(progress_sink, {"ingest": total_in, "channel": total_in / decimation,
                                         "demod": total_in / decimation,
                                         "encode": total_in / sample_rate * 48_000.0})
        pass_through = self._is_pass_through_mode()
        outputs = [self._output_for(f, len(targets), fmt) for f in targets]
        writers: list = []
        dumps: list = []
        results: list[ProcessingResult] = []
        processed = 0
        try:
            # reference chunks per GPU call: enough rows to fill the persistent kernel's CTA slots at high decimation
            # (one 4 Mi-sample chunk at D = 640 is 7 block sets for 296 slots), bounded by 128 MiB of pinned ring
            # -- about 48 k channel-rate rows per call; a 10 MS/s capture already has 40 k rows per chunk and stays at 2
            batch = max(1, min(8, -(-49_152 * decimation // max(chunk, 1)), (128 << 20) // max(chunk * fmt.bytes_per_frame, 1)))
            if os.environ.get("IQ2A_BIND_NUMA", "1") != "0":
                # the reader thread and the pinned ring next to the GPU's PCIe root (numa.py); affects this thread only
                from .numa import bind_to_device_numa
                LOG.debug("host placement: %s", bind_to_device_numa(cfg.device))
            with IQReader(Path(cfg.in_path), chunk, cfg.iq_order, fmt, sample_rate=sample_rate, pinned=True,
                          batch=batch, device=cfg.device) as reader:
                self._check_cancel("initialization")
                first = reader.read_raw_block(max_samples)
                if first is None:
                    raise RuntimeError("Input stream produced no samples.")
                # mixer sign from the warm-up chunk (processing.py:1028-1042)
                signs = []
                if cfg.mix_sign_override in (1, -1):
                    signs = [cfg.mix_sign_override] * len(targets)
                else:
                    lib = _lib.load()
                    nfr = first.size // fmt.bytes_per_frame
                    warm = np.empty(nfr, dtype=np.complex64)
                    _lib.check(lib.iq2a_unpack_mix(first.ctypes.data, nfr, _lib.CODEC_IDS[fmt.codec],
                                                   _lib.ORDER_IDS[cfg.iq_order], 0.0, 0.0, warm.ctypes.data, cfg.device))
                    signs = choose_mix_signs(warm, sample_rate, [f - center for f in targets], taps, decimation, device=cfg.device)
                LOG.info("Selected mixer sign(s) %s based on warm-up snippet.", signs)
                self._check_cancel("warm-up")
                if cfg.probe_only:
                    prog.advance("ingest", first.size // fmt.bytes_per_frame)
                    return [ProcessingResult(probe, center, f, f - center, decimation, fs_channel, s, 0.0)
                            for f, s in zip(targets, signs)]

                bank = ChannelBank(sample_rate, decimation,
                                   [Target(f - center, taps, s, cfg.demod_mode, cfg.deemph_us, cfg.agc_enabled)
                                    for f, s in zip(targets, signs)],
                                   codec=fmt.codec, iq_order=cfg.iq_order, ref_chunk=chunk, device=cfg.device)
                if pass_through:                                   # processing.py:1014-1015
                    writers = [IQSliceWriter(o, fs_channel, fmt) for o in outputs]
                else:
                    writers = [AudioWriter(o, fs_channel, device=cfg.device) for o in outputs]
                if cfg.dump_iq_path:                               # one cf32 dump per target (cli.py:623, :666)
                    dumps = [self._annotate(Path(cfg.dump_iq_path), f, len(targets)).open("wb") for f in targets]
                want_bb = pass_through or bool(dumps)

                def blocks():
                    nonlocal processed
                    blk = first
                    idx = 0
                    while blk is not None:
                        n = blk.size // fmt.bytes_per_frame
                        self._check_cancel(f"chunk {idx + 1}")
                        prog.advance("ingest", n)
                        prog.status(f"Chunk {idx + 1}: channelize + demodulate")
                        processed += n
                        yield blk
                        idx += 1
                        left = None if max_samples is None else max_samples - processed
                        if left is not None and left <= 0:
                            break
                        blk = reader.read_raw_block(left, batched=True)

                with bank:
                    for res in bank.stream(blocks(), want_baseband=want_bb, want_audio=False,
                                           want_clipped=not pass_through, chunk_frames=chunk):
                        prog.advance("channel", float(res.count))
                        for fd, row in zip(dumps, res.baseband if dumps else ()):   # IQDebugWriter, :363-378
                            fd.write(row.tobytes())
                        prog.advance("demod", float(res.count))
                        if pass_through:                            # processing.py:1114-1121
                            for w, row in zip(writers, res.baseband):
                                w.write(row)
                        else:
                            for w, row in zip(writers, res.clipped):
                                w.write_clipped(row)
                        self._check_cancel("encode")
                        prog.advance("encode", res.count / max(fs_channel, 1e-9) * 48_000.0)
                    peaks = [w.peak for w in writers] if pass_through else bank.peaks
            for w, pk in zip(writers, peaks):
                w.peak = pk
            for f, s, o, pk in zip(targets, signs, outputs, peaks):
                if pass_through:
                    LOG.info("IQ slice peak magnitude %.2f dBFS (complex).", 20.0 * math.log10(max(pk, 1e-6)))
                else:
                    LOG.info("Audio peak level %.2f dBFS.", 20.0 * math.log10(max(pk, 1e-6)))
                results.append(ProcessingResult(probe, center, f, f - center, decimation, fs_channel, s, pk, o, processed))
            return results
        except ProcessingCancelled:
            for w in writers:
                try:
                    w.close()
                except Exception:
                    pass
            writers = []
            for o in outputs:
                try:
                    Path(o).unlink(missing_ok=True)          # ref: processing.py:1205-1211
                except OSError:
                    LOG.debug("Failed to remove cancelled output %s", o)
            raise
        finally:
            # every writer and dump is closed even if one close() raises; the first failure is re-raised
            failure = None
            for w in writers:
                try:
                    w.close()
                except Exception as exc:                      # noqa: BLE001
                    failure = failure or exc
            for fd in dumps:
                fd.close()
            prog.close()
            if failure is not None and sys.exc_info()[0] is None:
                raise failure
