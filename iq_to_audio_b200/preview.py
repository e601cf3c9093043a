"""Preview snapshots for the interactive front end, with the spectrum work on the GPU.

Same entry points and result type as the reference's worker helpers
(``src/iq_to_audio/interactive/workers.py:36-161`` gather_snapshot, ``:164-300`` compute_full_psd,
``interactive/models.py:37-49`` SnapshotData): the capture is read in raw PCM chunks and pushed straight to the
device accumulator (sample-format conversion, IQ order, windows, transforms and averaging all happen there), so
no complex64 copy of the stream is made on the host except the bounded `samples` excerpt the front end keeps for
its own re-plots.
"""
from __future__ import annotations

import contextlib
import math
from collections.abc import Callable
from dataclasses import dataclass
from pathlib import Path
from typing import Any

import numpy as np

from . import _lib
from .pipeline import (IQReader, ProcessingConfig, SampleRateProbe, center_frequency_from_filename, resolve_input)
from .processing import tune_chunk_size
from .spectrum import SpectrumAccumulator, WaterfallResult


@dataclass
class SnapshotData:
    path: Path
    sample_rate: float
    center_freq: float
    probe: SampleRateProbe
    seconds: float
    mode: str
    freqs: np.ndarray
    psd_db: np.ndarray
    waterfall: tuple[np.ndarray, np.ndarray, np.ndarray] | None
    samples: np.ndarray | None
    params: dict[str, Any]
    fft_frames: int


def waterfall_to_tuple(waterfall: WaterfallResult | None):
    if waterfall is None:
        return None
    return (np.asarray(waterfall.freqs, dtype=np.float64), np.asarray(waterfall.times, dtype=np.float32),
            np.asarray(waterfall.matrix, dtype=np.float32))


def _resolve(config: ProcessingConfig, what: str):
    manual = config.input_sample_rate
    if manual is not None and manual <= 0:
        raise ValueError("Input sample rate override must be positive.")
    fmt = resolve_input(Path(config.in_path), config.input_format, config.input_container, manual)
    if fmt.container == "raw" and manual is None:
        raise ValueError("Raw IQ inputs require a sample rate override before previewing.")
    fs = float(manual) if manual is not None else float(fmt.sample_rate)
    center = config.center_freq
    if center is None:
        center = center_frequency_from_filename(Path(config.in_path))
        if center is None:
            raise ValueError("Center frequency not provided and could not be inferred from WAV metadata or "
                             f"filename. {what}")
    return fmt, fs, SampleRateProbe(wave=fs), float(center)


def _to_complex(raw: np.ndarray, fmt, iq_order: str, device: int) -> np.ndarray:
    n = raw.size // fmt.bytes_per_frame
    out = np.empty(n, dtype=np.complex64)
    _lib.check(_lib.load().iq2a_unpack_mix(raw.ctypes.data, n, _lib.CODEC_IDS[fmt.codec], _lib.ORDER_IDS[iq_order],
                                           0.0, 0.0, out.ctypes.data, device))
    return out


def gather_snapshot(config: ProcessingConfig, seconds: float, *, nfft: int, hop: int | None, max_slices: int,
                    fft_workers: int | None = None, max_in_memory_samples: int = 0,
                    progress_cb: Callable[[float, float], None] | None = None) -> SnapshotData:
    """Stream the first `seconds` of the capture into an averaged PSD + capped waterfall."""
    fmt, fs, probe, center = _resolve(config, "Provide --fc or enter a value before previewing.")
    total = int(max(1, round(fs * seconds)))
    hop = max(1, hop or nfft // 4)
    chunk = max(tune_chunk_size(fs, config.chunk_size), nfft)
    retain = min(int(max_in_memory_samples), total)
    kept: list[np.ndarray] = []
    kept_n = consumed = 0
    last = -1.0
    with SpectrumAccumulator(fs, nfft=nfft, hop=hop, max_slices=max_slices, codec=fmt.codec,
                             iq_order=config.iq_order, device=config.device) as acc, \
            IQReader(Path(config.in_path), chunk, config.iq_order, fmt, sample_rate=fs) as reader:
        left = total
        while left > 0:
            raw = reader.read_raw_block(left)
            if raw is None:
                break
            n = raw.size // fmt.bytes_per_frame
            left -= n
            consumed += n
            if kept_n < retain:
                take = min(retain - kept_n, n)
                kept.append(_to_complex(raw[: take * fmt.bytes_per_frame], fmt, config.iq_order, config.device))
                kept_n += take
            acc.push(raw)
            if progress_cb:
                with contextlib.suppress(Exception):
                    frac = min(consumed / total, 1.0)
                    if frac - last >= 0.02 or frac >= 0.999:
                        progress_cb(consumed / fs, frac)
                        last = frac
        freqs, avg, wf, frames = acc.finish()
    samples = np.concatenate(kept) if kept else None
    params = {"nfft": nfft, "hop": hop, "max_slices": max_slices, "fft_workers": fft_workers, "seconds": seconds,
              "full_capture": False, "max_in_memory_samples": max_in_memory_samples}
    snap = SnapshotData(path=Path(config.in_path), sample_rate=fs, center_freq=center, probe=probe,
                        seconds=consumed / fs, mode="samples" if samples is not None else "precomputed",
                        freqs=freqs, psd_db=avg, waterfall=waterfall_to_tuple(wf), samples=samples, params=params,
                        fft_frames=frames)
    if progress_cb:
        with contextlib.suppress(Exception):
            progress_cb(snap.seconds, 1.0)
    return snap


def compute_full_psd(config: ProcessingConfig, *, nfft: int, hop: int, max_slices: int,
                     fft_workers: int | None = None,
                     status_cb: Callable[[str], None] | None = None) -> SnapshotData:
    """Averaged PSD + waterfall of the whole recording."""
    fmt, fs, probe, center = _resolve(config, "Enter a value before using full-record preview.")
    chunk = max(int(config.chunk_size), nfft)
    try:
        payload = max(Path(config.in_path).stat().st_size - fmt.data_offset, 0)
    except OSError:
        payload = 0
    est_total = payload // fmt.bytes_per_frame
    est_chunks = int(math.ceil(est_total / chunk)) if est_total else 0
    stride = max(1, est_chunks // 25) if est_chunks else 4
    if status_cb:
        with contextlib.suppress(Exception):
            status_cb("Reading full recording for spectrum analysis…")
    consumed = idx = 0
    with SpectrumAccumulator(fs, nfft=nfft, hop=hop, max_slices=max_slices, codec=fmt.codec,
                             iq_order=config.iq_order, device=config.device) as acc, \
            IQReader(Path(config.in_path), chunk, config.iq_order, fmt, sample_rate=fs) as reader:
        while True:
            raw = reader.read_raw_block()
            if raw is None:
                break
            consumed += raw.size // fmt.bytes_per_frame
            idx += 1
            if status_cb and (idx == 1 or idx % stride == 0 or (est_chunks and idx >= est_chunks)):
                with contextlib.suppress(Exception):
                    if est_chunks:
                        status_cb(f"Averaging PSD chunk {idx}/{est_chunks} "
                                  f"({min(idx / est_chunks, 1.0) * 100.0:4.1f}% ≈ {consumed / fs:.1f}s/{est_total / fs:.1f}s)")
                    else:
                        status_cb(f"Averaging PSD chunk {idx}…")
            acc.push(raw)
        freqs, avg, wf, frames = acc.finish()
    params = {"nfft": nfft, "hop": hop, "max_slices": max_slices, "fft_workers": fft_workers,
              "seconds": consumed / fs, "full_capture": True}
    return SnapshotData(path=Path(config.in_path), sample_rate=fs, center_freq=center, probe=probe,
                        seconds=consumed / fs, mode="precomputed", freqs=freqs, psd_db=avg,
                        waterfall=waterfall_to_tuple(wf), samples=None, params=params, fft_frames=frames)
