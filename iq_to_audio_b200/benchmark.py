"""`--benchmark`: synthetic capture + timed pipeline run.

Keeps the contract of the reference harness (src/iq_to_audio/benchmark.py:41-127): same keyword
arguments, same synthetic input (tone at the target offset + AWGN from seed 42, clipped to +-0.999,
16-bit stereo WAV whose name carries the centre frequency), only `ProcessingPipeline.run()` inside
the timed region, the "Benchmark processed ... IQ samples in ... s (...x realtime)." report line,
return code 0.  Adds the throughput in Msamples/s.
"""
from __future__ import annotations

import logging
import math
import tempfile
import time
import wave
from collections.abc import Mapping
from dataclasses import dataclass
from pathlib import Path

import numpy as np

from .pipeline import ProcessingConfig, ProcessingPipeline

LOG = logging.getLogger(__name__)
_DEFAULT_CENTER_HZ = 400_000_000.0
_WAV_BLOCK = 1 << 22


@dataclass(frozen=True)
class _Tuning:
    center: float
    target: float

    @property
    def offset(self) -> float:
        return self.target - self.center


def _resolve_tuning(center: float | None, target: float | None, offset: float) -> _Tuning:
    """Whichever of centre / target is missing follows from the requested offset."""
    if center is None and target is None:
        center = _DEFAULT_CENTER_HZ
    if center is None:
        center = target - offset
    if target is None:
        target = center + offset
    return _Tuning(float(center), float(target))


def write_synthetic_capture(path: Path, sample_rate: float, seconds: float, freq_offset: float, *,
                            amplitude: float = 0.7, noise_std: float = 0.02) -> int:
    """Tone + noise capture as PCM_16 stereo WAV; returns the number of complex samples."""
    count = int(round(sample_rate * seconds))
    if count <= 0:
        raise ValueError("Benchmark duration is too short to generate samples.")
    noise = np.random.default_rng(42).normal(scale=noise_std, size=(count, 2))     # one draw for the whole capture
    with wave.open(str(path), "wb") as wav:
        wav.setparams((2, 2, int(sample_rate), 0, "NONE", "not compressed"))
        for lo in range(0, count, _WAV_BLOCK):
            hi = min(count, lo + _WAV_BLOCK)
            # every float64 operation in the reference's order (benchmark.py:31-37): t = n / fs, the phase as
            # ((2 pi) * offset) * t, the tone through the complex exponential -- a phase formed as
            # (2 pi offset / fs) * n differs in the last bits at n ~ 1e7 and flips a few int16 roundings
            t = np.arange(lo, hi, dtype=np.float64) / sample_rate
            tone = np.exp(1j * 2.0 * math.pi * freq_offset * t)
            i = amplitude * tone.real + noise[lo:hi, 0]
            q = amplitude * tone.imag + noise[lo:hi, 1]
            frame = np.clip(np.column_stack((i, q)).astype(np.float32), -0.999, 0.999)
            # libsndfile's float -> PCM_16 (pcm.c f2s_array): lrintf(x * 32767.0f), float32 arithmetic
            wav.writeframes(np.rint(frame * np.float32(32767.0)).astype("<i2").tobytes())
    return count


def _check_request(seconds: float, sample_rate: float, freq_offset: float) -> None:
    if seconds <= 0:
        raise ValueError("Benchmark duration must be positive.")
    if sample_rate <= 0:
        raise ValueError("Benchmark sample rate must be positive.")
    if abs(freq_offset) >= 0.5 * sample_rate:
        raise ValueError("Benchmark offset must be within half the sample rate.")


def run_benchmark(*, seconds: float, sample_rate: float, freq_offset: float, center_freq: float | None,
                  target_freq: float | None, base_kwargs: Mapping[str, object] | None) -> int:
    _check_request(seconds, sample_rate, freq_offset)
    options = {k: v for k, v in (base_kwargs or {}).items() if k != "target_freqs"}
    requested_mode = options.get("demod_mode")
    mode = requested_mode.lower() if isinstance(requested_mode, str) else "nfm"
    tuning = _resolve_tuning(center_freq, target_freq, freq_offset)
    LOG.info("Running benchmark: %.2f s at %.2f MS/s, demod=%s, offset %.1f kHz", seconds, sample_rate / 1e6,
             mode.upper(), tuning.offset / 1e3)
    with tempfile.TemporaryDirectory() as scratch:
        scratch_dir = Path(scratch)
        capture = scratch_dir / f"benchmark_fc-{int(tuning.center)}Hz.wav"
        write_synthetic_capture(capture, sample_rate, seconds, tuning.offset)
        options |= dict(target_freq=tuning.target, center_freq=tuning.center, center_freq_source="benchmark",
                        demod_mode=mode, probe_only=False, output_path=scratch_dir / f"benchmark_audio_{mode}.wav")
        pipeline = ProcessingPipeline(ProcessingConfig(in_path=capture, **options))
        started = time.perf_counter()
        outcome = pipeline.run(progress_sink=None)          # the only timed statement, as in the reference
        elapsed = time.perf_counter() - started
    total = sample_rate * seconds
    speed = seconds / elapsed if elapsed > 0 else float("inf")
    LOG.info("Benchmark processed %.0f IQ samples in %.2f s (%.2f× realtime).", total, elapsed, speed)
    LOG.info("Throughput %.1f Msamples/s end to end (file read + H2D + kernels + D2H + encode).",
             total / max(elapsed, 1e-12) / 1e6)
    LOG.info("Channel decimation %d -> %.1f Hz; audio peak %.2f dBFS.", outcome.decimation, outcome.fs_channel,
             20.0 * math.log10(max(outcome.audio_peak, 1e-6)))
    return 0


__all__ = ["run_benchmark", "write_synthetic_capture"]
