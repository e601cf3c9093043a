"""`--benchmark`: synthetic capture + timed pipeline run (ref: src/iq_to_audio/benchmark.py:41-127).

Same inputs as the reference's harness -- tone at the target offset, AWGN with seed 42, clipped to
+-0.999, PCM_16 stereo WAV named `benchmark_fc-<fc>Hz.wav` -- and the same log line
("Benchmark processed N IQ samples in T s (X x realtime)."), plus the throughput in Msamples/s.
Only `ProcessingPipeline.run()` is timed, as in the reference (benchmark.py:107-109)."""
from __future__ import annotations

import logging
import math
import tempfile
import time
import wave
from collections.abc import Mapping
from pathlib import Path
from typing import Any

import numpy as np

from .pipeline import ProcessingConfig, ProcessingPipeline

LOG = logging.getLogger(__name__)


def write_synthetic_capture(path: Path, sample_rate: float, seconds: float, freq_offset: float, *,
                            amplitude: float = 0.7, noise_std: float = 0.02) -> int:
    count = int(round(sample_rate * seconds))
    if count <= 0:
        raise ValueError("Benchmark duration is too short to generate samples.")
    rng = np.random.default_rng(42)
    with wave.open(str(path), "wb") as out:
        out.setnchannels(2)
        out.setsampwidth(2)
        out.setframerate(int(sample_rate))
        noise = rng.normal(scale=noise_std, size=(count, 2))      # one draw, as the reference does
        step = 1 << 22
        for s in range(0, count, step):
            t = np.arange(s, min(s + step, count), dtype=np.float64) / sample_rate
            lo = np.exp(1j * 2.0 * math.pi * freq_offset * t)
            iq = np.column_stack((amplitude * lo.real + noise[s:s + t.size, 0],
                                  amplitude * lo.imag + noise[s:s + t.size, 1])).astype(np.float32)
            pcm = np.round(np.clip(iq, -0.999, 0.999) * 32767.0).astype("<i2")
            out.writeframes(pcm.tobytes())
    return count


def run_benchmark(*, seconds: float, sample_rate: float, freq_offset: float, center_freq: float | None,
                  target_freq: float | None, base_kwargs: Mapping[str, object] | None) -> int:
    if seconds <= 0:
        raise ValueError("Benchmark duration must be positive.")
    if sample_rate <= 0:
        raise ValueError("Benchmark sample rate must be positive.")
    if abs(freq_offset) >= sample_rate / 2.0:
        raise ValueError("Benchmark offset must be within half the sample rate.")
    mode = (base_kwargs or {}).get("demod_mode")
    mode = mode.lower() if isinstance(mode, str) else "nfm"
    if center_freq is not None and target_freq is not None:
        offset = target_freq - center_freq
    elif center_freq is not None:
        target_freq, offset = center_freq + freq_offset, freq_offset
    elif target_freq is not None:
        center_freq, offset = target_freq - freq_offset, freq_offset
    else:
        center_freq = 400_000_000.0
        target_freq, offset = center_freq + freq_offset, freq_offset
    LOG.info("Running benchmark: %.2f s at %.2f MS/s, demod=%s, offset %.1f kHz", seconds, sample_rate / 1e6,
             mode.upper(), offset / 1e3)
    with tempfile.TemporaryDirectory() as tmp:
        tmp_path = Path(tmp)
        capture = tmp_path / f"benchmark_fc-{int(center_freq)}Hz.wav"
        write_synthetic_capture(capture, sample_rate, seconds, offset)
        kwargs: dict[str, Any] = dict(base_kwargs) if base_kwargs is not None else {}
        kwargs.pop("target_freqs", None)
        kwargs.update(target_freq=target_freq, center_freq=center_freq, center_freq_source="benchmark",
                      demod_mode=mode, output_path=tmp_path / f"benchmark_audio_{mode}.wav", probe_only=False)
        pipeline = ProcessingPipeline(ProcessingConfig(in_path=capture, **kwargs))
        t0 = time.perf_counter()
        result = pipeline.run(progress_sink=None)
        elapsed = time.perf_counter() - t0
    iq_samples = sample_rate * seconds
    realtime = seconds / elapsed if elapsed > 0 else float("inf")
    LOG.info("Benchmark processed %.0f IQ samples in %.2f s (%.2f× realtime).", iq_samples, elapsed, realtime)
    LOG.info("Throughput %.1f Msamples/s end to end (file read + H2D + kernels + D2H + encode).",
             iq_samples / max(elapsed, 1e-12) / 1e6)
    LOG.info("Channel decimation %d -> %.1f Hz; audio peak %.2f dBFS.", result.decimation, result.fs_channel,
             20.0 * math.log10(max(result.audio_peak, 1e-6)))
    return 0


__all__ = ["run_benchmark", "write_synthetic_capture"]
