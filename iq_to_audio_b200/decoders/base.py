"""Decoder plug-in contract (ref: src/iq_to_audio/decoders/base.py:9-37): `setup(fs)`,
`process(complex64[]) -> (float32[], DecoderStats | None)`, `finalize()`, `intermediates()`."""
from __future__ import annotations

import ctypes as C
from abc import ABC, abstractmethod
from dataclasses import dataclass

import numpy as np

from .. import _lib


@dataclass
class DecoderStats:
    """Per-chunk decoder statistic."""

    rms_dbfs: float


class Decoder(ABC):
    """Interface every demodulator implements; the pipeline only talks to this."""

    name: str = "decoder"

    @abstractmethod
    def setup(self, sample_rate: float) -> None: ...

    @abstractmethod
    def finalize(self) -> None: ...

    @abstractmethod
    def process(self, samples: np.ndarray) -> tuple[np.ndarray, DecoderStats | None]: ...

    def intermediates(self) -> dict[str, tuple[np.ndarray, float]]:
        return {}


class _GpuChannelDecoder(Decoder):
    """Shared plumbing: one channel's carried state + a call into iq2a_demod."""

    _mode_id = 0
    _agc = False
    _alpha = 0.0
    device = 0

    def __init__(self) -> None:
        self._state = _lib.ChannelState.fresh()
        self._sample_rate = 0.0
        self._last_stats: DecoderStats | None = None
        self._intermediates: dict[str, tuple[np.ndarray, float]] = {}

    def setup(self, sample_rate: float) -> None:
        self._sample_rate = sample_rate

    def finalize(self) -> None:
        return

    @property
    def last_stats(self) -> DecoderStats | None:
        return self._last_stats

    def intermediates(self) -> dict[str, tuple[np.ndarray, float]]:
        return dict(self._intermediates)

    def _run(self, samples: np.ndarray) -> tuple[np.ndarray, DecoderStats]:
        if self._sample_rate == 0.0:
            raise RuntimeError("Decoder.setup(sample_rate) must be called before processing data.")
        x = np.ascontiguousarray(samples, dtype=np.complex64)
        audio = np.empty(x.size, dtype=np.float32)
        rms = C.c_double(0.0)
        if x.size:
            _lib.check(_lib.load().iq2a_demod(self._mode_id, 1 if self._agc else 0, float(self._alpha),
                                              x.ctypes.data, x.size, C.byref(self._state), audio.ctypes.data,
                                              C.byref(rms), self.device))
            stats = DecoderStats(rms_dbfs=float(rms.value))
        else:
            stats = DecoderStats(rms_dbfs=20.0 * np.log10(np.sqrt(1e-18) + 1e-12))
        self._last_stats = stats
        return audio, stats


def run_scan(kind: int, alpha: float, x: np.ndarray, state: "_lib.ChannelState", device: int = 0) -> np.ndarray:
    xin = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(xin)
    if xin.size:
        _lib.check(_lib.load().iq2a_scan(kind, float(alpha), xin.ctypes.data, xin.size, C.byref(state),
                                         out.ctypes.data, device))
    return out
