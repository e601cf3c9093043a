"""Narrow-band FM: discriminator + de-emphasis (ref: src/iq_to_audio/decoders/nfm.py)."""
from __future__ import annotations

import math

import numpy as np

from .. import _lib
from .base import DecoderStats, _GpuChannelDecoder, run_scan


class QuadratureDemod:
    """angle(s[n] * conj(s[n-1])), lag carried across calls (ref: nfm.py:11-24)."""

    def __init__(self):
        self._st = _lib.ChannelState.fresh()

    @property
    def prev(self) -> np.complex64:
        return np.complex64(complex(self._st.prev_re, self._st.prev_im))

    def process(self, samples: np.ndarray) -> np.ndarray:
        if samples.size == 0:
            return np.empty(0, dtype=np.float32)
        import ctypes as C
        x = np.ascontiguousarray(samples, dtype=np.complex64)
        out = np.empty(x.size, dtype=np.float32)
        # de-emphasis pole 0 turns the NFM chain into the bare discriminator (y = 1*x + 0)
        self._st.deemph_z = 0.0
        _lib.check(_lib.load().iq2a_demod(_lib.MODE_IDS["nfm"], 0, 0.0, x.ctypes.data, x.size,
                                          C.byref(self._st), out.ctypes.data, None, 0))
        return out


class DeemphasisFilter:
    """Single-pole de-emphasis y = beta*x + alpha*y1 (ref: nfm.py:27-62)."""

    def __init__(self, tau_us: float, sample_rate: float | None = None):
        self.tau_us = tau_us
        self.alpha = 0.0
        self.beta = 0.0
        self._st = _lib.ChannelState.fresh()
        if sample_rate is not None:
            self.configure(sample_rate)

    def configure(self, sample_rate: float) -> None:
        tau_sec = max(self.tau_us * 1e-6, 1e-6)
        self.alpha = math.exp(-1.0 / (sample_rate * tau_sec))
        self.beta = 1.0 - self.alpha
        self._st.deemph_z = 0.0

    @property
    def state(self) -> float:
        return float(self._st.deemph_z)

    def process(self, samples: np.ndarray) -> np.ndarray:
        if samples.size == 0:
            return samples
        return run_scan(0, self.alpha, samples, self._st)


class NarrowbandFMDecoder(_GpuChannelDecoder):
    name = "narrowband_fm"
    _mode_id = _lib.MODE_IDS["nfm"]

    def __init__(self, deemph_us: float):
        super().__init__()
        self._deemph_us = deemph_us

    def setup(self, sample_rate: float) -> None:
        tau_sec = max(self._deemph_us * 1e-6, 1e-6)
        self._alpha = math.exp(-1.0 / (sample_rate * tau_sec))
        self._state.deemph_z = 0.0
        self._sample_rate = sample_rate

    def process(self, samples: np.ndarray) -> tuple[np.ndarray, DecoderStats | None]:
        audio, stats = self._run(samples)
        if samples.size:
            self._intermediates = {"deemph": (audio.copy(), self._sample_rate),
                                   "audio": (audio.copy(), self._sample_rate)}
        return audio, stats


__all__ = ["DecoderStats", "DeemphasisFilter", "NarrowbandFMDecoder", "QuadratureDemod"]
