"""48 kHz output stage on the GPU (K14): the arithmetic of the reference's encode-side ffmpeg
(`-ar 48000 -acodec pcm_s16le`, ref: src/iq_to_audio/processing.py:399-418), i.e. libswresample's
default resampler and its float -> int16 conversion, for several channels at once."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


class Resampler48k:
    def __init__(self, in_rate: int, n_channels: int = 1, *, out_rate: int = 48_000, device: int = 0):
        self._lib = _lib.load()
        self.in_rate, self.out_rate, self.n_channels = int(in_rate), int(out_rate), int(n_channels)
        h = C.c_void_p()
        _lib.check(self._lib.iq2a_resampler_create(self.in_rate, self.out_rate, self.n_channels, int(device), C.byref(h)))
        self._h = h

    def close(self) -> None:
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.iq2a_resampler_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _cap(self, n: int) -> int:
        k = C.c_int64(0)
        _lib.check(self._lib.iq2a_resampler_max_outputs(self._h, int(n), C.byref(k)))
        return int(k.value)

    def process(self, audio: np.ndarray) -> np.ndarray:
        """float32 [C, n] (or [n] for one channel) -> int16 [C, m]: every output whose window is complete."""
        a = np.ascontiguousarray(np.atleast_2d(audio), dtype=np.float32)
        if a.shape[0] != self.n_channels:
            raise ValueError(f"expected {self.n_channels} channel rows, got {a.shape[0]}")
        cap = self._cap(a.shape[1])
        out = np.empty((self.n_channels, cap), dtype=np.int16)
        k = C.c_int64(0)
        _lib.check(self._lib.iq2a_resampler_process(self._h, a.ctypes.data, a.shape[1], a.shape[1], out.ctypes.data,
                                                    cap, C.byref(k)))
        return out[:, : k.value]

    def flush(self) -> np.ndarray:
        cap = self._cap(0)
        out = np.empty((self.n_channels, cap), dtype=np.int16)
        k = C.c_int64(0)
        _lib.check(self._lib.iq2a_resampler_flush(self._h, out.ctypes.data, cap, C.byref(k)))
        return out[:, : k.value]
