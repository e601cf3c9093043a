"""Spectrum previews on the GPU: same names and results as the reference's ``spectrum.py``.

``compute_psd`` (ref: src/iq_to_audio/spectrum.py:15-45) and ``streaming_waterfall`` (:54-92) keep the
reference's signatures, return types and ValueErrors; the transforms, the dB conversion, the running mean and the
capped slice list run in float64 on the device (csrc/spectrum.cu).  ``fft_workers`` is accepted and ignored: it
only selects SciPy's thread count in the reference.  nfft must be a power of two (the front end offers
65536 ... 524288); anything else raises ValueError instead of silently computing on the CPU.
"""
from __future__ import annotations

import ctypes as C
from collections.abc import Iterable
from dataclasses import dataclass

import numpy as np

from . import _lib


def _freq_axis(nfft: int, sample_rate: float) -> np.ndarray:
    # fftshift(fftfreq(nfft, d=1/fs)) for even nfft: (k - nfft/2) / (nfft * d), formed the way numpy does
    k = np.arange(-(nfft // 2), nfft - nfft // 2, dtype=np.int64)
    return (k * (1.0 / (nfft * (1.0 / sample_rate)))).astype(np.float64)


def _as_frames(samples: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(samples).ravel(), dtype=np.complex64)


def compute_psd(samples: np.ndarray, sample_rate: float, nfft: int = 1 << 18, *,
                fft_workers: int | None = None, device: int = 0) -> tuple[np.ndarray, np.ndarray]:
    """Single Hann-windowed PSD in dBFS/Hz of the first ``nfft`` complex samples (zero padded when shorter)."""
    del fft_workers
    x = _as_frames(samples)
    if x.size == 0:
        raise ValueError("Cannot compute PSD for an empty signal.")
    lib = _lib.load()
    out = np.empty(int(nfft), dtype=np.float64)
    _lib.check(lib.iq2a_psd(x.ctypes.data, x.size, _lib.CODEC_IDS["complex64"], _lib.ORDER_IDS["iq"], int(nfft),
                            float(sample_rate), out.ctypes.data, int(device)))
    return _freq_axis(int(nfft), float(sample_rate)), out


@dataclass
class WaterfallResult:
    freqs: np.ndarray
    times: np.ndarray
    matrix: np.ndarray


class SpectrumAccumulator:
    """Incremental form of ``streaming_waterfall``: push chunks, then ``finish()``.

    ``codec``/``iq_order`` let the reader hand over raw PCM frames (half the PCIe bytes of complex64)."""

    def __init__(self, sample_rate: float, *, nfft: int, hop: int | None = None, max_slices: int = 400,
                 codec: str = "complex64", iq_order: str = "iq", device: int = 0):
        self._lib = _lib.load()
        self.nfft, self.sample_rate = int(nfft), float(sample_rate)
        self.hop = max(1, hop or self.nfft // 4)
        self._codec = _lib.CODEC_IDS[codec]
        h = C.c_void_p()
        _lib.check(self._lib.iq2a_spectrum_create(self.nfft, self.hop, int(max_slices), self.sample_rate, self._codec,
                                                  _lib.ORDER_IDS[iq_order], int(device), C.byref(h)))
        self._h = h

    def push(self, chunk: np.ndarray | bytes | None) -> None:
        if chunk is None:
            return
        if isinstance(chunk, (bytes, bytearray, memoryview)):
            buf = np.frombuffer(chunk, dtype=np.uint8)
        elif self._codec == _lib.CODEC_IDS["complex64"] and np.asarray(chunk).dtype != np.uint8:
            buf = _as_frames(chunk).view(np.uint8)
        else:
            buf = np.ascontiguousarray(chunk).view(np.uint8).ravel()
        n = buf.size // _lib.FRAME_BYTES[self._codec]
        if n == 0:
            return
        _lib.check(self._lib.iq2a_spectrum_push(self._h, buf.ctypes.data, n))

    def counts(self) -> tuple[int, int, int]:
        frames, slices, launches = C.c_int64(0), C.c_int32(0), C.c_int64(0)
        _lib.check(self._lib.iq2a_spectrum_counts(self._h, C.byref(frames), C.byref(slices), C.byref(launches)))
        return int(frames.value), int(slices.value), int(launches.value)

    def finish(self) -> tuple[np.ndarray, np.ndarray, WaterfallResult, int]:
        frames, slices, _ = self.counts()
        if frames == 0:
            raise ValueError("Input did not contain enough samples for one FFT frame.")
        avg = np.empty(self.nfft, dtype=np.float64)
        times = np.empty(slices, dtype=np.float32)
        matrix = np.empty((slices, self.nfft), dtype=np.float32)
        _lib.check(self._lib.iq2a_spectrum_result(self._h, avg.ctypes.data, times.ctypes.data, matrix.ctypes.data))
        freqs = _freq_axis(self.nfft, self.sample_rate)
        return freqs.copy(), avg, WaterfallResult(freqs=freqs, times=times, matrix=matrix), frames

    def close(self) -> None:
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.iq2a_spectrum_destroy(h)

    def __enter__(self) -> "SpectrumAccumulator":
        return self

    def __exit__(self, *exc) -> None:
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def streaming_waterfall(chunks: Iterable[np.ndarray | None], sample_rate: float, *, nfft: int,
                        hop: int | None = None, max_slices: int = 400, fft_workers: int | None = None,
                        device: int = 0) -> tuple[np.ndarray, np.ndarray, WaterfallResult, int]:
    """Averaged PSD, capped waterfall and frame count for a stream of complex64 blocks."""
    del fft_workers
    with SpectrumAccumulator(sample_rate, nfft=nfft, hop=hop, max_slices=max_slices, device=device) as acc:
        for chunk in chunks:
            acc.push(chunk)
        return acc.finish()
