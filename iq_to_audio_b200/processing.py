"""Drop-in stage classes with the reference's names and call signatures, running on the GPU.

Mirrors the public surface of ``src/iq_to_audio/processing.py`` for the hot path
(SURVEY.md 8b): `ComplexOscillator`, `OverlapSaveFIR`, `Decimator`,
`design_channel_filter`, `choose_mix_sign`, `tune_chunk_size`.  Arrays in and out
are numpy (as in the reference); every `process`/`mix` call goes through the C ABI
into a CUDA kernel -- there is no numpy fallback for the arithmetic.

The fused multi-target path is `iq_to_audio_b200.bank.ChannelBank`; the pipeline
object that drives it is in `iq_to_audio_b200.pipeline`.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
from scipy.signal import firwin, kaiser_beta

from . import _lib

_DEVICE = 0


def set_default_device(device: int) -> None:
    global _DEVICE
    _DEVICE = int(device)


def tune_chunk_size(sample_rate: float, requested: int) -> int:
    """Chunk-size heuristic (ref: processing.py:65-81): at least `requested`, else the next
    power of two above 0.25 / 0.40 / 0.50 s of input, capped at 4 Mi samples."""
    floor_ = max(1, requested)
    if sample_rate <= 0:
        return floor_
    span = 0.50 if sample_rate >= 5_000_000.0 else 0.40 if sample_rate >= 2_000_000.0 else 0.25
    wanted = int(round(sample_rate * span))
    if wanted <= floor_:
        return floor_
    ceiling = 4_194_304
    wanted = min(ceiling, max(floor_, wanted))
    return int(min(max(1 << math.ceil(math.log2(wanted)), floor_), ceiling))


def channel_decimation(sample_rate: float, fs_ch_target: float) -> tuple[int, float]:
    """Decimation factor and channel rate (ref: processing.py:885-890)."""
    d = max(1, int(round(sample_rate / fs_ch_target)))
    if sample_rate / d > fs_ch_target * 1.5:
        d = max(int(math.floor(sample_rate / fs_ch_target)), 1)
    return d, sample_rate / d


def design_channel_filter(sample_rate: float, bandwidth: float, decimation: int) -> np.ndarray:
    """Kaiser-windowed low-pass for one channel (ref: processing.py:599-620).  Host-side, once
    per target; uses the same scipy design routine so the taps are bit-identical."""
    # every product below is formed in the reference's order of operations: (fs / (2 D)) * 0.9 and 0.9 * fs / (2 D)
    # differ by one ulp for many (fs, D), and firwin would turn that into different taps
    transition = max(1_000.0, bandwidth * 0.5)
    edge = min(bandwidth * 0.5 * 1.05, (sample_rate / (2.0 * max(decimation, 1))) * 0.9)
    if edge <= 0:
        raise ValueError("Invalid cutoff frequency for channel filter.")
    length = int(np.clip(4.0 / max(transition / sample_rate, 1e-8), 1024, 32768))
    if length % 2 == 0:
        length += 1
    return np.asarray(firwin(length, cutoff=edge, window=("kaiser", kaiser_beta(80.0)), fs=sample_rate),
                      dtype=np.float64)


class ComplexOscillator:
    """Frequency translation by a continuous complex exponential (ref: processing.py:282-297)."""

    def __init__(self, freq_offset_hz: float, sample_rate: float):
        self.phase = 0.0
        self.increment = -2.0 * np.pi * freq_offset_hz / sample_rate

    def mix(self, samples: np.ndarray, sign: int) -> np.ndarray:
        if samples.size == 0:
            return samples
        x = np.ascontiguousarray(samples, dtype=np.complex64)
        out = np.empty_like(x)
        w = sign * self.increment
        _lib.check(_lib.load().iq2a_unpack_mix(x.ctypes.data, x.size, _lib.CODEC_IDS["complex64"],
                                               _lib.ORDER_IDS["iq"], float(self.phase), float(w),
                                               out.ctypes.data, _DEVICE))
        self.phase = (self.phase + sign * self.increment * samples.size) % (2.0 * np.pi)
        return out


class OverlapSaveFIR:
    """Streaming FIR with the reference's semantics (ref: processing.py:300-346): output length
    equals input length, zero initial history, history = last ntaps-1 input samples."""

    def __init__(self, taps: np.ndarray, block_size: int, *, workers: int | None = None):
        if block_size <= 0:
            raise ValueError("block_size must be positive")
        self._real_taps = np.ascontiguousarray(np.real(taps), dtype=np.float64)
        self.taps = np.asarray(taps).astype(np.complex128)
        self.filter_len = len(taps)
        self.overlap = self.filter_len - 1
        self.block_size = block_size
        self.fft_size = 1 << math.ceil(math.log2(self.block_size + self.filter_len - 1))
        self.workers = workers if workers and workers > 1 else None
        self.state = np.zeros(self.overlap, dtype=np.complex64)

    @property
    def taps_fft(self) -> np.ndarray:
        padded = np.zeros(self.fft_size, dtype=np.complex128)
        padded[: self.filter_len] = self.taps
        return np.fft.fft(padded)

    def process(self, samples: np.ndarray) -> np.ndarray:
        if samples.size == 0:
            return samples
        x = np.ascontiguousarray(samples, dtype=np.complex64)
        out = np.empty_like(x)
        hist = np.ascontiguousarray(self.state, dtype=np.complex64)
        _lib.check(_lib.load().iq2a_fir(x.ctypes.data, x.size, hist.ctypes.data if self.overlap else None,
                                        self._real_taps.ctypes.data_as(C.POINTER(C.c_double)),
                                        self.filter_len, out.ctypes.data, _DEVICE))
        if self.overlap:
            joined = x if x.size >= self.overlap else np.concatenate([self.state, x])
            self.state = joined[-self.overlap:].copy()
        return out


class Decimator:
    """Keep samples whose global index is a multiple of `factor` (ref: processing.py:349-360)."""

    def __init__(self, factor: int):
        self.factor = max(1, factor)
        self.offset = 0

    def process(self, samples: np.ndarray) -> np.ndarray:
        if self.factor == 1 or samples.size == 0:
            return samples
        x = np.ascontiguousarray(samples, dtype=np.complex64)
        cap = x.size // self.factor + 1
        out = np.empty(cap, dtype=np.complex64)
        n_out = C.c_int64(0)
        _lib.check(_lib.load().iq2a_decimate(x.ctypes.data, x.size, self.factor, self.offset, out.ctypes.data,
                                             C.byref(n_out), _DEVICE))
        self.offset = (self.offset + samples.size) % self.factor
        return out[: n_out.value]


def choose_mix_sign(warmup: np.ndarray, sample_rate: float, freq_offset: float, taps: np.ndarray,
                    decimation: int, *, device: int | None = None) -> int:
    """Pick the mixer sign that leaves more power in the channel (ref: processing.py:623-663).

    Both candidate signs are run as two channels of one bank over the warm-up snippet; the
    mean power of the settled channel samples is compared in float64 on the host, with the
    reference's tie-break (strictly greater wins, +1 first)."""
    return choose_mix_signs(warmup, sample_rate, [freq_offset], taps, decimation, device=device)[0]


def choose_mix_signs(warmup: np.ndarray, sample_rate: float, freq_offsets, taps: np.ndarray,
                     decimation: int, *, device: int | None = None) -> list[int]:
    """`choose_mix_sign` for several targets at once: all 2 x T candidates are channels of ONE bank (a bank per
    target costs ~40 ms of set-up and tear-down each, more than the probing itself)."""
    offsets = [float(f) for f in freq_offsets]
    if warmup.size == 0:
        return [1] * len(offsets)
    from .bank import ChannelBank, Target
    span = max(int(sample_rate * 0.05), len(taps) * 4, 131_072)
    take = min(warmup.size, span)
    if take < len(taps):
        take = min(warmup.size, len(taps) * 2)
    snippet = np.ascontiguousarray(warmup[:take], dtype=np.complex64)
    d = max(decimation, 1)
    cands = (1, -1)
    with ChannelBank(sample_rate, d, [Target(off, taps, s, "iq") for off in offsets for s in cands],
                     codec="complex64", ref_chunk=max(take, 1), device=_DEVICE if device is None else int(device)) as bank:
        bb = bank.process_chunk(snippet, want_baseband=True).baseband
    signs = []
    for t in range(len(offsets)):
        best_sign, best_power = 1, -np.inf
        for row, sign in zip(bb[2 * t:2 * t + 2], cands):
            if row.size == 0:
                power = -np.inf
            else:
                settled = row[min(len(taps), row.size // 4):]
                if settled.size == 0:
                    settled = row
                power = float(np.mean(np.abs(settled.astype(np.complex128)) ** 2))
            if power > best_power:
                best_power, best_sign = power, sign
        signs.append(best_sign)
    return signs
