"""CPU oracle for the spectrum previews.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench tooling's CPU leg may import this; the product path
(iq_to_audio_b200/spectrum.py -> csrc/spectrum.cu) never does.

Restates, in numpy float64, what the reference computes in src/iq_to_audio/spectrum.py:
  psd_one              <- compute_psd :15-45
  window_starts        <- _sliding_windows :95-128 (incl. its start-index bookkeeping: after a chunk that yielded
                          windows the index base slips back by the pending length, :108-110 with :125)
  frame_psd            <- _SlidingFFT.psd :159-171
  SliceList            <- _WaterfallAggregator :174-208
  waterfall            <- streaming_waterfall :54-92
Parity status: PINNED -- tests/test_oracle_spectrum.py checks every function against outputs of the reference
module itself (tests/golden/spectrum_vectors.npz, written by tests/golden/make_spectrum_golden.py).
"""
from __future__ import annotations

import numpy as np

EPS = 1e-18


def freq_axis(nfft: int, fs: float) -> np.ndarray:
    return np.fft.fftshift(np.fft.fftfreq(nfft, d=1.0 / fs)).astype(np.float64)


def psd_one(x: np.ndarray, fs: float, nfft: int) -> np.ndarray:
    if x.size == 0:
        raise ValueError("Cannot compute PSD for an empty signal.")
    use = x[:nfft]
    w = np.hanning(use.size).astype(np.float64)
    power = np.sum(w * w) / use.size
    spec = np.fft.fftshift(np.fft.fft(use.astype(np.complex128) * w, n=nfft))
    p = spec * np.conj(spec) / (use.size * fs * power + EPS)
    return 10.0 * np.log10(np.abs(p) + EPS)


def frame_psd(x: np.ndarray, fs: float, w: np.ndarray, power: float) -> np.ndarray:
    n = x.size
    spec = np.fft.fftshift(np.fft.fft(x.astype(np.complex128) * w, n=n))
    p = spec * np.conj(spec) / (n * fs * power + EPS)
    return 10.0 * np.log10(np.abs(p) + EPS)


def window_starts(chunk_sizes, nfft: int, hop: int):
    """(reported start index, true position in the concatenated stream) of every window, in order."""
    out = []
    pend = 0            # samples carried over
    base = 0            # the reference's `offset`
    consumed = 0        # true stream position of the first carried-over sample
    for n in chunk_sizes:
        if n == 0:
            continue
        total = pend + n
        if pend:
            base -= pend
        if total < nfft:
            pend = total
            base += total
            continue
        s = 0
        while s + nfft <= total:
            out.append((base + s, consumed + s))
            s += hop
        keep = max(total - s, 0)
        base += total - keep
        keep = min(keep, nfft)
        consumed += total - keep
        pend = keep
    return out


class SliceList:
    def __init__(self, cap: int):
        self.cap = max(1, int(cap))
        self.rows: list[np.ndarray] = []
        self.t: list[float] = []

    def add(self, row: np.ndarray, t: float) -> None:
        self.rows.append(row.astype(np.float32))
        self.t.append(float(t))
        while len(self.rows) > self.cap:
            rows, ts = [], []
            for i in range(0, len(self.rows), 2):
                if i + 1 < len(self.rows):
                    rows.append(((self.rows[i].astype(np.float64) + self.rows[i + 1].astype(np.float64)) / 2.0)
                                .astype(np.float32))
                else:
                    rows.append(self.rows[i])
                ts.append(self.t[i])
            self.rows, self.t = rows, ts


def waterfall(chunks, fs: float, nfft: int, hop: int | None, max_slices: int):
    """-> (avg_psd_db float64 [nfft], times float32, matrix float32 [slices, nfft], frames)"""
    hop = max(1, hop or nfft // 4)
    chunks = [np.asarray(c, dtype=np.complex64) for c in chunks if c is not None]
    stream = np.concatenate(chunks) if chunks else np.empty(0, np.complex64)
    w = np.hanning(nfft).astype(np.float64)
    power = np.sum(w * w) / nfft
    total = None
    sl = SliceList(max_slices)
    frames = 0
    for reported, pos in window_starts([c.size for c in chunks], nfft, hop):
        row = frame_psd(stream[pos:pos + nfft], fs, w, power)
        total = row.copy() if total is None else total + row
        sl.add(row, reported / fs)
        frames += 1
    if frames == 0:
        raise ValueError("Input did not contain enough samples for one FFT frame.")
    return total / frames, np.asarray(sl.t, dtype=np.float32), np.stack(sl.rows), frames
