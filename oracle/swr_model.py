"""CPU oracle for the 48 kHz output stage (K14).  TEST INFRASTRUCTURE ONLY.

The reference does not resample in Python: `AudioWriter` pipes float32 audio to
`ffmpeg -f f32le -ar round(fs_ch) -i - -acodec pcm_s16le -ar 48000`
(``src/iq_to_audio/processing.py:399-418``), i.e. the arithmetic is libswresample's default
resampler plus its flt -> s16 conversion.  FFmpeg is a third-party dependency that is not in
``/root/reference`` and is version-unpinned there (CI installs "latest"); this module restates
the published algorithm of libswresample (resample.c: `resample_init`, `build_filter`,
`swri_resample` linear-interpolated float path, `invert_initial_buffer`, `resample_flush`) with
all-default options: filter_size 32, phase_shift 10, linear_interp, exact_rational, Kaiser beta 9,
cutoff 0.97, no dither.

Parity status: PINNED against outputs of a real libswresample 6.1.100 (FFmpeg 8 series, the copy
bundled in the opencv wheel of the build image, driven through ctypes by
``tests/golden/make_resampler_golden.py``): float outputs agree to 1e-7, int16 outputs and sample
counts exactly (``tests/test_oracle_resampler.py``).

Facts established against the library (each one was a mismatch until modelled):
  * every phase is normalised by the tap sum of phase 0 (DC gain of other phases is 1 +- 2.7e-6);
  * at the last phase (index == phase_count-1) the interpolation partner is phase 0 at the NEXT
    input sample with all of its taps (the SIMD loop runs over the padded filter length);
  * the stream start is extended by reflection x[-n] = x[n], the end by x[N+j] = x[N-1-j] over
    R = (min(leftover, filter_length) + 1) // 2 samples, where `leftover` is what the last regular
    call could not consume -- so the total count is not simply ceil(N * out / in);
  * flt -> s16 is clip(lrintf(x * 32768)).
"""
from __future__ import annotations

import math

import numpy as np


def _bessel_i0(x: np.ndarray) -> np.ndarray:
    """Modified Bessel function I0 by its power series (exact to double precision for x <= ~20)."""
    x = np.asarray(x, dtype=np.float64)
    q = x * x / 4.0
    term = np.ones_like(q)
    total = np.ones_like(q)
    for k in range(1, 64):
        term = term * q / (k * k)
        total = total + term
    return total


class SwrModel:
    def __init__(self, in_rate: int, out_rate: int = 48_000, *, filter_size: int = 32, phase_shift: int = 10,
                 cutoff: float = 0.97, beta: float = 9.0):
        self.in_rate, self.out_rate = int(in_rate), int(out_rate)
        self.passthrough = self.in_rate == self.out_rate        # swr inserts no resampler at all
        factor = min(out_rate * cutoff / in_rate, 1.0)
        pc = 1 << phase_shift
        fl = max(int(math.ceil(filter_size / factor)), 1)
        if fl > 1:
            fl = (fl + 1) & ~1
        g = math.gcd(self.out_rate, self.in_rate)
        if self.out_rate // g <= pc:                            # exact_rational
            pc = self.out_rate // g
        self.factor, self.phase_count, self.filter_length = factor, pc, fl
        self.center = (fl - 1) // 2
        i = np.arange(fl)
        bank = np.zeros((pc, fl), dtype=np.float64)
        for ph in range(pc):
            t = (i - self.center) - ph / pc
            x = np.pi * t * factor
            y = np.where(x == 0, 1.0, np.sin(x) / np.where(x == 0, 1.0, x))
            w = 2.0 * np.abs(t) / fl
            bank[ph] = y * _bessel_i0(beta * np.sqrt(np.maximum(1.0 - w * w, 0.0)))
        bank /= bank[0].sum()
        self.bank = bank.astype(np.float32)
        num, den = self.out_rate, self.in_rate * pc
        g = math.gcd(num, den)
        src_incr, dst_incr = num // g, den // g
        while dst_incr < (1 << 20) and src_incr < (1 << 20):
            dst_incr *= 2
            src_incr *= 2
        self.src_incr, self.dst_incr = src_incr, dst_incr

    # output k reads x[start_k : start_k + filter_length], start_k = floor(t_k) - center
    def positions(self, k: np.ndarray):
        tot = k.astype(np.int64) * self.dst_incr
        idx, frac = tot // self.src_incr, tot % self.src_incr
        return idx // self.phase_count, idx % self.phase_count, frac

    def output_count(self, n_in: int) -> tuple[int, int]:
        """(total outputs after flush, outputs a single regular call yields) for n_in input samples."""
        if self.passthrough:
            return n_in, n_in
        fl = self.filter_length
        k = np.arange(int(n_in * self.out_rate / self.in_rate) + 8, dtype=np.int64)
        start = self.positions(k)[0] - self.center
        n_main = int(np.count_nonzero(start + fl <= n_in))
        leftover = n_in - int(start[n_main])
        refl = (min(leftover, fl) + 1) // 2
        return int(np.count_nonzero(start + fl <= n_in + refl)), n_main

    def resample(self, x: np.ndarray) -> np.ndarray:
        """float32 in -> float64 out (the value libswresample rounds to float32 / int16)."""
        x = np.asarray(x, dtype=np.float32)
        if self.passthrough:
            return x.astype(np.float64)
        n = x.size
        fl, pc = self.filter_length, self.phase_count
        n_out, _ = self.output_count(n)
        samp, ph, frac = self.positions(np.arange(n_out, dtype=np.int64))
        pad = fl + 2
        xe = np.concatenate([x[1:pad + 1][::-1], x, x[::-1][:pad]]).astype(np.float64)
        taps = np.arange(fl)[None, :]
        base = samp - self.center + pad
        v1 = (xe[base[:, None] + taps] * self.bank[ph].astype(np.float64)).sum(1)
        wrap = ph == pc - 1
        v2 = (xe[(base + wrap)[:, None] + taps] * self.bank[np.where(wrap, 0, ph + 1)].astype(np.float64)).sum(1)
        return v1 + (v2 - v1) * frac / self.src_incr

    def resample_s16(self, x: np.ndarray) -> np.ndarray:
        y = self.resample(x).astype(np.float32)
        return np.clip(np.rint(y * np.float32(32768.0)), -32768, 32767).astype(np.int16)
