"""Seeded inputs of the spectrum golden vectors (shared by the generator and the tests)."""
from __future__ import annotations

import numpy as np

PSD_CASES = {
    # name: sample count vs nfft covers truncation (n > nfft), exact fit and zero padding (n < nfft)
    "trunc": dict(seed=11, n=6000, nfft=4096, fs=2.5e6),
    "pad": dict(seed=12, n=3000, nfft=4096, fs=1.0e6),
    "big": dict(seed=13, n=16384, nfft=16384, fs=10.0e6),
    "one": dict(seed=14, n=1, nfft=64, fs=48_000.0),
}

WATERFALL_CASES = {
    # ragged chunking, None / empty chunks, chunks shorter than nfft, slice-list halving (frames > max_slices)
    "ragged": dict(seed=21, nfft=1024, hop=None, max_slices=6, fs=2.5e6,
                   sizes=[700, 0, 500, 4096, 100, 100, 3000, 1024, 5, 2500]),
    "hop_big": dict(seed=22, nfft=512, hop=800, max_slices=400, fs=1.0e6, sizes=[3000, 1700, 512, 2100]),
    "odd_cap": dict(seed=23, nfft=256, hop=64, max_slices=7, fs=96_000.0, sizes=[4000, 333, 2000]),
    "large": dict(seed=24, nfft=16384, hop=None, max_slices=3, fs=10.0e6, sizes=[40000, 30000]),
}


def _signal(seed: int, n: int, fs: float) -> np.ndarray:
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64)
    x = 0.4 * np.exp(2j * np.pi * 0.1234 * t) + 0.05 * np.exp(-2j * np.pi * 0.31 * t + 1j)
    x += 1e-3 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    # quantise like an int16 capture so that the inputs are exactly representable in every format
    q = np.round(x.real * 32767.0) / 32768.0 + 1j * (np.round(x.imag * 32767.0) / 32768.0)
    return q.astype(np.complex64)


def psd_input(case: dict) -> np.ndarray:
    return _signal(case["seed"], case["n"], case["fs"])


def waterfall_chunks(case: dict) -> list:
    x = _signal(case["seed"], int(sum(case["sizes"])), case["fs"])
    out, pos = [], 0
    for i, n in enumerate(case["sizes"]):
        out.append(x[pos:pos + n])
        pos += n
        if i == 1:
            out.append(None)
    return out
