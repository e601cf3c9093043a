"""CPU-only models of two pieces of device arithmetic whose correctness rests on constants and index algebra:
the bank-conflict-free tile index of the float64 filter (csrc/precise_fft.cu: tix) and the Cody-Waite sine / cosine of
the bit-faithful mixer (csrc/sincos_cw.cuh).  The GPU tests check the kernels end to end; these pin the reasoning."""
from __future__ import annotations

import re
from fractions import Fraction
from pathlib import Path

import numpy as np

CSRC = Path(__file__).resolve().parents[1] / "iq_to_audio_b200" / "csrc"


def tix(row: int, c: int) -> int:
    return ((row << 2) | c) ^ (((row >> 1) & 3) | (((row >> 5) & 1) << 2))


def test_tile_index_is_a_bijection_and_conflict_free():
    """Element (row, column) -> double2 index.  A quarter-warp (8 threads x 16 bytes) is one shared-memory wavefront when
    its eight 16-byte units are distinct modulo 8 (128 bytes): true for the pass-1 stores, the pass-2 loads / stores
    and the multiply-accumulate reads, with the thread mappings of k_fir_fft64r (c = tid & 3, g = tid >> 2)."""
    src = (CSRC / "precise_fft.cu").read_text()
    assert "((row << 2) | c) ^ (((row >> 1) & 3) | (((row >> 5) & 1) << 2))" in src       # the model is the device's formula
    idx = {tix(r, c) for r in range(1024) for c in range(4)}
    assert idx == set(range(4096))                                                       # bijection onto the 64 KB tile

    def distinct(units):
        return len({u % 8 for u in units}) == 8

    for q in range(0, 128, 8):                                  # every quarter-warp of the 128-thread CTA
        lanes = [(t & 3, t >> 2) for t in range(q, q + 8)]      # (c, g)
        for k1 in range(32):                                    # pass-1 store of result k1: row k1 * 32 + g
            assert distinct([tix(k1 * 32 + g, c) for c, g in lanes])
        for i in range(32):                                     # pass-2 load / store of element i: row g * 32 + i
            assert distinct([tix(g * 32 + i, c) for c, g in lanes])
        for i in range(8):                                      # multiply-accumulate: row tid + 128 i, one column at a time
            for c in range(4):
                assert distinct([tix(t + 128 * i, c) for t in range(q, q + 8)])
    # the padded layout this replaced (row stride 5 units) needed two wavefronts for both transform passes
    assert not distinct([(g * 5 + c) for c, g in [(t & 3, t >> 2) for t in range(8)]])
    assert not distinct([((g * 32) * 5 + c) for c, g in [(t & 3, t >> 2) for t in range(8)]])


def _constants():
    text = (CSRC / "sincos_cw.cuh").read_text()
    nums = [float(x) for x in re.findall(r"-?\d\.\d{10,}e[+-]\d\d", text)]
    return text, nums


def sincos_cw_model(x: np.ndarray):
    """The device routine in numpy float64; the one fused multiply-add whose exactness the reduction relies on is done
    in rational arithmetic and rounded once, the polynomial steps use separate multiplies and adds (slightly less
    accurate than the device's fma chain)."""
    _, k = _constants()
    two_over_pi, p1, p1t = k[0], k[1], k[2]
    s6, s5, s4, s3, s2, s1 = k[3], k[4], k[5], k[6], k[7], k[8]
    c6, c5, c4, c3, c2, c1 = k[9], k[10], k[11], k[12], k[13], k[14]
    fn = np.rint(x * two_over_pi)
    r = np.array([float(Fraction(float(xi)) - Fraction(float(f)) * Fraction(p1)) for xi, f in zip(x, fn)])
    # x - fn * P1 is a multiple of 2^-32 below 2 in magnitude: exactly representable, so float() above rounds nothing
    assert all(Fraction(float(ri)) == Fraction(float(xi)) - Fraction(float(f)) * Fraction(p1) for ri, xi, f in zip(r[:64], x[:64], fn[:64]))
    w = fn * p1t
    y = r - w
    yt = (r - y) - w
    z = y * y
    v = z * y
    rs = s2 + z * (s3 + z * (s4 + z * (s5 + z * s6)))
    ks = y - ((z * (0.5 * yt - v * rs) - yt) - v * s1)
    rc = z * (c1 + z * (c2 + z * (c3 + z * (c4 + z * (c5 + z * c6)))))
    hz = 0.5 * z
    a = 1.0 - hz
    kc = a + (((1.0 - a) - hz) + (z * rc - y * yt))
    n = fn.astype(np.int64)
    s = np.where(n & 1, kc, ks)
    c = np.where(n & 1, ks, kc)
    return np.where(n & 2, -s, s), np.where((n + 1) & 2, -c, c)


def test_cody_waite_sincos_constants_and_quadrants():
    text, k = _constants()
    assert len(k) == 15, k                                    # 2/pi, P1, P1t, six sine and six cosine coefficients
    assert abs(k[0] - 2 / np.pi) < 1e-16 and abs((k[1] + k[2]) - np.pi / 2) < 1e-16
    assert float(k[1]).hex().endswith("00000p+0")             # P1 carries 33 bits: the low 20 bits of its mantissa are zero
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(-4.9e7, 4.9e7, 20_000), rng.uniform(-10.0, 10.0, 5_000), rng.uniform(2.5e7, 2.7e7, 5_000),
                        np.array([0.0, np.pi / 4, -np.pi / 4, 1e-9, 3e7 + 0.5])])
    s, c = sincos_cw_model(x)
    # numpy's sin / cos are within 1 ulp of the true values for arguments of this size (glibc); the model's plain
    # multiply-add chain adds about one more
    assert np.abs(s - np.sin(x)).max() < 4e-16 and np.abs(c - np.cos(x)).max() < 4e-16
    # and what the mixer uses -- the float32 roundings -- agree except where the value sits on a rounding boundary
    assert np.mean(s.astype(np.float32) == np.sin(x).astype(np.float32)) > 0.9999
    assert np.mean(c.astype(np.float32) == np.cos(x).astype(np.float32)) > 0.9999
