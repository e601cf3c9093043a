"""CPU-only: the polyphase overlap-save factorisation the kernel implements
(iq_to_audio_b200/plan.py) reproduces the reference's decimated channel samples."""
from __future__ import annotations

import numpy as np
import pytest

from iq_to_audio_b200 import plan as P
from iq_to_audio_b200 import processing as gp
from oracle import iq_oracle as orc
from tests import _cases


@pytest.mark.parametrize("m_fft", [256, 512, 1024])
def test_block_math_matches_reference_baseband(m_fft):
    g = _cases.load("case_a_nfm_2p5M")
    x = _cases.complex_input("case_a_nfm_2p5M")
    taps = orc.channel_taps(2.5e6, 12_500.0, 26)
    pl = P.build_plan(2.5e6, 26, [P.ChannelSpec(25e3, taps, int(g["mix_sign"]))], m_fft=m_fft)
    chunk = 65_536
    ph = [P.phase_table(pl.increments[0], chunk, (x.size + chunk - 1) // chunk)]
    s = P.emulate_block_math(pl, x, g["baseband"].size, ph, chunk)
    # 3.7e-8 == the reference's own complex64 mixer rounding (exact FIR vs reference 3.5e-8)
    assert np.abs(s[0] - g["baseband"]).max() < 1e-7


def test_multi_channel_plan_and_phase_table():
    m = _cases.manifest()["case_b_nfm_10M"]
    x = _cases.complex_input("case_b_nfm_10M")[:300_000]
    taps = orc.channel_taps(10e6, 12_500.0, 104)
    chans = [P.ChannelSpec(t["f_off"], taps, 1) for t in m["targets"][:2]]
    pl = P.build_plan(10e6, 104, chans, m_fft=512)
    chunk = m["chunk"]
    ph = [P.phase_table(w, chunk, 2) for w in pl.increments]
    n_out = orc.decimated_count(0, x.size, 104)
    s = P.emulate_block_math(pl, x, n_out, ph, chunk)
    for i in range(2):
        g = _cases.load(f"case_b_nfm_10M_t{i}")
        assert np.abs(s[i] - g["baseband"][:n_out]).max() < 1e-7
    # the phase table is the reference's recurrence, bit for bit
    st = orc.NcoState.for_offset(m["targets"][0]["f_off"], 10e6)
    for k in range(2):
        assert ph[0][k] == st.phase
        st.phase = orc.nco_advance(st.phase, st.increment, 1, chunk)


def test_slot_permutation_is_a_bijection():
    for m in P.SUPPORTED_M:
        perm = P.spectrum_slot_to_bin(m)
        assert sorted(perm) == list(range(m))


def test_host_helpers_match_oracle():
    for fs in (0.0, 250e3, 1e6, 2.5e6, 10e6, 20e6, 61.44e6):
        for req in (1, 65_536, 1_048_576, 8_000_000):
            assert gp.tune_chunk_size(fs, req) == orc.plan_chunk(fs, req)
    for fs, tgt in ((2.5e6, 96e3), (10e6, 96e3), (61.44e6, 96e3), (250e3, 96e3), (48e3, 96e3), (150e3, 96e3)):
        assert gp.channel_decimation(fs, tgt) == orc.plan_decimation(fs, tgt)
    for fs, bw, d in ((2.5e6, 12_500.0, 26), (10e6, 12_500.0, 104), (20e6, 2_800.0, 208), (250e3, 200_000.0, 3)):
        np.testing.assert_array_equal(gp.design_channel_filter(fs, bw, d), orc.channel_taps(fs, bw, d))
    with pytest.raises(ValueError):
        gp.design_channel_filter(1e6, -1.0, 10)


def test_int16_unpack_bit_trick_is_exact_for_every_value():
    """The channel bank turns an int16 pair into two floats without an integer-to-float conversion
    (csrc/channelizer2.cuh): flip the sign bit of each half (offset binary; for a negated Q complement the other
    15 bits as well), put each half under the exponent bytes 0x4B00 and subtract 8421376 (8421375 for the negated
    half).  Restated in numpy for all 65 536 values of both halves, all four IQ orders."""
    v = np.arange(-32768, 32768, dtype=np.int64)
    lo, hi = np.meshgrid(v[::257], v, indexing="ij")              # every high half against a spread of low halves
    lo = np.concatenate([lo.ravel(), v, v[::-1]])
    hi = np.concatenate([hi.ravel(), v[::-1], v])
    w = ((hi & 0xFFFF) << 16 | (lo & 0xFFFF)).astype(np.uint32)   # frame word: first int16 in the low half
    for swap in (0, 1):
        for q_neg in (0, 1):
            qmask = 0x7FFF if q_neg else 0x8000
            xmask = np.uint32((0x80000000 | qmask) if swap else ((qmask << 16) | 0x8000))
            x = w ^ xmask
            half_lo = (x & np.uint32(0xFFFF)) | np.uint32(0x4B000000)
            half_hi = (x >> np.uint32(16)) | np.uint32(0x4B000000)
            i_bits, q_bits = (half_hi, half_lo) if swap else (half_lo, half_hi)
            re = i_bits.view(np.float32) + np.float32(-8421376.0)
            im = q_bits.view(np.float32) + np.float32(-8421375.0 if q_neg else -8421376.0)
            want_i, want_q = (hi, lo) if swap else (lo, hi)
            np.testing.assert_array_equal(re, want_i.astype(np.float32))
            np.testing.assert_array_equal(im, (-want_q if q_neg else want_q).astype(np.float32))


@pytest.mark.parametrize("fs,d,bw", [(10e6, 104, 12_500.0), (2.4e6, 24, 12_500.0), (1.0e6, 12, 10_000.0),
                                     (0.768e6, 8, 12_500.0), (2.0e6, 20, 12_500.0), (1.6e6, 16, 12_500.0)])
def test_mirror_pair_form_equals_unpaired_form(fs, d, bw):
    """The mirror-pair factorisation of csrc/channelizer5.cuh (one real table entry and four multiply-adds per
    branch PAIR, aligned column groups, one spectrum carried from tile to tile) is an identity for firwin's
    symmetric taps: same channel samples as the unpaired model -- for r = 0 (no class-1 tiles), even and odd
    numbers of column groups per class (a middle group that mirrors itself) and single-tile classes."""
    taps = orc.channel_taps(fs, bw, d)
    rng = np.random.default_rng(d)
    n = 3 * 449 * d + 11
    x = (rng.normal(size=n) + 1j * rng.normal(size=n)) * 0.2
    chans = [P.ChannelSpec(0.13 * fs, taps, 1), P.ChannelSpec(-0.31 * fs, taps, -1)]
    pl = P.build_plan(fs, d, chans, m_fft=512)
    chunk = 1 << 16
    ph = [P.phase_table(w, chunk, n // chunk + 2) for w in pl.increments]
    n_out = orc.decimated_count(0, n, d)
    a = P.emulate_block_math(pl, x, n_out, ph, chunk)
    b = P.emulate_block_math_paired(pl, x, n_out, ph, chunk)
    # the only difference is the float32 rounding of the unpaired table (plan.g_table is complex64)
    assert np.abs(a - b).max() < 2e-7 * np.abs(a).max()
    tiles, aa, r = P.pair_tiles(len(taps), d)
    assert aa == -(-(len(taps) - 1) // d) and 0 <= r < d and r % 4 == 0
    groups = sorted([t.u for t in tiles] + [t.w for t in tiles if t.w != t.u])
    assert groups == list(range(d // 4))                             # every column group staged exactly once


def test_mirror_pair_form_rejects_what_it_cannot_pair():
    taps = orc.channel_taps(1.0e6, 10_000.0, 12).copy()
    taps[3] *= 1.0 + 1e-9
    pl = P.build_plan(1.0e6, 12, [P.ChannelSpec(1e5, taps, 1)], m_fft=512)
    with pytest.raises(ValueError):
        P.emulate_block_math_paired(pl, np.zeros(100, complex), 5, [np.zeros(2)], 1 << 16)
    with pytest.raises(ValueError):
        P.pair_tiles(1601, 26)                                       # D % 4 != 0: generation 4 takes it
    with pytest.raises(ValueError):
        P.pair_tiles(1311, 20)                                       # (ntaps - 1) % 4 != 0
