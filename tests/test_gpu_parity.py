"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the
same seeded inputs and against the golden vectors recorded from the reference.

Tolerances (BASELINE.json north_star):
  * sample counts per chunk, decimation phase, chunk boundaries: exact
  * float audio: max abs error <= 1e-4 full scale
  * complex64 channel samples: <= 1e-6 (float32 transform path vs the reference's complex128)
"""
from __future__ import annotations

import numpy as np
import pytest

from oracle import iq_oracle as orc
from tests import _cases

pytestmark = pytest.mark.gpu

AUDIO_TOL = 1e-4
BB_TOL = 1e-6


@pytest.fixture(scope="module")
def gpu():
    from iq_to_audio_b200 import _lib
    if _lib.device_count() < 1:
        pytest.fail("GPU tests need a CUDA device; the product path has no CPU fallback")
    import iq_to_audio_b200.decoders as gd
    import iq_to_audio_b200.processing as gp
    from iq_to_audio_b200.bank import ChannelBank, Target
    return dict(gp=gp, gd=gd, ChannelBank=ChannelBank, Target=Target, lib=_lib)


@pytest.fixture(scope="module")
def sv():
    return _cases.load("stage_vectors")


# ------------------------------------------------------------------------------- stage level
def test_gtable_matches_host_plan(gpu):
    from iq_to_audio_b200 import plan as P
    taps = orc.channel_taps(2.5e6, 12_500.0, 26)
    for m_fft in (512, 1024):
        pl = P.build_plan(2.5e6, 26, [P.ChannelSpec(25e3, taps, 1), P.ChannelSpec(-300e3, taps, -1)], m_fft=m_fft)
        with gpu["ChannelBank"](2.5e6, 26, [gpu["Target"](25e3, taps, 1), gpu["Target"](-300e3, taps, -1)],
                                fft_size=m_fft) as b:
            assert (b.fft_size, b.overlap_rows, b.rows_per_block) == (m_fft, pl.vd, pl.ld)
            assert np.abs(b.g_table() - pl.g_table).max() <= 1e-11


def test_mixer_matches_reference(gpu, sv):
    z = sv["mix_in"]
    osc = gpu["gp"].ComplexOscillator(123_456.7, 2.4e6)
    a = osc.mix(z[:17_000], -1)
    assert osc.phase == float(sv["mix_phase_a"])                 # phase carry: same float64 expression
    b = osc.mix(z[17_000:], -1)
    assert osc.phase == float(sv["mix_phase_b"])
    # LO from CUDA's float64 sincos rounded to float32: <= 1 ulp of the LO from numpy's
    assert np.abs(a - sv["mix_out_a"]).max() <= 3e-7 and np.abs(b - sv["mix_out_b"]).max() <= 3e-7


def test_mixer_with_unpack_formats(gpu):
    rng = np.random.default_rng(5)
    for codec, raw in (("pcm_s16le", rng.integers(-32768, 32767, 2 * 4097, dtype=np.int16)),
                       ("pcm_u8", rng.integers(0, 255, 2 * 4097, dtype=np.uint8)),
                       ("pcm_f32le", rng.normal(size=2 * 4097).astype(np.float32))):
        for order in orc.IQ_ORDERS:
            x = orc.order_iq(orc.unpack_interleaved(raw, codec), order)
            st = orc.NcoState(increment=-0.1234, phase=1.25)
            want = orc.nco_mix(st, x, 1)
            got = np.empty_like(x)
            lib = gpu["lib"]
            lib.check(lib.load().iq2a_unpack_mix(raw.ctypes.data, x.size, lib.CODEC_IDS[codec], lib.ORDER_IDS[order],
                                                 1.25, -0.1234, got.ctypes.data, 0))
            assert np.abs(got - want).max() <= 3e-7 * max(1.0, np.abs(want).max()), (codec, order)


def test_fir_ragged_calls_match_reference(gpu, sv):
    z = sv["mix_in"]
    fir = gpu["gp"].OverlapSaveFIR(sv["fir_taps"], 4096)
    out = np.concatenate([fir.process(z[:5_000]), fir.process(z[5_000:5_700]), fir.process(z[5_700:])])
    assert out.size == z.size
    assert np.abs(out - sv["fir_out"]).max() <= 1e-7
    np.testing.assert_array_equal(fir.state, sv["fir_state"])


def test_decimator_known_answers(gpu, sv):
    Decimator = gpu["gp"].Decimator
    d3 = Decimator(3)
    got = np.concatenate([d3.process(np.arange(9, dtype=np.complex64)), d3.process(np.arange(9, 18, dtype=np.complex64))])
    np.testing.assert_array_equal(got, np.arange(0, 18, 3, dtype=np.complex64))   # ref tests/test_processing.py:22-28
    z = sv["mix_in"]
    d7 = Decimator(7)
    got = np.concatenate([d7.process(z[:10]), d7.process(z[10:11]), d7.process(z[11:400])])
    np.testing.assert_array_equal(got, sv["dec7"])
    assert d7.offset == int(sv["dec7_offset"])
    assert Decimator(1).process(z[:5]) is not None and Decimator(1).process(z[:5]).size == 5


def test_discriminator_deemphasis_dc_agc(gpu, sv):
    gd = gpu["gd"]
    from iq_to_audio_b200.decoders.base import run_scan
    from iq_to_audio_b200.decoders.common import DCBlocker
    from iq_to_audio_b200.decoders.nfm import DeemphasisFilter, QuadratureDemod
    z = sv["mix_in"]
    qd = QuadratureDemod()
    a = np.concatenate([qd.process(z[:9_000]), qd.process(z[9_000:20_000])])
    assert a.size == 20_000 and np.abs(a - sv["disc"]).max() <= 1e-6           # atan2f vs numpy arctan2: ulps
    de = DeemphasisFilter(300.0, 96_153.846)
    assert de.alpha == float(sv["deemph_alpha"])
    y = np.concatenate([de.process(sv["disc"][:9_000]), de.process(sv["disc"][9_000:])])
    assert np.abs(y - sv["deemph"]).max() <= 1e-7
    assert abs(de.state - float(sv["deemph_state"])) <= 1e-12
    dc = DCBlocker()
    r = sv["dc_in"]
    y = np.concatenate([dc.process(r[:2_500]), dc.process(r[2_500:])])
    assert np.abs(y - sv["dc_out"]).max() <= 1e-6                               # float64 scan vs float32 loop
    # AGC alone on the reference's own DC-blocked audio (identical input): float64 affine scan vs the
    # reference's float32 sequential loop
    fresh = gpu["lib"].ChannelState.fresh
    g = np.concatenate([run_scan(2, 0.0, sv["dc_out"][:2_500], fresh()), run_scan(2, 0.0, sv["dc_out"][2_500:], fresh())])
    ref = sv["agc_out"]
    assert np.abs(g - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max())
    assert np.abs(np.clip(g, -0.99, 0.99) - np.clip(ref, -0.99, 0.99)).max() <= AUDIO_TOL


def test_decoder_plugins_match_oracle(gpu):
    rng = np.random.default_rng(11)
    n = 50_000
    t = np.arange(n)
    s = (0.3 * np.exp(1j * (0.02 * t + 2.0 * np.sin(0.004 * t))) * (1.0 + 0.5 * np.sin(0.01 * t))
         + 0.01 * (rng.normal(size=n) + 1j * rng.normal(size=n))).astype(np.complex64)
    for mode, agc in (("nfm", True), ("am", True), ("usb", False), ("lsb", False)):
        dec = gpu["gd"].create_decoder(mode, deemph_us=300.0, agc_enabled=agc)
        dec.setup(96_000.0)
        st = orc.DemodState.create(mode, 96_000.0, deemph_us=300.0, agc_enabled=agc)
        for lo, hi in ((0, 20_000), (20_000, 20_001), (20_001, n)):
            got, stats = dec.process(s[lo:hi])
            want, db = orc.demodulate(st, s[lo:hi])
            assert got.size == want.size
            assert np.abs(got - want).max() <= 2e-6, mode
            assert abs(stats.rms_dbfs - db) <= 1e-3
        assert dec.process(s[:0])[0].size == 0


def test_choose_mix_sign_known_answers(gpu):
    taps = orc.channel_taps(1e6, 12_500.0, 10)
    nn = np.arange(0, int(1e6 * 0.1))
    warm = np.exp(1j * 2.0 * np.pi * 12_500.0 * nn / 1e6).astype(np.complex64)
    choose = gpu["gp"].choose_mix_sign
    assert choose(warm, 1e6, 12_500.0, taps, 10) == 1            # ref tests/test_processing.py:31-40
    assert choose(np.conj(warm), 1e6, 12_500.0, taps, 10) == -1
    assert choose(warm[:0], 1e6, 12_500.0, taps, 10) == 1


# ------------------------------------------------------------------------------- fused path
def _targets(gpu, key, names):
    m = _cases.manifest()[key]
    fs = m["fs"]
    d, _ = orc.plan_decimation(fs, 96_000.0)
    gold = [_cases.load(n) for n in names]
    tg = []
    for t, g in zip(m["targets"], gold):
        bw = t.get("bw", 12_500.0 if t["mode"] == "nfm" else 2_800.0)
        tg.append(gpu["Target"](t["f_off"], orc.channel_taps(fs, bw, d), int(g["mix_sign"]), t["mode"], 300.0,
                                t.get("agc", True)))
    return m, fs, d, tg, gold


def _stream(gpu, key, names, fft_size=0, limit=None):
    m, fs, d, tg, gold = _targets(gpu, key, names)
    raw = _cases.raw_input(key).view(np.uint8)
    fb = orc.FRAME_BYTES[m["codec"]]
    chunk = m["chunk"]
    nfr = raw.size // fb if limit is None else min(raw.size // fb, limit)
    out = dict(audio=[], clipped=[], bb=[], counts=[], rms=[])
    with gpu["ChannelBank"](fs, d, tg, codec=m["codec"], iq_order=m["iq_order"], ref_chunk=chunk,
                            fft_size=fft_size) as bank:
        for s in range(0, nfr, chunk):
            r = bank.process_chunk(raw[s * fb:min(s + chunk, nfr) * fb], want_baseband=True)
            out["audio"].append(r.audio.copy()); out["clipped"].append(r.clipped.copy())
            out["bb"].append(r.baseband.copy()); out["counts"].append(r.count); out["rms"].append(r.rms_dbfs.copy())
        peaks = bank.peaks
        consumed = bank.get_state()[1]
    assert consumed == nfr
    cat = {k: np.concatenate(out[k], axis=1) for k in ("audio", "clipped", "bb")}
    return cat, out["counts"], np.asarray(out["rms"]), peaks, gold


def _check(cat, counts, rms, peaks, gold, audio_tol=AUDIO_TOL):
    for i, g in enumerate(gold):
        assert counts == list(g["counts"])                                    # exact chunk boundaries
        assert np.abs(cat["bb"][i] - g["baseband"]).max() <= BB_TOL
        assert np.abs(cat["audio"][i] - g["audio"]).max() <= audio_tol
        assert np.abs(cat["clipped"][i] - g["clipped"]).max() <= audio_tol
        assert np.abs(rms[:, i] - g["rms_dbfs"]).max() <= 1e-3
        assert abs(peaks[i] - float(g["peak"])) <= audio_tol * max(1.0, float(g["peak"]))


@pytest.mark.parametrize("fft_size", [512, 1024])
def test_stream_case_a_benchmark_shape(gpu, fft_size):
    _check(*_stream(gpu, "case_a_nfm_2p5M", ["case_a_nfm_2p5M"], fft_size))


@pytest.mark.parametrize("fft_size", [512, 1024])
def test_stream_case_b_five_nfm_targets(gpu, fft_size):
    _check(*_stream(gpu, "case_b_nfm_10M", [f"case_b_nfm_10M_t{i}" for i in range(5)], fft_size))


def test_stream_case_c_am_usb_lsb_with_agc(gpu):
    """cfg3 shape: AM + USB + LSB, AGC on.  The SSB channels run on the bit-faithful path (complex64
    mixer emulation, float64 decimating FIR, float32 sequential DC blocker + per-chunk AGC): the AGC
    output -- values up to ~58 before the clip -- matches the reference within 1e-4 of full scale."""
    names = ["case_c_20M_am", "case_c_20M_usb", "case_c_20M_lsb"]
    cat, counts, rms, peaks, gold = _stream(gpu, "case_c_20M_am_ssb", names)
    _check({k: v[:1] for k, v in cat.items()}, counts, rms[:, :1], peaks[:1], gold[:1])
    for i in (1, 2):
        g = gold[i]
        assert counts == list(g["counts"])
        assert np.mean(cat["bb"][i] == g["baseband"]) > 0.99999          # bit-identical channel samples
        assert np.abs(cat["clipped"][i] - g["clipped"]).max() <= AUDIO_TOL
        # the AGC output reaches ~58 before the clip: held to 1e-4 ABSOLUTE (measured: every sample is bit-identical)
        assert np.abs(cat["audio"][i] - g["audio"]).max() <= AUDIO_TOL
        assert abs(peaks[i] - float(g["peak"])) <= AUDIO_TOL
        assert np.abs(rms[:, i] - g["rms_dbfs"]).max() <= 1e-3


def test_resident_case_c_agc_chunks_in_parallel(gpu):
    """Same capture through the resident API: the four reference chunks are processed in parallel
    (speculative DC-blocker start + verification, AGC restart per chunk) and still match."""
    import torch
    names = ["case_c_20M_am", "case_c_20M_usb", "case_c_20M_lsb"]
    m, fs, d, tg, gold = _targets(gpu, "case_c_20M_am_ssb", names)
    raw = _cases.raw_input("case_c_20M_am_ssb")
    n = raw.size // 2
    d_raw = torch.from_numpy(raw.copy()).cuda()
    with gpu["ChannelBank"](fs, d, tg, ref_chunk=m["chunk"]) as bank:
        rows = bank.rows_in(0, n)
        d_audio = torch.zeros((3, rows), dtype=torch.float32, device="cuda")
        d_clip = torch.zeros((3, rows), dtype=torch.float32, device="cuda")
        k, rms = bank.process_resident(d_raw.data_ptr(), 0, n, 0, n, dev_audio=d_audio.data_ptr(),
                                       dev_clipped=d_clip.data_ptr(), out_stride=rows, want_rms=True)
        a, c = d_audio.cpu().numpy(), d_clip.cpu().numpy()
    for i, g in enumerate(gold):
        assert k == g["audio"].size
        assert np.abs(c[i] - g["clipped"]).max() <= AUDIO_TOL
        assert np.abs(a[i] - g["audio"]).max() <= AUDIO_TOL                   # absolute, also for the AGC's ~58 peaks
        assert np.abs(rms[i] - g["rms_dbfs"]).max() <= 1e-3


@pytest.mark.parametrize("name", ["case_d_pcm_u8_qi_nfm", "case_d_pcm_f32le_iq_inv_nfm", "case_d_pcm_s16le_qi_inv_usb"])
def test_stream_case_d_formats_and_iq_order(gpu, name):
    _check(*_stream(gpu, name, [name]))


def test_stream_case_e_preview_truncation(gpu):
    m = _cases.manifest()["case_e_truncated"]
    _cases.manifest()["case_e_truncated"].setdefault("fs", 2.5e6)
    m.setdefault("codec", "pcm_s16le"); m.setdefault("iq_order", "iq")
    m.setdefault("targets", [dict(f_off=25e3, bw=12_500.0, mode="nfm")])
    _check(*_stream(gpu, "case_e_truncated", ["case_e_truncated"], limit=m["max_input_samples"]))


@pytest.mark.parametrize("order", ["iq", "qi", "iq_inv", "qi_inv"])
def test_bulk_kernel_equals_first_generation_kernel(gpu, order, monkeypatch):
    """The mirror-pair kernel (generation 5), the unpaired TMA / packed-f32x2 kernel (generation 4) and the
    bounds-checked kernel (generation 1) implement the same arithmetic: same capture, ragged call sizes
    -> same channel samples.  A filter that is not symmetric cannot be paired and lands on generation 4."""
    fs, d = 10e6, 104
    taps = orc.channel_taps(fs, 12_500.0, d)
    rng = np.random.default_rng(17)
    n = 700_001
    raw = rng.integers(-20_000, 20_000, 2 * n, dtype=np.int16)
    # both ends of the int16 range on both axes: the bit-level unpack (offset binary, Q negation by
    # complementing) has to agree with generation 1's plain conversion there too
    raw[1000:1008] = [-32768, -32768, 32767, 32767, -32768, 32767, 32767, -32768]
    raw[2 * 350_000:2 * 350_000 + 4] = [-32768, 0, 0, -32768]
    T = gpu["Target"]
    sizes = [300_000, 3, 101, 104, 250_000, n]

    def run(taps):
        tg = [T(1.0e6, taps, 1, "iq"), T(-2.2e6, taps, -1, "iq"), T(3.05e6, taps, 1, "iq")]
        with gpu["ChannelBank"](fs, d, tg, iq_order=order, ref_chunk=1 << 18, fft_size=512) as bank:
            gen, pos, parts = bank.kernel_generation, 0, []
            for sz in sizes:
                e = min(n, pos + sz)
                if e > pos:
                    parts.append(bank.process_chunk(raw[2 * pos:2 * e], want_baseband=True).baseband.copy())
                pos = e
            return gen, np.concatenate(parts, axis=1)
    gen5, bb5 = run(taps)
    skew = taps.copy()
    skew[7] *= 1.0 + 1e-9                                   # no longer exactly symmetric
    gen4s, _ = run(skew)
    monkeypatch.setenv("IQ2A_CHANNELIZER", "v4")
    gen4, bb4 = run(taps)
    monkeypatch.setenv("IQ2A_CHANNELIZER", "v1")
    gen1, bb1 = run(taps)
    assert (gen5, gen4s, gen4, gen1) == (5, 4, 4, 1)
    assert bb1.shape == bb4.shape == bb5.shape == (3, orc.decimated_count(0, n, d))
    assert np.abs(bb1 - bb5).max() <= 1e-6 * 20_000 / 32768 * 4
    assert np.abs(bb1 - bb4).max() <= 1e-6 * 20_000 / 32768 * 4


@pytest.mark.parametrize("fs,d,bw", [(2.4e6, 24, 12_500.0), (1.0e6, 12, 10_000.0), (20e6, 208, 10_000.0),
                                     (61.44e6, 640, 12_500.0), (0.768e6, 8, 12_500.0)])
def test_mirror_pair_kernel_geometries(gpu, fs, d, bw, monkeypatch):
    """Generation 5 over the pair geometries of other rates: r = 0 (one class is a single self-paired branch),
    both classes ragged, long rotations (A = 158) -- against generation 1 on the same random capture."""
    taps = orc.channel_taps(fs, bw, d)
    rng = np.random.default_rng(d)
    n = 40 * 449 * d // 8 + 77
    raw = rng.integers(-20_000, 20_000, 2 * n, dtype=np.int16)
    T = gpu["Target"]
    tg = [T(0.13 * fs, taps, 1, "iq"), T(-0.31 * fs, taps, -1, "iq")]

    def run():
        with gpu["ChannelBank"](fs, d, tg, ref_chunk=1 << 18, fft_size=512) as bank:
            return bank.kernel_generation, bank.process_chunk(raw, want_baseband=True).baseband.copy()
    gen5, bb5 = run()
    monkeypatch.setenv("IQ2A_CHANNELIZER", "v1")
    gen1, bb1 = run()
    assert (gen5, gen1) == (5, 1)
    assert np.abs(bb1 - bb5).max() <= 1e-6 * 20_000 / 32768 * 4


def test_pipelined_stream_equals_synchronous_chunks(gpu):
    """ChannelBank.stream (two chunks in flight: H2D of chunk k+1 overlaps the kernels of chunk k)
    returns exactly what the synchronous per-chunk call returns."""
    m, fs, d, tg, gold = _targets(gpu, "case_b_nfm_10M", [f"case_b_nfm_10M_t{i}" for i in range(5)])
    raw = _cases.raw_input("case_b_nfm_10M")
    n = raw.size // 2
    chunk = 100_000
    views = [raw[2 * s:2 * min(s + chunk, n)] for s in range(0, n, chunk)]
    with gpu["ChannelBank"](fs, d, tg, ref_chunk=chunk) as bank:
        sync = [bank.process_chunk(v, want_baseband=True) for v in views]
        sync = [(r.audio.copy(), r.clipped.copy(), r.baseband.copy(), r.rms_dbfs.copy()) for r in sync]
        bank.reset()
        piped = [(r.audio.copy(), r.clipped.copy(), r.baseband.copy(), r.rms_dbfs.copy())
                 for r in bank.stream(views, want_baseband=True)]
        with pytest.raises(RuntimeError, match="no chunk in flight"):
            gpu["lib"].check(bank._lib.iq2a_bank_collect_chunk(bank._h, None, None, None, 0, None, None))
    assert len(sync) == len(piped) == len(views)
    for a, b in zip(sync, piped):
        for x, y in zip(a[:3], b[:3]):
            np.testing.assert_array_equal(x, y)
        np.testing.assert_allclose(a[3], b[3], rtol=0, atol=1e-9)   # float64 atomics: summation order varies


@pytest.mark.parametrize("fs,mode", [(120_000.0, "nfm"), (96_000.0, "am"), (130_000.0, "usb")])
def test_low_rate_capture_direct_mode(gpu, fs, mode):
    """Sample rates up to ~1.5 x fs_ch give D = 1 and, with the 1025+ taps the reference always designs, more history
    rows than any transform size holds.  The reference handles every rate (OverlapSaveFIR has no limit;
    Decimator(1) returns its input, processing.py:354-356); the bank switches to direct mode -- the float64 mixer and
    direct-form filter of the bit-faithful path for every channel -- instead of rejecting the capture."""
    d, fs_ch = orc.plan_decimation(fs, 96_000.0)
    assert d == 1
    n = 150_003
    kind = {"nfm": "fm", "am": "am", "usb": "usb"}[mode]
    off = 0.17 * fs
    car = dict(offset=off, amp=0.3, kind=kind, tone=800.0)
    if kind == "fm":
        car["dev"] = 2500.0
    raw = orc.to_s16(orc.multi_carrier_capture(fs, n, [car], noise_std=0.01, seed=3))
    bw = 12_500.0 if mode == "nfm" else 6_000.0
    taps = orc.channel_taps(fs, bw, d)
    assert len(taps) >= 1025
    chunk = 1 << 16
    T = gpu["Target"]
    with gpu["ChannelBank"](fs, d, [T(off, taps, 1, mode, 300.0, True)], ref_chunk=chunk) as bank:
        assert bank.kernel_generation == 0
        parts = [bank.process_chunk(raw[2 * s:2 * min(s + chunk, n)], want_baseband=True) for s in range(0, n, chunk)]
    bb = np.concatenate([p.baseband for p in parts], axis=1)[0]
    clipped = np.concatenate([p.clipped for p in parts], axis=1)[0]
    x = orc.order_iq(orc.unpack_interleaved(raw, "pcm_s16le"), "iq")
    plan = orc.TargetPlan(sample_rate=fs, freq_offset=off, bandwidth=bw, mode=mode, agc_enabled=True, mix_sign=1)
    want = orc.run_target(x, plan, chunk)
    assert bb.size == want.baseband.size == n
    assert np.mean(bb == want.baseband) > 0.9999                   # float64 direct sum vs the reference's complex128 FFT
    assert np.abs(bb - want.baseband).max() <= BB_TOL
    skip = 0
    if mode == "nfm" and abs(want.baseband[0]) < 1e-12:
        # h[0] of this filter is ~1e-20 (the window edge falls on a zero of the sinc), so the reference's s[0] is pure
        # transform round-off (|s[0]| ~ 1e-16 where the exact value is 4e-21) and its first discriminator output the
        # angle of that noise; the de-emphasis recurrence carries the arbitrary value for ~100 rows (0.97^n)
        assert abs(bb[0]) < 1e-18
        skip = 400
    assert np.abs(clipped[skip:] - want.clipped[skip:]).max() <= AUDIO_TOL


def test_many_channel_form_equals_fused_kernel(gpu, monkeypatch):
    """25+ channels of one filter take the many-channel form (csrc/channelizer5s.cuh: forward transforms once per
    wave of block sets, then a multiply-accumulate kernel per group of <= 4 channels).  Same arithmetic in the same
    order as the fused kernel: identical channel samples, ragged calls included."""
    fs, d = 10e6, 104
    taps = orc.channel_taps(fs, 12_500.0, d)
    rng = np.random.default_rng(29)
    n = 900_011
    raw = rng.integers(-20_000, 20_000, 2 * n, dtype=np.int16)
    T = gpu["Target"]
    tg = [T((-4.7 + 0.35 * i) * 1e6, taps, 1 if i % 3 else -1, "iq") for i in range(27)]
    sizes = [400_000, 7, 300_000, n]

    def run():
        with gpu["ChannelBank"](fs, d, tg, ref_chunk=1 << 18, fft_size=512) as bank:
            pos, parts, launches = 0, [], bank.launches
            for sz in sizes:
                e = min(n, pos + sz)
                if e > pos:
                    parts.append(bank.process_chunk(raw[2 * pos:2 * e], want_baseband=True).baseband.copy())
                pos = e
            return np.concatenate(parts, axis=1), bank.launches - launches
    many, l_many = run()
    monkeypatch.setenv("IQ2A_MANY", "0")
    fused, l_fused = run()
    assert many.shape == fused.shape == (27, orc.decimated_count(0, n, d))
    assert np.array_equal(many, fused)
    assert l_many != l_fused                                       # the two forms really are different launch sequences


def test_batched_stream_keeps_reference_chunk_semantics(gpu):
    """ChannelBank.stream(..., chunk_frames=): several reference chunks per GPU call.  Phase wraps, AGC restarts and
    DecoderStats stay per reference chunk, so the result equals the one-chunk-per-call run (the block grid of the
    transforms moves with the call, hence float32 rounding and not bit equality on the fast path).  USB with AGC on
    is on the bit-faithful path: both ways of calling must reproduce the ORACLE bit for bit -- which needs the
    filter history of a call to be mixed with the phase of the chunk it belonged to (the reference mixed it then),
    not with a phase extrapolated backwards from the current chunk."""
    fs = 10e6
    d, fs_ch = orc.plan_decimation(fs, 96_000.0)
    n = 1_900_037
    carriers = [dict(offset=1.0e6, amp=0.2, kind="fm", tone=700.0, dev=2500.0), dict(offset=-2.2e6, amp=0.2, kind="am", tone=800.0),
                dict(offset=3.05e6, amp=0.2, kind="usb", tone=900.0)]
    raw = orc.to_s16(orc.multi_carrier_capture(fs, n, carriers, noise_std=0.01, seed=11))
    T = gpu["Target"]
    tg = [T(1.0e6, orc.channel_taps(fs, 12_500.0, d), 1, "nfm"), T(-2.2e6, orc.channel_taps(fs, 10_000.0, d), 1, "am"),
          T(3.05e6, orc.channel_taps(fs, 2_800.0, d), -1, "usb", 300.0, True)]
    chunk = 1 << 18
    with gpu["ChannelBank"](fs, d, tg, ref_chunk=chunk) as bank:
        one = [bank.process_chunk(raw[2 * s:2 * min(s + chunk, n)]) for s in range(0, n, chunk)]
    a1 = np.concatenate([r.audio for r in one], axis=1)
    c1 = np.concatenate([r.clipped for r in one], axis=1)
    rms1 = np.stack([r.rms_dbfs for r in one], axis=1)
    cnt1 = np.array([r.count for r in one])
    with gpu["ChannelBank"](fs, d, tg, ref_chunk=chunk) as bank:
        k = 3                                                    # 3 reference chunks per call, a ragged last call
        views = [raw[2 * s:2 * min(s + k * chunk, n)] for s in range(0, n, k * chunk)]
        many = list(bank.stream(views, chunk_frames=chunk))
    a2 = np.concatenate([r.audio for r in many], axis=1)
    c2 = np.concatenate([r.clipped for r in many], axis=1)
    rms2 = np.concatenate([r.rms_dbfs for r in many], axis=1)
    cnt2 = np.concatenate([r.window_counts for r in many])
    assert np.array_equal(cnt1, cnt2) and a1.shape == a2.shape and rms1.shape == rms2.shape
    assert np.abs(a1[:2] - a2[:2]).max() <= 2e-6 and np.abs(c1[:2] - c2[:2]).max() <= 2e-6
    assert np.abs(rms1 - rms2).max() <= 1e-4
    x = orc.order_iq(orc.unpack_interleaved(raw, "pcm_s16le"), "iq")
    plan = orc.TargetPlan(sample_rate=fs, freq_offset=3.05e6, bandwidth=2_800.0, mode="usb", agc_enabled=True, mix_sign=-1)
    want = orc.run_target(x, plan, chunk)
    for audio, clipped in ((a1[2], c1[2]), (a2[2], c2[2])):
        assert np.array_equal(audio, want.audio) and np.array_equal(clipped, want.clipped)


def test_chunk_split_invariance_and_state_roundtrip(gpu):
    """Feeding the same capture in different call sizes gives the same audio (NFM has no per-chunk
    semantics); get_state/set_state moves the carried decoder state between banks."""
    m, fs, d, tg, gold = _targets(gpu, "case_a_nfm_2p5M", ["case_a_nfm_2p5M"])
    raw = _cases.raw_input("case_a_nfm_2p5M")
    n = raw.size // 2
    outs = []
    for sizes in ([n], [1, 25, 26, 27, 100_003, n]):
        with gpu["ChannelBank"](fs, d, tg, ref_chunk=65_536) as bank:
            pos, parts = 0, []
            for sz in sizes:
                e = min(n, pos + sz)
                if e > pos:
                    parts.append(bank.process_chunk(raw[2 * pos:2 * e]).audio[0].copy())
                pos = e
            if pos < n:
                parts.append(bank.process_chunk(raw[2 * pos:]).audio[0].copy())
            outs.append(np.concatenate(parts))
    assert outs[0].size == outs[1].size == gold[0]["audio"].size
    assert np.abs(outs[0] - outs[1]).max() <= 2e-6
    assert np.abs(outs[0] - gold[0]["audio"]).max() <= AUDIO_TOL


def test_resident_whole_capture_and_time_shards(gpu):
    """The resident API on the whole capture, then as two time shards (halo + 600-row warm-up):
    both reproduce the reference's single-stream audio."""
    import torch
    key = "case_b_nfm_10M"
    names = [f"case_b_nfm_10M_t{i}" for i in range(5)]
    m, fs, d, tg, gold = _targets(gpu, key, names)
    raw = _cases.raw_input(key)
    n = raw.size // 2
    d_raw = torch.from_numpy(raw.copy()).cuda()
    chunk = m["chunk"]
    with gpu["ChannelBank"](fs, d, tg, ref_chunk=chunk) as bank:
        rows = bank.rows_in(0, n)
        d_audio = torch.zeros((5, rows), dtype=torch.float32, device="cuda")
        d_clip = torch.zeros((5, rows), dtype=torch.float32, device="cuda")
        d_bb = torch.zeros((5, rows), dtype=torch.complex64, device="cuda")
        k, rms = bank.process_resident(d_raw.data_ptr(), 0, n, 0, n, dev_audio=d_audio.data_ptr(),
                                       dev_clipped=d_clip.data_ptr(), dev_baseband=d_bb.data_ptr(),
                                       out_stride=rows, want_rms=True)
        assert k == rows == gold[0]["audio"].size
        a, c, b = d_audio.cpu().numpy(), d_clip.cpu().numpy(), d_bb.cpu().numpy()
        for i, g in enumerate(gold):
            assert np.abs(a[i] - g["audio"]).max() <= AUDIO_TOL
            assert np.abs(c[i] - g["clipped"]).max() <= AUDIO_TOL
            assert np.abs(b[i] - g["baseband"]).max() <= BB_TOL
            assert np.abs(rms[i] - g["rms_dbfs"]).max() <= 1e-3
        peaks = bank.peaks
        for i, g in enumerate(gold):
            assert abs(peaks[i] - float(g["peak"])) <= 1e-6
        # two shards, the second starts at a reference-chunk boundary with its own halo
        half = 2 * chunk
        r0 = bank.rows_in(0, half)
        bank.process_resident(d_raw.data_ptr(), 0, n, 0, half, dev_audio=d_audio.data_ptr(), out_stride=rows)
        a0 = d_audio.cpu().numpy()[:, :r0].copy()
        first = half - bank.halo - 601 * d
        view = d_raw[2 * first:]
        bank.process_resident(view.data_ptr(), first, n - first, half, n, warmup_rows=600,
                              dev_audio=d_audio.data_ptr(), out_stride=rows)
        a1 = d_audio.cpu().numpy()[:, :rows - r0]
        both = np.concatenate([a0, a1], axis=1)
        for i, g in enumerate(gold):
            assert np.abs(both[i] - g["audio"]).max() <= AUDIO_TOL
        # missing halo is an error, not a silent zero-fill
        with pytest.raises(ValueError, match="history"):
            bank.process_resident(d_raw[2 * half:].data_ptr(), half, n - half, half, n, warmup_rows=600,
                                  dev_audio=d_audio.data_ptr(), out_stride=rows)


def test_linearity_and_channel_independence_at_scale(gpu):
    """Size-independent properties on a larger capture (32 Mi samples, device-generated):
    (1) a channel's output does not depend on which other channels share the bank;
    (2) channel samples are linear in the input: bank(a) + bank(b) == bank(a+b) for complex64 input."""
    import torch
    fs, d = 10e6, 104
    taps = orc.channel_taps(fs, 12_500.0, d)
    n = 1 << 22
    rng = np.random.default_rng(3)
    t = np.arange(n) / fs
    a = (0.2 * np.exp(2j * np.pi * 400e3 * t) + 0.02 * (rng.normal(size=n) + 1j * rng.normal(size=n))).astype(np.complex64)
    b = (0.1 * np.exp(2j * np.pi * (400e3 + 700.0) * t)).astype(np.complex64)
    T = gpu["Target"]

    def run(x, targets):
        with gpu["ChannelBank"](fs, d, targets, codec="complex64", ref_chunk=1 << 20) as bank:
            return bank.process_chunk(x, want_baseband=True).baseband
    both = run(a, [T(400e3, taps, 1, "iq"), T(-2.5e6, taps, 1, "iq"), T(3.3e6, taps, -1, "iq")])
    alone = run(a, [T(400e3, taps, 1, "iq")])
    assert np.abs(both[0] - alone[0]).max() <= 1e-6
    sa, sb, sab = alone[0], run(b, [T(400e3, taps, 1, "iq")])[0], run((a + b).astype(np.complex64), [T(400e3, taps, 1, "iq")])[0]
    assert np.abs(sa + sb - sab).max() <= 2e-6


def test_cfg4_cfg5_shapes_against_oracle(gpu):
    """61.44 MS/s (D = 640, 32 769 taps): cfg4's five NFM targets and a cfg5-style 16-channel bank
    (three launch groups) on a short capture, against the CPU oracle run per target."""
    fs = 61.44e6
    d, fs_ch = orc.plan_decimation(fs, 96_000.0)
    assert d == 640
    taps = orc.channel_taps(fs, 12_500.0, d)
    assert len(taps) == 32_769
    n = 1_500_000
    offs = [(-28.0 + 3.7 * i) * 1e6 for i in range(16)]
    carriers = [dict(offset=o, amp=0.05, kind="fm", tone=600.0 + 90.0 * i, dev=2500.0) for i, o in enumerate(offs)]
    raw = orc.to_s16(orc.multi_carrier_capture(fs, n, carriers, noise_std=0.01, seed=5))
    x = orc.order_iq(orc.unpack_interleaved(raw, "pcm_s16le"), "iq")
    chunk = 1 << 20
    T = gpu["Target"]
    with gpu["ChannelBank"](fs, d, [T(o, taps, 1, "nfm") for o in offs], ref_chunk=chunk) as bank:
        assert bank.kernel_generation == 5 and bank.fft_size == 512
        parts = [bank.process_chunk(raw[2 * s:2 * min(s + chunk, n)], want_baseband=True) for s in range(0, n, chunk)]
        audio = np.concatenate([p.audio for p in parts], axis=1)
        bb = np.concatenate([p.baseband for p in parts], axis=1)
    for i in (0, 4, 7, 15):                                   # one channel from each launch group + the last
        plan = orc.TargetPlan(sample_rate=fs, freq_offset=offs[i], mix_sign=1)
        want = orc.run_target(x, plan, chunk)
        assert audio.shape[1] == want.audio.size == orc.decimated_count(0, n, d)
        assert np.abs(bb[i] - want.baseband).max() <= BB_TOL
        assert np.abs(audio[i] - want.audio).max() <= AUDIO_TOL


def test_cfg5_256_channels(gpu):
    """The widest bank of the sweep configuration: 256 NFM channels at 61.44 MS/s in one pass (32 launch groups
    of 8 sharing the capture in L2).  Three channels against the oracle, and the lifted CLI cap's bound."""
    fs = 61.44e6
    d, _ = orc.plan_decimation(fs, 96_000.0)
    taps = orc.channel_taps(fs, 12_500.0, d)
    n = 1_200_000
    offs = [(-29.0 + 58.0 * i / 255.0) * 1e6 for i in range(256)]
    pick = (0, 131, 255)
    carriers = [dict(offset=offs[i], amp=0.2, kind="fm", tone=500.0 + 150.0 * j, dev=2500.0) for j, i in enumerate(pick)]
    raw = orc.to_s16(orc.multi_carrier_capture(fs, n, carriers, noise_std=0.01, seed=9))
    x = orc.order_iq(orc.unpack_interleaved(raw, "pcm_s16le"), "iq")
    chunk = 1 << 20
    T = gpu["Target"]
    with gpu["ChannelBank"](fs, d, [T(o, taps, 1, "nfm") for o in offs], ref_chunk=chunk) as bank:
        assert bank.n_channels == 256
        parts = [bank.process_chunk(raw[2 * s:2 * min(s + chunk, n)], want_baseband=True) for s in range(0, n, chunk)]
        audio = np.concatenate([p.audio for p in parts], axis=1)
        bb = np.concatenate([p.baseband for p in parts], axis=1)
    assert audio.shape == (256, orc.decimated_count(0, n, d))
    for i in pick:
        want = orc.run_target(x, orc.TargetPlan(sample_rate=fs, freq_offset=offs[i], mix_sign=1), chunk)
        assert np.abs(bb[i] - want.baseband).max() <= BB_TOL
        assert np.abs(audio[i] - want.audio).max() <= AUDIO_TOL
    with pytest.raises(ValueError, match="n_channels"):
        gpu["ChannelBank"](fs, d, [T(o, taps, 1, "nfm") for o in offs + [1.0]], ref_chunk=chunk)


@pytest.mark.parametrize("modes", [("nfm", "nfm", "nfm"), ("am", "usb", "nfm"), ("lsb", "am", "am")])
def test_single_pass_tail_equals_three_kernel_scan(gpu, modes, monkeypatch):
    """The single-pass tail (overlapped history per CTA, no inter-CTA carry) and the reduce / carry / apply scan
    compute the same recurrences: same audio, clipped audio, peaks, statistics and carried state, for ragged
    streaming calls (state carried across calls, calls shorter than the history window) and AGC off."""
    fs, d = 10e6, 104
    taps = orc.channel_taps(fs, 12_500.0, d)
    n = 3_000_001
    carriers = [dict(offset=o, amp=0.2, kind=k, tone=700.0 + 100 * i, dev=2500.0)
                for i, (o, k) in enumerate([(1.0e6, "fm"), (-2.2e6, "am"), (3.05e6, "usb")])]
    raw = orc.to_s16(orc.multi_carrier_capture(fs, n, carriers, noise_std=0.01, seed=4))
    T = gpu["Target"]
    tg = [T(o, taps, 1, m, 300.0, False) for o, m in zip((1.0e6, -2.2e6, 3.05e6), modes)]
    sizes = [1_500_000, 5, 300, 20_000, 104 * 1024, n]

    def run():
        with gpu["ChannelBank"](fs, d, tg, ref_chunk=1 << 20) as bank:
            pos, audio, clip, rms = 0, [], [], []
            for sz in sizes:
                e = min(n, pos + sz)
                if e > pos:
                    r = bank.process_chunk(raw[2 * pos:2 * e])
                    audio.append(r.audio.copy()); clip.append(r.clipped.copy()); rms.append(np.array(r.rms_dbfs))
                pos = e
            st, _ = bank.get_state()
            return np.concatenate(audio, axis=1), np.concatenate(clip, axis=1), np.array(rms), bank.peaks, st, bank.launches
    a2, c2, r2, p2, s2, l2 = run()
    monkeypatch.setenv("IQ2A_TAIL", "v1")
    a1, c1, r1, p1, s1, l1 = run()
    assert l2 < l1                                             # 2 tail launches per call instead of 5
    assert a1.shape == a2.shape == (3, orc.decimated_count(0, n, d))
    assert np.abs(a1 - a2).max() <= 2e-7 and np.abs(c1 - c2).max() <= 2e-7
    assert np.abs(r1 - r2).max() <= 1e-6 and np.abs(np.array(p1) - np.array(p2)).max() <= 2e-7
    for x, y in zip(s1, s2):
        for key in x:
            assert abs(x[key] - y[key]) <= 1e-6 * max(1.0, abs(x[key])), key


@pytest.mark.parametrize("d", [26, 105, 7])
def test_cp_async_staging_equals_first_generation_kernel(gpu, d, monkeypatch):
    """Decimations whose row pitch is not a multiple of 16 bytes (cfg1: D = 26) cannot use the tensor map; the same
    kernel with cp.async staging takes them, ragged call sizes included, and agrees with generation 1."""
    fs = 96_000.0 * d
    taps = orc.channel_taps(fs, 12_500.0, d)
    rng = np.random.default_rng(23)
    n = 400_003
    raw = rng.integers(-20_000, 20_000, 2 * n, dtype=np.int16)
    T = gpu["Target"]
    tg = [T(0.11 * fs, taps, 1, "iq"), T(-0.23 * fs, taps, -1, "iq"), T(0.31 * fs, taps, 1, "iq")]
    sizes = [150_000, 1, 2 * d + 1, 100_000, n]

    def run():
        with gpu["ChannelBank"](fs, d, tg, iq_order="qi", ref_chunk=1 << 17, fft_size=512) as bank:
            gen, pos, parts = bank.kernel_generation, 0, []
            for sz in sizes:
                e = min(n, pos + sz)
                if e > pos:
                    parts.append(bank.process_chunk(raw[2 * pos:2 * e], want_baseband=True).baseband.copy())
                pos = e
            return gen, np.concatenate(parts, axis=1)
    gen4, bb4 = run()
    monkeypatch.setenv("IQ2A_CHANNELIZER", "v1")
    gen1, bb1 = run()
    assert (gen4, gen1) == (4, 1)
    assert bb1.shape == bb4.shape == (3, orc.decimated_count(0, n, d))
    assert np.abs(bb1 - bb4).max() <= 1e-6 * 20_000 / 32768 * 4


def test_resident_unaligned_pointer_uses_cp_async(gpu):
    """process_resident on a device pointer that is only 4-byte aligned: the tensor-map path is not applicable,
    the cp.async path gives the same rows as the aligned call."""
    import torch
    fs, d = 10e6, 104
    taps = orc.channel_taps(fs, 12_500.0, d)
    n = 600_000
    rng = np.random.default_rng(5)
    raw = rng.integers(-15_000, 15_000, 2 * n + 2, dtype=np.int16)
    T = gpu["Target"]
    buf = torch.from_numpy(np.concatenate([np.zeros(2, np.int16), raw])).cuda()       # payload starts 4 bytes in
    aligned = torch.from_numpy(raw.copy()).cuda()
    with gpu["ChannelBank"](fs, d, [T(1.2e6, taps, 1, "nfm"), T(-3.0e6, taps, 1, "am")], ref_chunk=1 << 18) as bank:
        rows = bank.rows_in(0, n)
        out = torch.zeros((2, 2, rows), dtype=torch.float32, device="cuda")
        bank.process_resident(aligned.data_ptr(), 0, n, 0, n, dev_audio=out[0].data_ptr(), out_stride=rows)
        bank.reset()
        assert (buf.data_ptr() + 4) % 16 == 4
        bank.process_resident(buf.data_ptr() + 4, 0, n, 0, n, dev_audio=out[1].data_ptr(), out_stride=rows)
    o = out.cpu().numpy()
    assert np.abs(o[0] - o[1]).max() <= 2e-6


def test_float64_scan_agc_path(gpu, monkeypatch):
    """IQ2A_PRECISE_SSB=0 puts SSB+AGC channels on the float64-scan AGC (DC-blocker scan, then the data-dependent
    gain recurrence as a second affine scan with restarts at the reference chunk boundaries).  The reference's AGC
    is ill-conditioned at audio zero crossings, so this path is not held to 1e-4 everywhere;
    it must agree with the reference on the channel samples, on the chunk structure, and on the audio almost
    everywhere, and the AM channel next to it must stay within tolerance."""
    monkeypatch.setenv("IQ2A_PRECISE_SSB", "0")
    names = ["case_c_20M_am", "case_c_20M_usb", "case_c_20M_lsb"]
    cat, counts, rms, peaks, gold = _stream(gpu, "case_c_20M_am_ssb", names)
    _check({k: v[:1] for k, v in cat.items()}, counts, rms[:, :1], peaks[:1], gold[:1])
    for i in (1, 2):
        g = gold[i]
        assert counts == list(g["counts"])
        assert np.abs(cat["bb"][i] - g["baseband"]).max() <= BB_TOL
        err = np.abs(cat["clipped"][i] - g["clipped"])
        assert np.isfinite(cat["audio"][i]).all()
        assert np.median(err) <= 1e-4 and np.mean(err <= 1e-2) >= 0.9
        assert np.abs(rms[:, i] - g["rms_dbfs"]).max() <= 0.5


def test_edge_inputs_empty_tiny_and_sub_row_calls(gpu):
    """Empty calls, calls shorter than one decimated row, a capture shorter than the filter: counts follow the
    reference's rule (rows = multiples of D inside the call) and a later normal call is unaffected."""
    m, fs, d, tg, gold = _targets(gpu, "case_a_nfm_2p5M", ["case_a_nfm_2p5M"])
    raw = _cases.raw_input("case_a_nfm_2p5M")
    g = gold[0]
    with gpu["ChannelBank"](fs, d, tg, ref_chunk=m["chunk"]) as bank:
        r = bank.process_chunk(raw[:0])
        assert r.count == 0 and r.audio.shape == (1, 0)
        pos, audio = 0, []
        for sz in (1, 0, d - 2, 1, 1, 5, 3 * d, 2):                 # row boundaries crossed one sample at a time
            r = bank.process_chunk(raw[2 * pos:2 * (pos + sz)])
            assert r.count == orc.decimated_count(pos, pos + sz, d)
            audio.append(r.audio.copy())
            pos += sz
        r = bank.process_chunk(raw[2 * pos:])
        audio.append(r.audio.copy())
        got = np.concatenate(audio, axis=1)[0]
        assert bank.get_state()[1] == raw.size // 2
    # NFM has no per-call semantics except the NCO phase wrap (1 ulp of phase per call): same stream as the golden
    assert got.size == g["audio"].size
    assert np.abs(got - g["audio"]).max() <= AUDIO_TOL
    # a capture shorter than the channel filter (6449 taps): only head rows exist
    with gpu["ChannelBank"](fs, d, tg, ref_chunk=m["chunk"]) as bank:
        r = bank.process_chunk(raw[:2 * 1000], want_baseband=True)
        plan = orc.TargetPlan(sample_rate=fs, freq_offset=m["targets"][0]["f_off"], mix_sign=int(g["mix_sign"]))
        x = orc.order_iq(orc.unpack_interleaved(raw[:2000], "pcm_s16le"), "iq")
        want = orc.run_target(x, plan, m["chunk"])
        assert r.count == want.audio.size == orc.decimated_count(0, 1000, d)
        assert np.abs(r.baseband[0] - want.baseband).max() <= BB_TOL
        assert np.abs(r.audio[0] - want.audio).max() <= AUDIO_TOL


def test_bit_faithful_filter_forms(gpu, monkeypatch):
    """The float64 channel filter of the bit-faithful path in its three forms.  Direct form (IQ2A_PRECISE_FIR=direct,
    2 * ntaps DFMAs per sample): reproduces the reference's complex64 channel samples.  Register-pass transform form
    with repair (default, csrc/precise_fft.cu): every sample within 1e-12 of a float32 rounding boundary is
    recomputed by the direct sum, so the result is the direct form's, bit for bit.  First transform form
    (IQ2A_PRECISE_FIR=fft): cancels ~80 dB of out-of-band signal in its polyphase sum and lands on the neighbouring
    float32 for about one sample in 10^4 -- close, but not what this path is for."""
    names = ["case_c_20M_usb", "case_c_20M_lsb"]
    m, fs, d, tg_all, gold_all = _targets(gpu, "case_c_20M_am_ssb", ["case_c_20M_am"] + names)
    raw = _cases.raw_input("case_c_20M_am_ssb")
    chunk = m["chunk"]

    def run():
        with gpu["ChannelBank"](fs, d, tg_all[1:], ref_chunk=chunk) as bank:
            parts = [bank.process_chunk(raw[2 * s:2 * min(s + chunk, raw.size // 2)], want_baseband=True)
                     for s in range(0, raw.size // 2, chunk)]
            return np.concatenate([p.baseband for p in parts], axis=1), np.concatenate([p.clipped for p in parts], axis=1)
    bb_reg, clip_reg = run()
    monkeypatch.setenv("IQ2A_PRECISE_FIR", "direct")
    bb_dir, clip_dir = run()
    monkeypatch.setenv("IQ2A_PRECISE_FIR", "fft")
    bb_fft, clip_fft = run()
    for i, g in enumerate(gold_all[1:]):
        assert np.mean(bb_dir[i] == g["baseband"]) > 0.99999
        assert np.abs(clip_dir[i] - g["clipped"]).max() <= AUDIO_TOL
        assert np.array_equal(bb_reg[i], bb_dir[i]) and np.array_equal(clip_reg[i], clip_dir[i])
        assert np.mean(bb_fft[i] == bb_dir[i]) > 0.999
        assert np.abs(bb_fft[i] - bb_dir[i]).max() <= 1e-7
