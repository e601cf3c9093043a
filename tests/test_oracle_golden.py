"""Pin the CPU oracle (oracle/iq_oracle.py) against vectors recorded from the
unmodified reference (tests/golden/make_golden.py).  CPU only."""
from __future__ import annotations

import numpy as np
import pytest

from oracle import iq_oracle as orc
from tests import _cases


def _plan(fs, t, *, filter_block, fs_ch=96_000.0):
    return orc.TargetPlan(sample_rate=fs, freq_offset=t["f_off"], bandwidth=t.get("bw", 12_500.0),
                          mode=t["mode"], agc_enabled=t.get("agc", True), filter_block=filter_block,
                          mix_sign=t.get("mix_sign"), fs_ch_target=fs_ch)


def _assert_stream_equal(res: orc.StreamResult, g) -> None:
    assert res.counts == list(g["counts"])                       # exact chunk boundaries / decimation phase
    assert res.mix_sign == int(g["mix_sign"])
    np.testing.assert_array_equal(res.baseband, g["baseband"])   # same numpy/scipy calls: bit-identical
    np.testing.assert_array_equal(res.audio, g["audio"])
    np.testing.assert_array_equal(res.clipped, g["clipped"])
    assert res.peak == float(g["peak"])
    np.testing.assert_allclose(res.rms_dbfs, g["rms_dbfs"], rtol=0, atol=0)
    assert res.final["phase"] == float(g["final_phase"])
    assert res.final["offset"] == int(g["final_offset"])


def test_case_a_benchmark_shape_nfm():
    m = _cases.manifest()["case_a_nfm_2p5M"]
    g = _cases.load("case_a_nfm_2p5M")
    plan = _plan(m["fs"], m["targets"][0] | {"mode": "nfm"}, filter_block=m["filter_block"])
    assert len(plan.taps) == int(g["ntaps"]) and plan.decimation == int(g["decimation"])
    res = orc.run_target(_cases.complex_input("case_a_nfm_2p5M"), plan, m["chunk"])
    _assert_stream_equal(res, g)


@pytest.mark.parametrize("t", range(5))
def test_case_b_five_nfm_targets(t):
    m = _cases.manifest()["case_b_nfm_10M"]
    g = _cases.load(f"case_b_nfm_10M_t{t}")
    res = orc.run_target(_cases.complex_input("case_b_nfm_10M"),
                         _plan(m["fs"], m["targets"][t], filter_block=m["filter_block"]), m["chunk"])
    _assert_stream_equal(res, g)


@pytest.mark.parametrize("t,mode", [(0, "am"), (1, "usb"), (2, "lsb")])
def test_case_c_am_ssb_agc(t, mode):
    m = _cases.manifest()["case_c_20M_am_ssb"]
    g = _cases.load(f"case_c_20M_{mode}")
    res = orc.run_target(_cases.complex_input("case_c_20M_am_ssb"),
                         _plan(m["fs"], m["targets"][t], filter_block=m["filter_block"]), m["chunk"])
    _assert_stream_equal(res, g)


@pytest.mark.parametrize("name", ["case_d_pcm_u8_qi_nfm", "case_d_pcm_f32le_iq_inv_nfm",
                                  "case_d_pcm_s16le_qi_inv_usb"])
def test_case_d_formats_and_iq_order(name):
    m = _cases.manifest()[name]
    g = _cases.load(name)
    t = dict(m["targets"][0])
    t["bw"] = 12_500.0 if t["mode"] == "nfm" else 2_800.0
    res = orc.run_target(_cases.complex_input(name), _plan(m["fs"], t, filter_block=m["filter_block"]), m["chunk"])
    _assert_stream_equal(res, g)


def test_case_e_preview_truncation():
    m = _cases.manifest()["case_e_truncated"]
    g = _cases.load("case_e_truncated")
    plan = orc.TargetPlan(sample_rate=2.5e6, freq_offset=25e3, filter_block=m["filter_block"])
    res = orc.run_target(_cases.complex_input("case_e_truncated"), plan, m["chunk"],
                         max_input_samples=m["max_input_samples"])
    _assert_stream_equal(res, g)
    assert sum(res.counts) == orc.decimated_count(0, m["max_input_samples"], plan.decimation)


# ---------------------------------------------------------------- stage level

@pytest.fixture(scope="module")
def sv():
    return _cases.load("stage_vectors")


def test_stage_mixer(sv):
    st = orc.NcoState.for_offset(123_456.7, 2.4e6)
    z = sv["mix_in"]
    np.testing.assert_array_equal(orc.nco_mix(st, z[:17_000], -1), sv["mix_out_a"])
    assert st.phase == float(sv["mix_phase_a"])
    np.testing.assert_array_equal(orc.nco_mix(st, z[17_000:], -1), sv["mix_out_b"])
    assert st.phase == float(sv["mix_phase_b"])
    assert orc.nco_mix(st, z[:0], 1).size == 0


def test_stage_fir_ragged_calls(sv):
    z = sv["mix_in"]
    st = orc.FirState(sv["fir_taps"], 4096)
    out = np.concatenate([orc.fir_overlap_save(st, z[:5_000]), orc.fir_overlap_save(st, z[5_000:5_700]),
                          orc.fir_overlap_save(st, z[5_700:])])
    np.testing.assert_array_equal(out, sv["fir_out"])
    np.testing.assert_array_equal(st.hist, sv["fir_state"])
    with pytest.raises(ValueError):
        orc.FirState(sv["fir_taps"], 0)


def test_stage_fir_matches_direct_form(sv):
    z = sv["mix_in"]
    at = np.array([0, 1, 7, 500, 1599, 1600, 1601, 12_345, 29_999])
    exact = orc.fir_direct_f64(sv["fir_taps"], z, at)
    np.testing.assert_allclose(sv["fir_out"][at], exact, rtol=0, atol=2e-7)


def test_stage_decimator_known_answer(sv):
    d3 = orc.DecimState(3)
    got = np.concatenate([orc.decimate(d3, np.arange(9, dtype=np.complex64)),
                          orc.decimate(d3, np.arange(9, 18, dtype=np.complex64))])
    np.testing.assert_array_equal(got, np.arange(0, 18, 3, dtype=np.complex64))   # ref tests/test_processing.py:22-28
    np.testing.assert_array_equal(got, sv["dec3"])
    z = sv["mix_in"]
    d7 = orc.DecimState(7)
    got = np.concatenate([orc.decimate(d7, z[:10]), orc.decimate(d7, z[10:11]), orc.decimate(d7, z[11:400])])
    np.testing.assert_array_equal(got, sv["dec7"])
    assert d7.offset == int(sv["dec7_offset"])
    assert got.size == orc.decimated_count(0, 400, 7)


def test_stage_discriminator_deemphasis(sv):
    z = sv["mix_in"]
    ds = orc.DiscState()
    de = orc.DeemphState.design(300.0, 96_153.846)
    assert de.alpha == float(sv["deemph_alpha"])
    a1 = orc.fm_discriminate(ds, z[:9_000])
    a2 = orc.fm_discriminate(ds, z[9_000:20_000])
    np.testing.assert_array_equal(np.concatenate([a1, a2]), sv["disc"])
    y = np.concatenate([orc.deemphasis(de, a1), orc.deemphasis(de, a2)])
    np.testing.assert_array_equal(y, sv["deemph"])
    assert de.z == float(sv["deemph_state"])


def test_stage_dc_blocker_and_agc_bit_exact(sv):
    r = sv["dc_in"]
    st = orc.DcBlockState()
    y = np.concatenate([orc.dc_block(st, r[:2_500]), orc.dc_block(st, r[2_500:])])
    np.testing.assert_array_equal(y, sv["dc_out"])            # C helper == reference's Python loop, bit for bit
    g = np.concatenate([orc.agc(y[:2_500]), orc.agc(y[2_500:])])
    np.testing.assert_array_equal(g, sv["agc_out"])
    # the literal Python-loop forms agree with the C helper as well
    st2 = orc.DcBlockState()
    y2 = np.concatenate([orc.dc_block_pyloop(st2, r[:2_500]), orc.dc_block_pyloop(st2, r[2_500:])])
    np.testing.assert_array_equal(y2, y)
    np.testing.assert_array_equal(orc.agc_pyloop(y[:2_500]), g[:2_500])


def test_mix_sign_known_answers(sv):
    taps = orc.channel_taps(1e6, 12_500.0, 10)
    nn = np.arange(0, int(1e6 * 0.1))
    warm = np.exp(1j * 2.0 * np.pi * 12_500.0 * nn / 1e6).astype(np.complex64)
    assert orc.pick_mix_sign(warm, 1e6, 12_500.0, taps, 10) == 1 == int(sv["mix_sign_pos_tone"])   # ref tests/test_processing.py:31-40
    assert orc.pick_mix_sign(np.conj(warm), 1e6, 12_500.0, taps, 10) == int(sv["mix_sign_neg_tone"]) == -1
    assert orc.pick_mix_sign(warm[:0], 1e6, 12_500.0, taps, 10) == 1


def test_planning_helpers(sv):
    for fs, req, want in sv["tune_chunk"]:
        assert orc.plan_chunk(float(fs), int(req)) == int(want)
    for fs, bw, d, n in sv["ntaps_table"]:
        assert len(orc.channel_taps(float(fs), float(bw), int(d))) == int(n)
    assert orc.plan_decimation(2.5e6, 96e3) == (26, 2.5e6 / 26)
    assert orc.plan_decimation(61.44e6, 96e3)[0] == 640


def test_unpack_rules_and_ragged_tail():
    raw = np.array([-32768, 32767, 0, 1, 5], dtype=np.int16)       # odd int16 count -> partial frame dropped
    il = orc.unpack_interleaved(raw.tobytes()[:9], "pcm_s16le")
    np.testing.assert_array_equal(il, np.array([-1.0, 32767 / 32768, 0.0, 1 / 32768], dtype=np.float32))
    u = orc.unpack_interleaved(np.array([0, 255, 128, 129], dtype=np.uint8), "pcm_u8")
    np.testing.assert_array_equal(u, np.array([-1.0, 127 / 128, 0.0, 1 / 128], dtype=np.float32))
    z = orc.order_iq(np.array([1, 2, 3, 4], dtype=np.float32), "qi_inv")
    np.testing.assert_array_equal(z, np.array([2 - 1j, 4 - 3j], dtype=np.complex64))
    with pytest.raises(ValueError):
        orc.order_iq(il, "xy")
    assert orc.unpack_interleaved(b"", "pcm_s16le").size == 0


def test_oracle_reproduces_the_reference_pipeline_run_bit_for_bit(tmp_path):
    """`tests/golden/pipeline_cfg1.npz` holds the float32 stream the UNMODIFIED reference handed to its encoder in a
    `run_benchmark -> ProcessingPipeline.run` run of the repo's own --benchmark configuration (2.5 MS/s, 5 s, NFM at
    +25 kHz, chunk 1 Mi, filter block 65 536, mix sign probed) -- `make_pipeline_golden.py` ran it with a stub
    soundfile and a shim ffmpeg.  The oracle's replay of the loop on the same capture must give the same bytes, and the
    capture written by the product's own `--benchmark` generator must be the capture the reference generated."""
    import hashlib
    from pathlib import Path
    from iq_to_audio_b200.benchmark import write_synthetic_capture

    g = np.load(Path(__file__).parent / "golden" / "pipeline_cfg1.npz")
    cap = tmp_path / "benchmark_fc-400000000Hz.wav"
    n = write_synthetic_capture(cap, float(g["sample_rate"]), float(g["seconds"]), float(g["freq_offset"]))
    blob = cap.read_bytes()
    pos = blob.index(b"data") + 8
    assert n == int(g["capture_frames"]) and hashlib.sha256(blob[pos:]).hexdigest() == str(g["capture_sha256"])
    raw = np.frombuffer(blob[pos:], dtype="<i2")
    x = orc.order_iq(orc.unpack_interleaved(raw, "pcm_s16le"), "iq")
    plan = orc.TargetPlan(sample_rate=float(g["sample_rate"]), freq_offset=float(g["freq_offset"]), mix_sign=None)
    res = orc.run_target(x, plan, orc.plan_chunk(float(g["sample_rate"]), 1_048_576))
    assert int(g["ffmpeg_rate"]) == int(round(plan.fs_channel)) == 96_154
    assert res.clipped.dtype == np.float32 and res.clipped.size == g["clipped"].size
    assert np.array_equal(res.clipped, g["clipped"])
