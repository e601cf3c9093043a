"""GPU: spectrum previews (csrc/spectrum.cu through the C-ABI) against the reference's golden vectors and the oracle."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import pytest

from oracle import iq_oracle as orc
from oracle import spectrum_oracle as so
from tests._spectrum_cases import PSD_CASES, WATERFALL_CASES, _signal, psd_input, waterfall_chunks

pytestmark = pytest.mark.gpu

GOLD = np.load(Path(__file__).parent / "golden" / "spectrum_vectors.npz")
# float64 on both sides; the transforms differ in operation order (four-step / radix-2^2 here, pocketfft there).
# Rounding noise is ~1e-16 of the strongest bin, which is ~1e-9 dB on bins 60 dB down and more on weaker ones.
DB_TOL = 1e-6


@pytest.mark.parametrize("name", list(PSD_CASES))
def test_compute_psd_golden(name):
    from iq_to_audio_b200.spectrum import compute_psd
    case = PSD_CASES[name]
    freqs, psd = compute_psd(psd_input(case), case["fs"], case["nfft"])
    assert psd.dtype == np.float64 and psd.shape == (case["nfft"],) and freqs.dtype == np.float64
    np.testing.assert_array_equal(np.concatenate([freqs[:4], freqs[-4:]]), GOLD[f"psd_{name}_freqs_edge"])
    np.testing.assert_allclose(psd, GOLD[f"psd_{name}_db"], rtol=0, atol=DB_TOL)


@pytest.mark.parametrize("name", list(WATERFALL_CASES))
def test_streaming_waterfall_golden(name):
    from iq_to_audio_b200.spectrum import streaming_waterfall
    case = WATERFALL_CASES[name]
    freqs, avg, wf, frames = streaming_waterfall(iter(waterfall_chunks(case)), case["fs"], nfft=case["nfft"],
                                                 hop=case["hop"], max_slices=case["max_slices"])
    assert frames == int(GOLD[f"wf_{name}_frames"])
    np.testing.assert_array_equal(np.concatenate([freqs[:4], freqs[-4:]]), GOLD[f"wf_{name}_freqs_edge"])
    np.testing.assert_array_equal(wf.times, GOLD[f"wf_{name}_times"])           # bit-exact bookkeeping
    assert wf.matrix.dtype == np.float32 and wf.matrix.shape == GOLD[f"wf_{name}_matrix"].shape
    np.testing.assert_allclose(avg, GOLD[f"wf_{name}_avg"], rtol=0, atol=DB_TOL)
    np.testing.assert_allclose(wf.matrix, GOLD[f"wf_{name}_matrix"], rtol=0, atol=2e-5)
    np.testing.assert_array_equal(wf.freqs, freqs)


@pytest.mark.parametrize("log2n", [1, 2, 3, 6, 11, 13, 14, 15, 17, 18, 19])
def test_every_transform_shape_vs_oracle(log2n):
    # single-CTA sizes (<= 8192), four-step with 16 columns (<= 2^18) and with 8 columns (2^19)
    from iq_to_audio_b200.spectrum import compute_psd
    nfft = 1 << log2n
    x = _signal(100 + log2n, nfft + 17, 10e6)
    _, psd = compute_psd(x, 10e6, nfft)
    np.testing.assert_allclose(psd, so.psd_one(x, 10e6, nfft), rtol=0, atol=DB_TOL)


def test_front_end_default_size_streaming():
    # interactive/state.py:82: nfft 262144, hop nfft/4; 12 windows with a cap of 5 slices -> two halvings
    from iq_to_audio_b200.spectrum import streaming_waterfall
    nfft = 1 << 18
    x = _signal(7, nfft + 11 * (nfft // 4) + 1000, 10e6)
    chunks = [x[i:i + (1 << 20)] for i in range(0, x.size, 1 << 20)]
    freqs, avg, wf, frames = streaming_waterfall(iter(chunks), 10e6, nfft=nfft, max_slices=5)
    o_avg, o_t, o_m, o_frames = so.waterfall(chunks, 10e6, nfft, None, 5)
    assert frames == o_frames == 12 and wf.matrix.shape == o_m.shape
    np.testing.assert_array_equal(wf.times, o_t)
    np.testing.assert_allclose(avg, o_avg, rtol=0, atol=DB_TOL)
    np.testing.assert_allclose(wf.matrix, o_m, rtol=0, atol=2e-5)
    # the tone at +0.1234 fs is the strongest bin
    assert abs(freqs[np.argmax(avg)] - 0.1234 * 10e6) < 10e6 / nfft


@pytest.mark.parametrize("codec,order", [("pcm_s16le", "iq"), ("pcm_s16le", "qi_inv"), ("pcm_u8", "qi"),
                                         ("pcm_f32le", "iq_inv")])
def test_raw_frames_match_the_reader_conversion(codec, order):
    # the accumulator also takes the reader's raw PCM frames; the result must equal feeding the complex64
    # samples the reference's IQReader would have produced from them (oracle.unpack_interleaved + order_iq)
    from iq_to_audio_b200.spectrum import SpectrumAccumulator
    x = _signal(31, 20000, 2.5e6)
    cols = np.column_stack((x.real, x.imag)).astype(np.float64)
    raw = {"pcm_s16le": orc.to_s16, "pcm_u8": orc.to_u8, "pcm_f32le": orc.to_f32}[codec](cols)
    want_x = orc.order_iq(orc.unpack_interleaved(raw, codec), order)
    with SpectrumAccumulator(2.5e6, nfft=2048, hop=512, max_slices=9, codec=codec, iq_order=order) as acc:
        b = raw.view(np.uint8)
        fb = {"pcm_s16le": 4, "pcm_u8": 2, "pcm_f32le": 8}[codec]
        for lo, hi in [(0, 3000), (3000, 3100), (3100, 20000)]:
            acc.push(b[lo * fb:hi * fb])
        freqs, avg, wf, frames = acc.finish()
        assert acc.counts()[2] > 0                      # kernels were launched
    o_avg, o_t, o_m, o_frames = so.waterfall([want_x[:3000], want_x[3000:3100], want_x[3100:]], 2.5e6, 2048, 512, 9)
    assert frames == o_frames
    np.testing.assert_array_equal(wf.times, o_t)
    np.testing.assert_allclose(avg, o_avg, rtol=0, atol=DB_TOL)
    np.testing.assert_allclose(wf.matrix, o_m, rtol=0, atol=2e-5)


def test_mean_is_linear_in_the_frames():
    # size-independent property: the mean over 2K frames of a periodic stream equals the mean over K frames
    from iq_to_audio_b200.spectrum import streaming_waterfall
    nfft, hop = 4096, 4096
    one = _signal(5, nfft * 8, 1e6)
    _, a1, _, f1 = streaming_waterfall([one], 1e6, nfft=nfft, hop=hop, max_slices=4)
    _, a2, _, f2 = streaming_waterfall([one, one], 1e6, nfft=nfft, hop=hop, max_slices=4)
    assert (f1, f2) == (8, 16)
    np.testing.assert_allclose(a1, a2, rtol=0, atol=1e-9)


def test_errors_match_the_reference():
    from iq_to_audio_b200.spectrum import compute_psd, streaming_waterfall
    with pytest.raises(ValueError, match="empty"):
        compute_psd(np.empty(0, np.complex64), 1e6, 1024)
    with pytest.raises(ValueError, match="enough samples"):
        streaming_waterfall([np.zeros(100, np.complex64), None], 1e6, nfft=256)
    with pytest.raises(ValueError, match="power of two"):
        compute_psd(np.zeros(100, np.complex64), 1e6, 1000)


def test_silence_hits_the_floor():
    from iq_to_audio_b200.spectrum import compute_psd
    _, psd = compute_psd(np.zeros(5000, np.complex64), 1e6, 4096)
    np.testing.assert_array_equal(psd, np.full(4096, -180.0))


def _capture(tmp_path):
    import wave
    from tests import _cases
    raw = _cases.raw_input("case_b_nfm_10M")
    cap = tmp_path / "baseband_100000000Hz_capture.wav"
    with wave.open(str(cap), "wb") as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(10_000_000)
        w.writeframes(raw.tobytes())
    return cap, orc.order_iq(orc.unpack_interleaved(raw, "pcm_s16le"), "qi")


def test_gather_snapshot_matches_reference_flow(tmp_path):
    # interactive/workers.py:36-161: first `seconds` of the capture, raw frames straight to the device
    from iq_to_audio_b200.pipeline import ProcessingConfig
    from iq_to_audio_b200.preview import gather_snapshot
    cap, x = _capture(tmp_path)
    calls = []
    snap = gather_snapshot(ProcessingConfig(in_path=cap, iq_order="qi"), 0.05, nfft=65536, hop=None, max_slices=4,
                           max_in_memory_samples=100_000, progress_cb=lambda s, f: calls.append((s, f)))
    use = x[:500_000]
    o_avg, o_t, o_m, o_frames = so.waterfall([use], 10e6, 65536, None, 4)
    assert snap.fft_frames == o_frames and snap.mode == "samples" and snap.center_freq == 100e6
    assert snap.seconds == 0.05 and snap.sample_rate == 10e6 and calls[-1] == (0.05, 1.0)
    np.testing.assert_array_equal(snap.samples, use[:100_000])
    np.testing.assert_allclose(snap.psd_db, o_avg, rtol=0, atol=DB_TOL)
    f, t, mat = snap.waterfall
    np.testing.assert_array_equal(t, o_t)
    np.testing.assert_allclose(mat, o_m, rtol=0, atol=2e-5)
    np.testing.assert_array_equal(f, so.freq_axis(65536, 10e6))


def test_compute_full_psd_ragged_chunks(tmp_path):
    # interactive/workers.py:164-287: whole recording, chunk = max(config.chunk_size, nfft)
    from iq_to_audio_b200.pipeline import ProcessingConfig
    from iq_to_audio_b200.preview import compute_full_psd
    cap, _ = _capture(tmp_path)
    x = orc.order_iq(orc.unpack_interleaved(__import__("tests._cases", fromlist=["x"]).raw_input("case_b_nfm_10M"),
                                            "pcm_s16le"), "iq")
    msgs = []
    snap = compute_full_psd(ProcessingConfig(in_path=cap, chunk_size=150_001), nfft=32768, hop=10_000, max_slices=16,
                            status_cb=msgs.append)
    chunks = [x[i:i + 150_001] for i in range(0, x.size, 150_001)]
    o_avg, o_t, o_m, o_frames = so.waterfall(chunks, 10e6, 32768, 10_000, 16)
    assert snap.fft_frames == o_frames and snap.samples is None and snap.params["full_capture"] is True
    assert msgs[0].startswith("Reading full recording") and any("Averaging PSD chunk" in m for m in msgs)
    np.testing.assert_array_equal(snap.waterfall[1], o_t)
    np.testing.assert_allclose(snap.psd_db, o_avg, rtol=0, atol=DB_TOL)
    np.testing.assert_allclose(snap.waterfall[2], o_m, rtol=0, atol=2e-5)


def test_preview_errors(tmp_path):
    from iq_to_audio_b200.pipeline import ProcessingConfig
    from iq_to_audio_b200.preview import gather_snapshot
    raw = tmp_path / "cap.cs16"
    raw.write_bytes(np.zeros(4000, np.int16).tobytes())
    with pytest.raises(ValueError, match="sample rate override"):
        gather_snapshot(ProcessingConfig(in_path=raw, center_freq=1e6), 1.0, nfft=1024, hop=None, max_slices=4)
    with pytest.raises(ValueError, match="Center frequency"):
        gather_snapshot(ProcessingConfig(in_path=raw, input_sample_rate=1e6), 1.0, nfft=1024, hop=None, max_slices=4)
    with pytest.raises(ValueError, match="enough samples"):
        gather_snapshot(ProcessingConfig(in_path=raw, input_sample_rate=1e6, center_freq=1e6), 1.0, nfft=4096,
                        hop=None, max_slices=4)


def test_compute_psd_output_shapes():
    # the reference's own test (tests/test_processing.py:63-69), unchanged apart from the import
    from iq_to_audio_b200.spectrum import compute_psd
    sr = 1_000_000.0
    t = np.arange(0, 8192)
    samples = np.exp(1j * 2.0 * np.pi * 100_000 * t / sr).astype(np.complex64)
    freqs, psd = compute_psd(samples, sr, nfft=4096)
    assert freqs.shape == psd.shape
    assert np.isfinite(psd).all()
    assert abs(freqs[np.argmax(psd)] - 100_000.0) <= sr / 4096
