"""GPU: the host pipeline (reader -> ChannelBank.stream -> writer), `--benchmark`, cancellation."""
from __future__ import annotations

import wave

import numpy as np
import pytest

from oracle import iq_oracle as orc
from oracle.swr_model import SwrModel
from tests import _cases

pytestmark = pytest.mark.gpu


def _write_wav(path, raw_i16, rate):
    with wave.open(str(path), "wb") as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(int(rate))
        w.writeframes(raw_i16.tobytes())


def _read_pcm(path):
    with wave.open(str(path), "rb") as w:
        assert w.getnchannels() == 1 and w.getsampwidth() == 2
        return w.getframerate(), np.frombuffer(w.readframes(w.getnframes()), dtype="<i2")


@pytest.fixture()
def no_ffmpeg(monkeypatch):
    # force the native encoder (GPU resampler to 48 kHz + PCM_16) even where an ffmpeg binary exists
    monkeypatch.setenv("IQ_TO_AUDIO_B200_NATIVE_WAV", "1")


def test_pipeline_five_targets_one_pass(tmp_path, no_ffmpeg):
    from iq_to_audio_b200.pipeline import ProcessingConfig, ProcessingPipeline
    m = _cases.manifest()["case_b_nfm_10M"]
    cap = tmp_path / "baseband_100000000Hz_capture.wav"
    _write_wav(cap, _cases.raw_input("case_b_nfm_10M"), m["fs"])
    fc = 100_000_000.0
    cfg = ProcessingConfig(in_path=cap, target_freqs=[fc + t["f_off"] for t in m["targets"]], target_freq=fc,
                           chunk_size=m["chunk"], mix_sign_override=1, output_path=tmp_path / "out.wav")
    # chunk_size is only a lower bound (tune_chunk_size): 10 MS/s -> 4 Mi, i.e. the whole test capture is one chunk
    res = ProcessingPipeline(cfg).run_many()
    assert len(res) == 5 and cfg.center_freq == fc and cfg.center_freq_source == "filename:sdrpp"
    for i, r in enumerate(res):
        g = _cases.load(f"case_b_nfm_10M_t{i}")
        assert r.decimation == 104 and r.mix_sign == 1 and r.output_path.name == f"out_{int(round(r.target_freq))}.wav"
        rate, pcm = _read_pcm(r.output_path)
        want = SwrModel(96_154).resample_s16(g["clipped"])        # ffmpeg is told round(fs_ch) = 96154 (ref :391)
        assert rate == 48_000 and pcm.size == want.size
        assert np.abs(pcm.astype(np.int64) - want.astype(np.int64)).max() <= 1    # +-1 LSB on the int16 WAV
        assert abs(r.audio_peak - float(g["peak"])) <= 1e-5


def test_pipeline_auto_mix_sign_and_preview(tmp_path, no_ffmpeg):
    from iq_to_audio_b200.pipeline import ProcessingConfig, ProcessingPipeline
    cap = tmp_path / "bench_fc-400000000Hz.wav"
    _write_wav(cap, _cases.raw_input("case_a_nfm_2p5M"), 2.5e6)
    cfg = ProcessingConfig(in_path=cap, target_freq=400_025_000.0, max_input_seconds=0.06, output_path=tmp_path / "a.wav")
    r = ProcessingPipeline(cfg).run()
    g = _cases.load("case_a_nfm_2p5M")
    assert r.mix_sign == int(g["mix_sign"]) and r.decimation == 26
    assert r.samples_processed == 150_000
    rate, pcm = _read_pcm(tmp_path / "a.wav")
    rows = orc.decimated_count(0, 150_000, 26)
    # the reference tunes the chunk to 1 Mi at 2.5 MS/s: the 150 000-sample preview is one chunk, NFM has no
    # per-chunk semantics, so the channel-rate audio equals the head of the golden stream
    want = SwrModel(96_154).resample_s16(g["clipped"][:rows])
    assert rate == 48_000 and pcm.size == want.size
    assert np.abs(pcm.astype(np.int64) - want.astype(np.int64)).max() <= 1


def test_pipeline_cancel_removes_output(tmp_path, no_ffmpeg):
    from iq_to_audio_b200.pipeline import ProcessingCancelled, ProcessingConfig, ProcessingPipeline
    cap = tmp_path / "x_fc-400000000Hz.wav"
    _write_wav(cap, _cases.raw_input("case_a_nfm_2p5M"), 2.5e6)
    cfg = ProcessingConfig(in_path=cap, target_freq=400_025_000.0, output_path=tmp_path / "c.wav", mix_sign_override=1)
    pipe = ProcessingPipeline(cfg)

    class CancelOnFirstAdvance:                      # ref tests/test_processing.py:98-151 (_AutoCancelSink)
        def start(self, phases, *, overall_total): pass
        def advance(self, phase, delta, *, overall_completed, overall_total): pipe.cancel()
        def status(self, message): pass
        def close(self): pass
    with pytest.raises(ProcessingCancelled):
        pipe.run(CancelOnFirstAdvance())
    assert not (tmp_path / "c.wav").exists()


def test_cli_benchmark_surface(no_ffmpeg, caplog):
    from iq_to_audio_b200 import cli
    import logging
    with caplog.at_level(logging.INFO):
        rc = cli.main(["--benchmark", "--benchmark-seconds", "0.5"])
    assert rc == 0
    assert any("Benchmark processed 1250000 IQ samples" in r.getMessage() for r in caplog.records)
    assert cli.main(["--benchmark", "--benchmark-offset", "2000000"]) == 1        # offset beyond fs/2 -> error exit
    with pytest.raises(SystemExit):                                               # duplicate targets (cli.py:516-519)
        cli.main(["--ft", "1000", "--ft", "1000.2", "--in", "x.wav"])
    with pytest.raises(SystemExit):                                               # the cap is the bank's, not five
        cli.main(sum((["--ft", str(1000 + 10 * i)] for i in range(cli.MAX_TARGETS + 1)), []) + ["--in", "x.wav"])


def test_pipeline_errors(tmp_path, no_ffmpeg):
    from iq_to_audio_b200.pipeline import ProcessingConfig, ProcessingPipeline
    cap = tmp_path / "nofreq.wav"
    _write_wav(cap, _cases.raw_input("case_a_nfm_2p5M")[:20_000], 2.5e6)
    with pytest.raises(ValueError, match="Center frequency"):
        ProcessingPipeline(ProcessingConfig(in_path=cap, target_freq=1e6)).run()
    with pytest.raises(ValueError, match="Target frequency must be positive"):
        ProcessingPipeline(ProcessingConfig(in_path=cap, center_freq=1e6)).run()
    raw = tmp_path / "cap.cs16"
    raw.write_bytes(_cases.raw_input("case_a_nfm_2p5M")[:20_000].tobytes())
    with pytest.raises(ValueError, match="input-sample-rate"):
        ProcessingPipeline(ProcessingConfig(in_path=raw, target_freq=1.0e6 + 25e3, center_freq=1e6)).run()
    r = ProcessingPipeline(ProcessingConfig(in_path=raw, target_freq=1.0e6 + 25e3, center_freq=1e6,
                                            input_sample_rate=2.5e6, output_path=tmp_path / "r.wav",
                                            mix_sign_override=1)).run()
    assert r.samples_processed == 10_000


@pytest.mark.parametrize("key,pieces", [("r96154_n100991", [40_330, 40_330, 20_331]), ("r96000_n26219", [6_554, 6_554, 6_554, 6_557]),
                                        ("r95238_n20000", [20_000]), ("r100000_n12345", [5_000, 7_345]), ("r96154_n333", [333])])
def test_gpu_resampler_matches_libswresample_vectors(key, pieces):
    """K14 on the GPU against vectors from the real libswresample: sample counts exact (per call and
    after flush), int16 within +-1 LSB."""
    from iq_to_audio_b200.resample import Resampler48k
    rv = _cases.load("resampler_vectors")
    in_rate = int(key.split("_")[0][1:])
    x = rv[key + "_in"]
    model = SwrModel(in_rate)
    res = Resampler48k(in_rate, 2)                         # two channels: x and -x
    outs, pos = [], 0
    for n in pieces:
        outs.append(res.process(np.stack([x[pos:pos + n], -x[pos:pos + n]])))
        pos += n
    outs.append(res.flush())
    res.close()
    got = np.concatenate(outs, axis=1)
    ref = rv[key + "_s16"]
    assert got.shape == (2, ref.size)
    if key + "_counts" in rv and len(pieces) == len(rv[key + "_counts"]) - 1:
        assert [o.shape[1] for o in outs] == list(rv[key + "_counts"])          # per-call counts as the library's
    assert np.abs(got[0].astype(np.int64) - ref.astype(np.int64)).max() <= 1
    assert np.mean(got[0] == ref) > 0.995
    assert np.abs(got[0].astype(np.int64) - model.resample_s16(x).astype(np.int64)).max() <= 1
    assert np.abs(got[1].astype(np.int64) + got[0].astype(np.int64)).max() <= 1   # odd symmetry up to rounding ties
    with pytest.raises(ValueError):
        Resampler48k(48_000, 1)


def test_pass_through_writes_iq_slices_wav(tmp_path):
    # --demod none (processing.py:693-695, :1114-1121): the channelised IQ of every target, in the capture's own
    # container and sample format, plus one cf32 debug dump per target (cli.py:623)
    from iq_to_audio_b200.pipeline import ProcessingConfig, ProcessingPipeline
    m = _cases.manifest()["case_b_nfm_10M"]
    cap = tmp_path / "baseband_100000000Hz_capture.wav"
    _write_wav(cap, _cases.raw_input("case_b_nfm_10M"), m["fs"])
    fc = 100_000_000.0
    cfg = ProcessingConfig(in_path=cap, target_freqs=[fc + t["f_off"] for t in m["targets"]], target_freq=fc,
                           chunk_size=m["chunk"], mix_sign_override=1, demod_mode="none",
                           dump_iq_path=tmp_path / "dbg.cf32")
    res = ProcessingPipeline(cfg).run_many()
    assert len(res) == 5
    for i, r in enumerate(res):
        g = _cases.load(f"case_b_nfm_10M_t{i}")
        bb = g["baseband"]
        assert r.output_path.name == f"slice_{int(r.target_freq)}.wav"           # processing.py:1216-1232
        with wave.open(str(r.output_path), "rb") as w:
            assert (w.getnchannels(), w.getsampwidth(), w.getframerate()) == (2, 2, 96_154)
            got = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").astype(np.int64)
        want = np.rint(bb.view(np.float32).astype(np.float64) * 32767.0).astype(np.int64)
        assert got.size == want.size
        assert np.max(np.abs(got - want)) <= 1
        assert abs(r.audio_peak - float(np.max(np.abs(bb)))) < 1e-5
        dbg = np.fromfile(tmp_path / f"dbg_{int(round(r.target_freq))}.cf32", dtype=np.complex64)
        assert dbg.size == bb.size and np.max(np.abs(dbg - bb)) < 2e-6


@pytest.mark.parametrize("suffix,codec", [(".cs16", "pcm_s16le"), (".cu8", "pcm_u8"), (".cf32", "pcm_f32le")])
def test_pass_through_raw_containers(tmp_path, suffix, codec):
    from iq_to_audio_b200.pipeline import ProcessingConfig, ProcessingPipeline
    m = _cases.manifest()["case_a_nfm_2p5M"]
    raw16 = _cases.raw_input("case_a_nfm_2p5M")
    x = raw16.astype(np.float64) / 32768.0
    raw = {"pcm_s16le": raw16, "pcm_u8": orc.to_u8(x.reshape(-1, 2)), "pcm_f32le": x.astype(np.float32)}[codec]
    cap = tmp_path / f"capture_433000000Hz{suffix}"
    raw.tofile(cap)
    cfg = ProcessingConfig(in_path=cap, target_freq=433_000_000.0 + m["targets"][0]["f_off"], chunk_size=m["chunk"],
                           mix_sign_override=1, demod_mode="pass", input_sample_rate=m["fs"])
    r = ProcessingPipeline(cfg).run()
    assert r.output_path.name == f"slice_{int(r.target_freq)}{suffix}"
    # what the reference stages produce from the same frames (oracle), then its raw encoding rule (:527-539)
    xin = orc.order_iq(orc.unpack_interleaved(raw, codec), "iq")
    plan = orc.TargetPlan(m["fs"], m["targets"][0]["f_off"], m["targets"][0]["bw"], "nfm", mix_sign=1)
    bb = orc.run_target(xin, plan, chunk=1 << 20).baseband.view(np.float32).astype(np.float64)
    data = np.fromfile(r.output_path, dtype={"pcm_s16le": "<i2", "pcm_u8": np.uint8, "pcm_f32le": "<f4"}[codec])
    assert data.size == bb.size
    if codec == "pcm_s16le":
        want = (np.clip(bb, -1.0, 0.999969) * 32767.0).astype(np.int64)          # truncation, as astype does
        assert np.max(np.abs(data.astype(np.int64) - want)) <= 1
    elif codec == "pcm_u8":
        want = np.round((np.clip(bb, -1.0, 1.0) + 1.0) * 127.5).astype(np.int64)
        assert np.max(np.abs(data.astype(np.int64) - want)) <= 1
    else:
        assert np.max(np.abs(data.astype(np.float64) - bb)) < 2e-6
    assert abs(r.audio_peak - float(np.max(np.abs(bb.view(np.complex128))))) < 1e-5


def test_benchmark_run_matches_the_unmodified_reference_pipeline(tmp_path, no_ffmpeg, monkeypatch):
    """Whole-pipeline surface at the repo's own --benchmark configuration (2.5 MS/s, 5 s, NFM at +25 kHz, chunk
    1 Mi, filter block 65 536, mix sign probed on the warm-up chunk): the float32 stream the product pipeline hands
    to its encoder against the stream the UNMODIFIED reference handed to ffmpeg (tests/golden/pipeline_cfg1.npz,
    make_pipeline_golden.py), <= 1e-4; the 48 kHz PCM_16 file against libswresample's arithmetic on the reference's
    stream, +-1 LSB; the running peak to 1e-5."""
    import hashlib
    from pathlib import Path
    from iq_to_audio_b200 import pipeline as gp
    from iq_to_audio_b200.benchmark import write_synthetic_capture

    g = np.load(Path(__file__).parent / "golden" / "pipeline_cfg1.npz")
    cap = tmp_path / "benchmark_fc-400000000Hz.wav"
    n = write_synthetic_capture(cap, float(g["sample_rate"]), float(g["seconds"]), float(g["freq_offset"]))
    blob = cap.read_bytes()
    assert n == int(g["capture_frames"])
    assert hashlib.sha256(blob[blob.index(b"data") + 8:]).hexdigest() == str(g["capture_sha256"])   # same input
    parts = []
    real = gp.AudioWriter.write_clipped

    def spy(self, safe):
        parts.append(np.array(safe, dtype=np.float32, copy=True))
        return real(self, safe)

    monkeypatch.setattr(gp.AudioWriter, "write_clipped", spy)
    cfg = gp.ProcessingConfig(in_path=cap, target_freq=400_025_000.0, center_freq=400_000_000.0,
                              center_freq_source="benchmark", demod_mode="nfm", output_path=tmp_path / "audio.wav")
    r = gp.ProcessingPipeline(cfg).run()
    stream = np.concatenate(parts)
    want = g["clipped"]
    assert r.decimation == 26 and r.mix_sign == 1 and stream.size == want.size == 480_770
    assert np.abs(stream - want).max() <= 1e-4                    # measured ~3e-8
    assert abs(r.audio_peak - float(np.abs(want).max())) <= 1e-5  # nothing clips in this run: peak == max |stream|
    rate, pcm = _read_pcm(r.output_path)
    ref_pcm = SwrModel(int(g["ffmpeg_rate"])).resample_s16(want)
    assert rate == 48_000 and pcm.size == ref_pcm.size
    assert np.abs(pcm.astype(np.int64) - ref_pcm.astype(np.int64)).max() <= 1
