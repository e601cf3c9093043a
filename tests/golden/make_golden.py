#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED reference classes.

Run in the build container only (``/root/reference`` does not exist on the GPU
box):

    python tests/golden/make_golden.py

It imports ``iq_to_audio.processing`` / ``iq_to_audio.decoders`` straight from
``/root/reference/src`` (only ``soundfile`` is stubbed -- the hot-path classes
never call it), drives them in the order of the reference's loop body
(``src/iq_to_audio/processing.py:1070-1154``) on small seeded captures and
stores what they return in ``tests/golden/*.npz``.  Inputs are stored too when
small, otherwise as a generator recipe + SHA-256 so the tests can rebuild them
bit for bit.

Nothing from the reference is copied into the repository: only its outputs.
"""
from __future__ import annotations

import hashlib
import json
import sys
import types
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
REF_SRC = Path("/root/reference/src")

sys.path.insert(0, str(ROOT))
from oracle import iq_oracle as orc  # noqa: E402  (only its synthetic-capture generators are used here)


def _import_reference():
    if not REF_SRC.exists():
        raise SystemExit("reference tree not present; goldens can only be regenerated in the build container")
    stub = types.ModuleType("soundfile")

    def _nope(*a, **k):
        raise RuntimeError("soundfile stub")

    stub.info = _nope
    stub.read = _nope
    stub.write = _nope
    stub.SoundFile = object
    sys.modules.setdefault("soundfile", stub)
    sys.path.insert(0, str(REF_SRC))
    import iq_to_audio.decoders as dec
    import iq_to_audio.processing as proc
    from iq_to_audio.decoders.common import DCBlocker
    from iq_to_audio.decoders.nfm import DeemphasisFilter, QuadratureDemod
    return proc, dec, DCBlocker, DeemphasisFilter, QuadratureDemod


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_reference_loop(proc, dec, x_c64, *, fs, f_off, bw, mode, chunk, filter_block,
                       deemph_us=300.0, agc=True, fs_ch=96_000.0, mix_sign=None,
                       max_input_samples=None):
    """The reference loop body on an in-memory capture, one target."""
    D = max(1, int(round(fs / fs_ch)))
    fs_channel = fs / D
    if fs_channel > fs_ch * 1.5:
        D = max(int(np.floor(fs / fs_ch)), 1)
        fs_channel = fs / D
    taps = proc.design_channel_filter(fs, bw, D)
    osc = proc.ComplexOscillator(f_off, fs)
    fir = proc.OverlapSaveFIR(taps, filter_block)
    decim = proc.Decimator(D)
    decoder = dec.create_decoder(mode, deemph_us=deemph_us, agc_enabled=agc)
    decoder.setup(fs_channel)
    warm = x_c64[:chunk]
    if max_input_samples is not None and warm.size > max_input_samples:
        warm = warm[:max_input_samples]
    sign = mix_sign if mix_sign in (1, -1) else proc.choose_mix_sign(warm, fs, f_off, taps, D)
    audio, clipped, bb, counts, rms = [], [], [], [], []
    first_mixed = None
    peak = 0.0
    done = 0
    for s in range(0, x_c64.size, chunk):
        blk = x_c64[s:s + chunk]
        if max_input_samples is not None:
            left = max_input_samples - done
            if left <= 0:
                break
            blk = blk[:left]
        done += blk.size
        mixed = osc.mix(blk, sign)
        if first_mixed is None:
            first_mixed = mixed[:4096].copy()
        filt = fir.process(mixed)
        d = decim.process(filt)
        a, st = decoder.process(d)
        # AudioWriter.write arithmetic (processing.py:449-453) without the ffmpeg pipe
        pk = float(np.max(np.abs(a))) if a.size else 0.0
        peak = max(peak, pk)
        clipped.append(np.clip(a, -0.99, 0.99).astype(np.float32, copy=False))
        audio.append(np.asarray(a, dtype=np.float32))
        bb.append(np.asarray(d, dtype=np.complex64))
        counts.append(int(d.size))
        rms.append(float(st.rms_dbfs))
        if max_input_samples is not None and done >= max_input_samples:
            break
    return dict(
        audio=np.concatenate(audio), clipped=np.concatenate(clipped), baseband=np.concatenate(bb),
        counts=np.asarray(counts, dtype=np.int64), rms_dbfs=np.asarray(rms, dtype=np.float64),
        peak=np.float64(peak), mix_sign=np.int64(sign), ntaps=np.int64(len(taps)), decimation=np.int64(D),
        fs_channel=np.float64(fs_channel), first_mixed=first_mixed,
        final_phase=np.float64(osc.phase), final_offset=np.int64(decim.offset),
    )


def main() -> None:
    proc, dec, DCBlocker, DeemphasisFilter, QuadratureDemod = _import_reference()
    manifest: dict[str, dict] = {}

    # ---- case A: the --benchmark shape (cfg1), shortened; int16 input ------------
    fs, f_off = 2.5e6, 25e3
    raw = orc.benchmark_capture_s16(fs, 0.12, f_off)
    x = orc.order_iq(orc.unpack_interleaved(raw, "pcm_s16le"), "iq")
    out = run_reference_loop(proc, dec, x, fs=fs, f_off=f_off, bw=12_500.0, mode="nfm",
                             chunk=65_536, filter_block=16_384)
    np.savez_compressed(HERE / "case_a_nfm_2p5M.npz", **out)
    manifest["case_a_nfm_2p5M"] = dict(
        gen="benchmark_capture_s16(2.5e6, 0.12, 25e3)", input_sha256=sha(raw), codec="pcm_s16le",
        iq_order="iq", fs=fs, targets=[dict(f_off=f_off, bw=12_500.0, mode="nfm")],
        chunk=65_536, filter_block=16_384)

    # ---- case B: cfg2 shape -- 10 MS/s, 5 NFM targets, int16 ----------------------
    fs = 10e6
    offs = [-3.2e6, -1.1e6, 0.4e6, 2.3e6, 4.1e6]
    carriers = [dict(offset=o, amp=0.12, kind="fm", tone=700.0 + 150.0 * i, dev=2500.0)
                for i, o in enumerate(offs)]
    n = 1_000_000
    raw = orc.to_s16(orc.multi_carrier_capture(fs, n, carriers))
    x = orc.order_iq(orc.unpack_interleaved(raw, "pcm_s16le"), "iq")
    tg = []
    for i, o in enumerate(offs):
        out = run_reference_loop(proc, dec, x, fs=fs, f_off=o, bw=12_500.0, mode="nfm",
                                 chunk=262_144, filter_block=65_536, mix_sign=1)
        np.savez_compressed(HERE / f"case_b_nfm_10M_t{i}.npz", **out)
        tg.append(dict(f_off=o, bw=12_500.0, mode="nfm", mix_sign=1))
    manifest["case_b_nfm_10M"] = dict(
        gen="to_s16(multi_carrier_capture(10e6, 1000000, carriers))", carriers=carriers,
        input_sha256=sha(raw), codec="pcm_s16le", iq_order="iq", fs=fs, targets=tg,
        chunk=262_144, filter_block=65_536)

    # ---- case C: cfg3 shape -- 20 MS/s, AM + USB + LSB, AGC on ---------------------
    fs = 20e6
    carriers = [dict(offset=-6.0e6, amp=0.2, kind="am", tone=1000.0, depth=0.8),
                dict(offset=1.5e6, amp=0.2, kind="usb", tone=1100.0),
                dict(offset=7.25e6, amp=0.2, kind="lsb", tone=900.0)]
    n = 1_600_000
    raw = orc.to_s16(orc.multi_carrier_capture(fs, n, carriers))
    x = orc.order_iq(orc.unpack_interleaved(raw, "pcm_s16le"), "iq")
    tg = []
    for i, (c, mode, bw) in enumerate(zip(carriers, ("am", "usb", "lsb"), (10_000.0, 2_800.0, 2_800.0))):
        out = run_reference_loop(proc, dec, x, fs=fs, f_off=c["offset"], bw=bw, mode=mode,
                                 chunk=524_288, filter_block=65_536, mix_sign=1, agc=True)
        np.savez_compressed(HERE / f"case_c_20M_{mode}.npz", **out)
        tg.append(dict(f_off=c["offset"], bw=bw, mode=mode, mix_sign=1, agc=True))
    manifest["case_c_20M_am_ssb"] = dict(
        gen="to_s16(multi_carrier_capture(20e6, 1600000, carriers))", carriers=carriers,
        input_sha256=sha(raw), codec="pcm_s16le", iq_order="iq", fs=fs, targets=tg,
        chunk=524_288, filter_block=65_536)

    # ---- case D: formats and IQ order (u8 / f32, qi / iq_inv / qi_inv), SSB without AGC
    fs = 2.4e6
    carriers = [dict(offset=300e3, amp=0.3, kind="fm", tone=1000.0, dev=2500.0),
                dict(offset=-450e3, amp=0.25, kind="usb", tone=800.0)]
    n = 200_000
    cols = orc.multi_carrier_capture(fs, n, carriers, seed=7)
    for codec, packer, order, mode, f_off, agc in (
            ("pcm_u8", orc.to_u8, "qi", "nfm", 300e3, True),
            ("pcm_f32le", orc.to_f32, "iq_inv", "nfm", 300e3, True),
            ("pcm_s16le", orc.to_s16, "qi_inv", "usb", -450e3, False)):
        raw = packer(cols)
        x = orc.order_iq(orc.unpack_interleaved(raw, codec), order)
        out = run_reference_loop(proc, dec, x, fs=fs, f_off=f_off, bw=12_500.0 if mode == "nfm" else 2_800.0,
                                 mode=mode, chunk=50_000, filter_block=8_192, agc=agc)
        name = f"case_d_{codec}_{order}_{mode}"
        np.savez_compressed(HERE / f"{name}.npz", **out)
        manifest[name] = dict(gen=f"{packer.__name__}(multi_carrier_capture(2.4e6, 200000, carriers, seed=7))",
                              carriers=carriers, input_sha256=sha(raw), codec=codec, iq_order=order, fs=fs,
                              targets=[dict(f_off=f_off, mode=mode, agc=agc)], chunk=50_000, filter_block=8_192)

    # ---- case E: preview truncation (max_input_samples) --------------------------
    fs, f_off = 2.5e6, 25e3
    raw = orc.benchmark_capture_s16(fs, 0.12, f_off)
    x = orc.order_iq(orc.unpack_interleaved(raw, "pcm_s16le"), "iq")
    out = run_reference_loop(proc, dec, x, fs=fs, f_off=f_off, bw=12_500.0, mode="nfm",
                             chunk=65_536, filter_block=16_384, max_input_samples=150_001)
    np.savez_compressed(HERE / "case_e_truncated.npz", **out)
    manifest["case_e_truncated"] = dict(gen="benchmark_capture_s16(2.5e6, 0.12, 25e3)", input_sha256=sha(raw),
                                        max_input_samples=150_001, chunk=65_536, filter_block=16_384)

    # ---- stage-level vectors -------------------------------------------------------
    rng = np.random.default_rng(2024)
    z = (rng.normal(size=30_000) + 1j * rng.normal(size=30_000)).astype(np.complex64) * np.float32(0.3)
    stage = {}
    # mixer: two calls, carried phase
    osc = proc.ComplexOscillator(123_456.7, 2.4e6)
    stage["mix_in"] = z
    stage["mix_out_a"] = osc.mix(z[:17_000], -1)
    stage["mix_phase_a"] = np.float64(osc.phase)
    stage["mix_out_b"] = osc.mix(z[17_000:], -1)
    stage["mix_phase_b"] = np.float64(osc.phase)
    # FIR: three ragged calls
    taps = proc.design_channel_filter(2.4e6, 12_500.0, 25)
    fir = proc.OverlapSaveFIR(taps, 4096)
    stage["fir_taps"] = taps
    stage["fir_out"] = np.concatenate([fir.process(z[:5_000]), fir.process(z[5_000:5_700]), fir.process(z[5_700:])])
    stage["fir_state"] = fir.state.copy()
    # decimator: the reference's own known-answer test (tests/test_processing.py:22-28) + ragged
    d3 = proc.Decimator(3)
    stage["dec3"] = np.concatenate([d3.process(np.arange(9, dtype=np.complex64)),
                                    d3.process(np.arange(9, 18, dtype=np.complex64))])
    d7 = proc.Decimator(7)
    stage["dec7"] = np.concatenate([d7.process(z[:10]), d7.process(z[10:11]), d7.process(z[11:400])])
    stage["dec7_offset"] = np.int64(d7.offset)
    # discriminator + de-emphasis over two calls
    qd = QuadratureDemod()
    de = DeemphasisFilter(300.0, 96_153.846)
    a1 = qd.process(z[:9_000]); a2 = qd.process(z[9_000:20_000])
    stage["disc"] = np.concatenate([a1, a2])
    stage["deemph"] = np.concatenate([de.process(a1), de.process(a2)])
    stage["deemph_state"] = np.float64(de.state)
    stage["deemph_alpha"] = np.float64(de.alpha)
    # DC blocker + AGC over two calls (AGC resets per call)
    r = (rng.normal(size=6_000) * 0.05 + 0.2 * np.sin(np.arange(6_000) * 0.07)).astype(np.float32)
    r[100:140] = 0.0          # exercise the 1e-6 floor branch
    dcb = DCBlocker()
    y1 = dcb.process(r[:2_500]); y2 = dcb.process(r[2_500:])
    stage["dc_in"] = r
    stage["dc_out"] = np.concatenate([y1, y2])
    ssb = dec.create_decoder("usb", deemph_us=300.0, agc_enabled=True)
    stage["agc_out"] = np.concatenate([ssb._apply_agc(y1), ssb._apply_agc(y2)])
    # choose_mix_sign known answer (tests/test_processing.py:31-40)
    t1 = proc.design_channel_filter(1e6, 12_500.0, 10)
    nn = np.arange(0, int(1e6 * 0.1))
    warm = np.exp(1j * 2.0 * np.pi * 12_500.0 * nn / 1e6).astype(np.complex64)
    stage["mix_sign_pos_tone"] = np.int64(proc.choose_mix_sign(warm, 1e6, 12_500.0, t1, 10))
    stage["mix_sign_neg_tone"] = np.int64(proc.choose_mix_sign(np.conj(warm), 1e6, 12_500.0, t1, 10))
    # planning helpers
    stage["tune_chunk"] = np.asarray([[fs_, req, proc.tune_chunk_size(fs_, req)]
                                      for fs_ in (0.0, 250e3, 1e6, 2.5e6, 10e6, 20e6, 61.44e6)
                                      for req in (1, 65_536, 1_048_576, 8_000_000)], dtype=np.float64)
    stage["ntaps_table"] = np.asarray([[fs_, bw_, d_, len(proc.design_channel_filter(fs_, bw_, d_))]
                                       for fs_, bw_, d_ in ((2.5e6, 12_500.0, 26), (10e6, 12_500.0, 104),
                                                            (20e6, 10_000.0, 208), (20e6, 2_800.0, 208),
                                                            (61.44e6, 12_500.0, 640), (250e3, 200_000.0, 3))],
                                      dtype=np.float64)
    np.savez_compressed(HERE / "stage_vectors.npz", **stage)

    (HERE / "manifest.json").write_text(json.dumps(manifest, indent=1, default=float) + "\n")
    total = sum(p.stat().st_size for p in HERE.glob("*.npz"))
    print(f"wrote {len(list(HERE.glob('*.npz')))} fixtures, {total/1e6:.2f} MB")


if __name__ == "__main__":
    main()
