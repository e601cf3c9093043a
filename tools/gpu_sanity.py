"""Scratch GPU sanity run: prints max errors of the CUDA path vs the golden fixtures."""
import sys, time, traceback
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
from oracle import iq_oracle as orc
from tests import _cases
from iq_to_audio_b200 import plan as P
from iq_to_audio_b200.bank import ChannelBank, Target
import iq_to_audio_b200.processing as gp
import iq_to_audio_b200.decoders as gd


def section(name):
    print(f"\n=== {name}", flush=True)


def guarded(fn):
    try:
        fn()
    except Exception:
        traceback.print_exc()


def t_gtable():
    section("G table vs numpy plan")
    taps = orc.channel_taps(2.5e6, 12500.0, 26)
    for M in (512, 1024):
        pl = P.build_plan(2.5e6, 26, [P.ChannelSpec(25e3, taps, 1), P.ChannelSpec(-300e3, taps, -1)], m_fft=M)
        with ChannelBank(2.5e6, 26, [Target(25e3, taps, 1), Target(-300e3, taps, -1)], fft_size=M) as b:
            g = b.g_table()
            print(M, "vd", b.overlap_rows, pl.vd, "max |dG|", np.abs(g - pl.g_table).max(), "max|G|", np.abs(pl.g_table).max())


def t_stage():
    section("stage-level")
    sv = _cases.load("stage_vectors")
    z = sv["mix_in"]
    osc = gp.ComplexOscillator(123_456.7, 2.4e6)
    a = osc.mix(z[:17000], -1); pa = osc.phase
    b = osc.mix(z[17000:], -1)
    print("mixer err", np.abs(a - sv["mix_out_a"]).max(), np.abs(b - sv["mix_out_b"]).max(), "phase eq", pa == float(sv["mix_phase_a"]), osc.phase == float(sv["mix_phase_b"]))
    fir = gp.OverlapSaveFIR(sv["fir_taps"], 4096)
    out = np.concatenate([fir.process(z[:5000]), fir.process(z[5000:5700]), fir.process(z[5700:])])
    print("fir err", np.abs(out - sv["fir_out"]).max(), "state eq", np.array_equal(fir.state, sv["fir_state"]))
    d7 = gp.Decimator(7)
    got = np.concatenate([d7.process(z[:10]), d7.process(z[10:11]), d7.process(z[11:400])])
    print("dec7 eq", np.array_equal(got, sv["dec7"]), d7.offset == int(sv["dec7_offset"]))
    qd = gd.nfm.QuadratureDemod(); de = gd.nfm.DeemphasisFilter(300.0, 96153.846)
    a1 = qd.process(z[:9000]); a2 = qd.process(z[9000:20000])
    print("disc err", np.abs(np.concatenate([a1, a2]) - sv["disc"]).max())
    y = np.concatenate([de.process(sv["disc"][:9000]), de.process(sv["disc"][9000:])])
    print("deemph err", np.abs(y - sv["deemph"]).max(), "state", de.state, float(sv["deemph_state"]))
    dc = gd.common.DCBlocker()
    r = sv["dc_in"]
    y = np.concatenate([dc.process(r[:2500]), dc.process(r[2500:])])
    print("dc err", np.abs(y - sv["dc_out"]).max())
    from iq_to_audio_b200.decoders.base import run_scan
    from iq_to_audio_b200 import _lib
    g = np.concatenate([run_scan(2, 0.0, sv["dc_out"][:2500], _lib.ChannelState.fresh()), run_scan(2, 0.0, sv["dc_out"][2500:], _lib.ChannelState.fresh())])
    print("agc err", np.abs(g - sv["agc_out"]).max(), "rel", (np.abs(g - sv["agc_out"]) / np.maximum(np.abs(sv["agc_out"]), 1e-3)).max())
    taps = orc.channel_taps(1e6, 12500.0, 10)
    nn = np.arange(0, int(1e6 * 0.1))
    warm = np.exp(1j * 2.0 * np.pi * 12500.0 * nn / 1e6).astype(np.complex64)
    print("mix sign", gp.choose_mix_sign(warm, 1e6, 12500.0, taps, 10), gp.choose_mix_sign(np.conj(warm), 1e6, 12500.0, taps, 10))


def run_stream(key, names, fft_size=0):
    m = _cases.manifest()[key]
    raw = _cases.raw_input(key)
    fs = m["fs"]
    D, fs_ch = orc.plan_decimation(fs, 96000.0)
    tg = []
    gold = [_cases.load(n) for n in names]
    for t, g in zip(m["targets"], gold):
        bw = t.get("bw", 12500.0 if t["mode"] == "nfm" else 2800.0)
        tg.append(Target(t["f_off"], orc.channel_taps(fs, bw, D), int(g["mix_sign"]), t["mode"], 300.0, t.get("agc", True)))
    chunk = m["chunk"]
    fb = orc.FRAME_BYTES[m["codec"]]
    rawb = raw.view(np.uint8)
    limit = m.get("max_input_samples")
    with ChannelBank(fs, D, tg, codec=m["codec"], iq_order=m["iq_order"], ref_chunk=chunk, fft_size=fft_size) as bank:
        print(key, "M", bank.fft_size, "vd", bank.overlap_rows)
        audio, clip, bb, cnt, rms = [], [], [], [], []
        nfr = rawb.size // fb
        if limit: nfr = min(nfr, limit)
        t0 = time.perf_counter()
        for s in range(0, nfr, chunk):
            e = min(s + chunk, nfr)
            r = bank.process_chunk(rawb[s * fb:e * fb], want_baseband=True)
            audio.append(r.audio.copy()); clip.append(r.clipped.copy()); bb.append(r.baseband.copy()); cnt.append(r.count); rms.append(r.rms_dbfs.copy())
        dt = time.perf_counter() - t0
        peaks = bank.peaks
    audio = np.concatenate(audio, axis=1); clip = np.concatenate(clip, axis=1); bb = np.concatenate(bb, axis=1)
    rms = np.array(rms)
    for i, g in enumerate(gold):
        ea = np.abs(audio[i] - g["audio"]); ec = np.abs(clip[i] - g["clipped"]); eb = np.abs(bb[i] - g["baseband"])
        print(f"  ch{i} counts_eq {cnt == list(g['counts'])} audio max {ea.max():.3e} (arg {ea.argmax()}) clipped max {ec.max():.3e} bb max {eb.max():.3e} "
              f"rms err {np.abs(rms[:, i] - g['rms_dbfs']).max():.2e} peak {peaks[i]:.6f} vs {float(g['peak']):.6f}")
    print(f"  streaming {nfr/dt/1e6:.1f} MS/s wall")


def t_stream():
    section("fused streaming vs golden")
    guarded(lambda: run_stream("case_a_nfm_2p5M", ["case_a_nfm_2p5M"]))
    guarded(lambda: run_stream("case_a_nfm_2p5M", ["case_a_nfm_2p5M"], 1024))
    guarded(lambda: run_stream("case_b_nfm_10M", [f"case_b_nfm_10M_t{i}" for i in range(5)]))
    guarded(lambda: run_stream("case_b_nfm_10M", [f"case_b_nfm_10M_t{i}" for i in range(5)], 1024))
    guarded(lambda: run_stream("case_c_20M_am_ssb", ["case_c_20M_am", "case_c_20M_usb", "case_c_20M_lsb"]))
    for n in ("case_d_pcm_u8_qi_nfm", "case_d_pcm_f32le_iq_inv_nfm", "case_d_pcm_s16le_qi_inv_usb"):
        guarded(lambda n=n: run_stream(n, [n]))
    guarded(lambda: run_stream("case_e_truncated", ["case_e_truncated"]))


def t_resident():
    section("resident whole-capture vs golden + timing")
    import torch
    key = "case_b_nfm_10M"
    m = _cases.manifest()[key]
    raw = _cases.raw_input(key)
    fs = m["fs"]; D, _ = orc.plan_decimation(fs, 96000.0)
    gold = [_cases.load(f"case_b_nfm_10M_t{i}") for i in range(5)]
    tg = [Target(t["f_off"], orc.channel_taps(fs, 12500.0, D), 1, "nfm") for t in m["targets"]]
    d_raw = torch.from_numpy(raw.copy()).cuda()
    n = raw.size // 2
    with ChannelBank(fs, D, tg, ref_chunk=m["chunk"]) as bank:
        rows = bank.rows_in(0, n)
        d_audio = torch.empty((5, rows), dtype=torch.float32, device="cuda")
        d_clip = torch.empty((5, rows), dtype=torch.float32, device="cuda")
        k, rms = bank.process_resident(d_raw.data_ptr(), 0, n, 0, n, dev_audio=d_audio.data_ptr(), dev_clipped=d_clip.data_ptr(), out_stride=rows, want_rms=True)
        a = d_audio.cpu().numpy()
        for i, g in enumerate(gold):
            print(f"  ch{i} rows {k} audio max err {np.abs(a[i] - g['audio']).max():.3e} rms err {np.abs(rms[i] - g['rms_dbfs']).max():.2e}")
        # split into two segments with warm-up
        half = (n // 2 // m["chunk"]) * m["chunk"]
        r0 = bank.rows_in(0, half)
        bank.process_resident(d_raw.data_ptr(), 0, n, 0, half, dev_audio=d_audio.data_ptr(), out_stride=rows)
        a0 = d_audio.cpu().numpy()[:, :r0].copy()
        bank.process_resident(d_raw.data_ptr(), 0, n, half, n, warmup_rows=600, dev_audio=d_audio.data_ptr(), out_stride=rows)
        a1 = d_audio.cpu().numpy()[:, :rows - r0]
        a2 = np.concatenate([a0, a1], axis=1)
        for i, g in enumerate(gold):
            print(f"  ch{i} 2-segment audio max err {np.abs(a2[i] - g['audio']).max():.3e}")
    # timing at a larger size: tile the capture to 64 M frames
    reps = 64
    big = torch.from_numpy(np.tile(raw, reps)).cuda()
    nb = big.numel() // 2
    with ChannelBank(fs, D, tg, ref_chunk=4194304) as bank:
        rows = bank.rows_in(0, nb)
        d_audio = torch.empty((5, rows), dtype=torch.float32, device="cuda")
        for it in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            bank.process_resident(big.data_ptr(), 0, nb, 0, nb, dev_audio=d_audio.data_ptr(), out_stride=rows)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            print(f"  resident {nb/1e6:.0f} M frames, 5 ch: {dt*1e3:.2f} ms -> {nb/dt/1e9:.2f} GS/s")
    with ChannelBank(fs, D, tg[:1], ref_chunk=4194304) as bank:
        d_audio = torch.empty((1, rows), dtype=torch.float32, device="cuda")
        for it in range(2):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            bank.process_resident(big.data_ptr(), 0, nb, 0, nb, dev_audio=d_audio.data_ptr(), out_stride=rows)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            print(f"  resident {nb/1e6:.0f} M frames, 1 ch: {dt*1e3:.2f} ms -> {nb/dt/1e9:.2f} GS/s")


if __name__ == "__main__":
    for f in (t_gtable, t_stage, t_stream, t_resident):
        guarded(f)
