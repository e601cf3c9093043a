"""GPU: BASELINE configs[1] at its full size (10 MS/s x 60 s = 6e8 int16 frames, 5 NFM targets).

The oracle cannot run 60 s in test time, so the checks are the ones that do not depend on size:
  * the single 60 s pass equals the same capture processed as 8 time shards with their own halo + warm-up
    (what bench.py --gpus 8 does), rows compared one to one;
  * two windows of the 60 s output, one of them across a reference-chunk boundary 52 s into the capture (byte
    offsets beyond 2^31), equal the CPU oracle started from the reference's own carried state at that chunk
    boundary (NCO phase table, decimator offset);
  * counts: rows out == multiples of D in [0, n).
"""
from __future__ import annotations

import numpy as np
import pytest

from oracle import iq_oracle as orc

pytestmark = pytest.mark.gpu

AUDIO_TOL = 1e-4          # north_star: 1e-4 full scale on the float audio
BB_TOL = 2e-6


def test_cfg2_full_size_shards_and_oracle_windows():
    import torch
    import bench
    from iq_to_audio_b200 import plan as planmod
    from iq_to_audio_b200.bank import ChannelBank
    from iq_to_audio_b200.sharding import plan_segments

    dev = torch.device("cuda", 0)
    fs, chunk = bench.FS, 4 << 20
    d, fs_ch, targets = bench.make_targets()
    n = int(fs * 60)
    raw = bench.synth_capture_device(0, n + d, dev, seed=1)
    assert raw.numel() * 2 > 2**31                                   # the point of this test
    with ChannelBank(fs, d, targets, codec="pcm_s16le", iq_order="iq", ref_chunk=chunk, device=0) as bank:
        assert bank.kernel_generation == 5
        rows = bank.rows_in(0, n)
        assert rows == orc.decimated_count(0, n, d)
        whole = torch.zeros((5, rows), dtype=torch.float32, device=dev)
        bb = torch.zeros((5, rows), dtype=torch.complex64, device=dev)
        k, _ = bank.process_resident(raw.data_ptr(), 0, n + d, 0, n, dev_audio=whole.data_ptr(),
                                     dev_baseband=bb.data_ptr(), out_stride=rows)
        assert k == rows
        peaks_whole = np.array(bank.peaks)

        # ---- 8 time shards ----
        parts = torch.zeros((5, rows), dtype=torch.float32, device=dev)
        segs = plan_segments(n, 8, chunk, d, bank.halo, ["nfm"] * 5)
        assert segs[0].begin == 0 and segs[-1].end == n and all(a.end == b.begin for a, b in zip(segs, segs[1:]))
        peaks = np.zeros(5)
        for sg in segs:
            view = raw[2 * sg.first_frame:]
            bank.reset()
            kk, _ = bank.process_resident(view.data_ptr(), sg.first_frame, n + d - sg.first_frame, sg.begin, sg.end,
                                          warmup_rows=sg.warmup_rows, dev_audio=parts[:, sg.row_begin:].data_ptr(),
                                          out_stride=rows)
            assert kk == sg.rows
            peaks = np.maximum(peaks, np.array(bank.peaks))
        diff = float((parts - whole).abs().max())
        assert diff <= 2e-6, diff                                     # de-emphasis warm-up residue only
        assert np.abs(peaks - peaks_whole).max() <= 2e-6

        # ---- oracle windows ----
        taps = np.asarray(targets[0].taps, dtype=np.float64)
        win = 600_000                                                 # input samples handed to the oracle per window
        # all five targets, each at the stream start and across a chunk boundary 52 s in (chunk 124 = 5.2e8 samples)
        for k0, ci in [(k, c) for k in (0, 124) for c in range(len(targets))]:
            s0 = k0 * chunk
            lo = s0 if k0 else 0
            if k0:
                lo = s0 + chunk - win // 2                            # window straddles the boundary s0+chunk
            x = orc.order_iq(orc.unpack_interleaved(raw[2 * lo:2 * (lo + win)].cpu().numpy(), "pcm_s16le"), "iq")
            t = targets[ci]
            nco = orc.NcoState.for_offset(t.freq_offset, fs)
            w_signed = t.mix_sign * nco.increment
            # reference carries the phase chunk by chunk (processing.py:295); inside a chunk it is phase + w*n
            tab = planmod.phase_table(w_signed, chunk, k0 + 3)
            fir = orc.FirState(taps, 65_536)
            dec = orc.DecimState(d, offset=lo % d)
            dem = orc.DemodState.create("nfm", fs_ch, deemph_us=bench.DEEMPH_US)
            pieces, pos = [], lo
            while pos < lo + win:
                kc = pos // chunk
                end = min(lo + win, (kc + 1) * chunk)
                nco.phase = float((tab[kc] + w_signed * (pos - kc * chunk)) % (2 * np.pi)) if pos != kc * chunk else float(tab[kc])
                pieces.append(orc.decimate(dec, orc.fir_overlap_save(fir, orc.nco_mix(nco, x[pos - lo:end - lo], t.mix_sign))))
                pos = end
            chan = np.concatenate(pieces)
            audio, _ = orc.demodulate(dem, chan)
            r_lo = orc.decimated_count(0, lo, d)
            got_bb = bb[ci, r_lo:r_lo + chan.size].cpu().numpy()
            got_audio = whole[ci, r_lo:r_lo + chan.size].cpu().numpy()
            skip_bb = 0 if lo == 0 else (len(taps) + d - 1) // d + 1  # FIR history was empty in the oracle
            skip_audio = skip_bb + (0 if lo == 0 else 700)            # plus de-emphasis settling (tau = 29 rows)
            assert np.abs(got_bb[skip_bb:] - chan[skip_bb:]).max() <= BB_TOL
            assert np.abs(got_audio[skip_audio:] - audio[skip_audio:]).max() <= AUDIO_TOL
