#!/usr/bin/env python
"""cuFFT timed as a COMPARISON only (never on the product path): the cfg2 channelizer built from library calls.

Two library formulations of the same work on the same device-resident int16 capture, through torch.fft (cuFFT)
and torch.einsum/bmm (cuBLAS):

  direct     the reference's own order of operations moved to the GPU: per target, unpack -> NCO mix ->
             overlap-save FFT filter (65 536-sample blocks, 131 072-point cuFFT) -> keep every D-th sample ->
             discriminator.  This is what a straight port of processing.py:1070-1154 would run.
  polyphase  the formulation the hand-written kernel uses (shared 512-point forward transforms of the D
             polyphase branches, per-bin multiply-accumulate with the branch filters, 512-point inverse per
             target), but with cuFFT/cuBLAS kernels and the intermediate spectra in HBM.

Both are checked against the product kernel on the measured segment (max abs error on the channel samples), then
timed with CUDA events; prints one JSON line.  Usage: python tools/cufft_compare.py [--msamples 64]
"""
from __future__ import annotations

import argparse
import json
import math
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402  (workload constants + device capture synthesis)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--msamples", type=int, default=64)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    import torch
    from iq_to_audio_b200 import plan as planmod
    from iq_to_audio_b200.bank import ChannelBank

    dev = torch.device("cuda", 0)
    fs = bench.FS
    d, fs_ch, targets = bench.make_targets()
    taps = np.asarray(targets[0].taps, dtype=np.float64)
    ntaps = taps.size
    chunk = 4 << 20
    bank = ChannelBank(fs, d, targets, codec="pcm_s16le", iq_order="iq", ref_chunk=chunk, device=0)
    m_fft, vd, ld = bank.fft_size, bank.overlap_rows, bank.rows_per_block
    n = (a.msamples << 20) // (ld * d) * (ld * d)          # whole overlap-save blocks
    raw = bench.synth_capture_device(0, n + d, dev, seed=5)
    x = (raw[: 2 * n].view(n, 2).float() * (1.0 / 32768.0))
    x = torch.view_as_complex(x.contiguous())
    ws = [-2.0 * math.pi * t.freq_offset / fs for t in targets]     # processing.py:287,:293 with mix_sign +1: LO = exp(+j w n)

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(a.reps):
            e0, e1 = ev(), ev()
            e0.record()
            out = fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best, out

    # ---- product kernel on the same segment (channel samples only) --------------------------------
    rows = n // d
    bb = torch.empty((len(targets), rows + 8), dtype=torch.complex64, device=dev)

    def product():
        bank.reset()
        bank.process_resident(raw.data_ptr(), 0, n + d, 0, n, dev_baseband=bb.data_ptr(), out_stride=rows + 8)
        return bb[:, :rows]

    ms_prod, ref = timed(product)
    ref = ref.clone()

    # ---- direct: mix, 131072-point overlap-save, decimate, per target ------------------------------
    nfft = 1 << 17
    hop = nfft - (ntaps - 1)
    h_f = torch.fft.fft(torch.tensor(taps, dtype=torch.float32, device=dev).to(torch.complex64), n=nfft)
    seg = 8 << 20                                            # samples mixed at a time (bounds the temporaries)

    def direct():
        outs = []
        for w in ws:
            parts = []
            hist = torch.zeros(ntaps - 1, dtype=torch.complex64, device=dev)
            for s in range(0, n, seg):
                m = min(seg, n - s)
                ph = torch.remainder(w * torch.arange(s, s + m, device=dev, dtype=torch.float64), 2.0 * math.pi)
                mixed = x[s:s + m] * torch.polar(torch.ones_like(ph), ph).to(torch.complex64)
                buf = torch.cat((hist, mixed))
                nblk = (m + hop - 1) // hop
                pad = nblk * hop + ntaps - 1 - buf.numel()
                if pad > 0:
                    buf = torch.cat((buf, torch.zeros(pad, dtype=torch.complex64, device=dev)))
                frames = buf.as_strided((nblk, nfft), (hop, 1))
                y = torch.fft.ifft(torch.fft.fft(frames, dim=1) * h_f, dim=1)[:, ntaps - 1:].reshape(-1)[:m]
                first = (-s) % d
                parts.append(y[first::d])
                hist = buf[m:m + ntaps - 1].clone()
            outs.append(torch.cat(parts))
        return torch.stack(outs)

    ms_direct, out_direct = timed(direct)
    err_direct = float((out_direct[:, vd:rows] - ref[:, vd:rows]).abs().max())

    # ---- polyphase with library kernels -----------------------------------------------------------
    g = np.stack([np.fft.fft(planmod.branch_filters(taps, w, d, m_fft), axis=1) for w in ws])          # [C, D, M]
    g_t = torch.tensor(g.astype(np.complex64), device=dev)
    xr = torch.cat((torch.zeros(vd * d, dtype=torch.complex64, device=dev), x,
                    torch.zeros(m_fft * d, dtype=torch.complex64, device=dev)))
    nblk = n // (ld * d)
    rot = torch.stack([torch.polar(torch.ones(rows, device=dev, dtype=torch.float64),
                                   torch.remainder(w * d * torch.arange(rows, device=dev, dtype=torch.float64),
                                                   2 * math.pi)).to(torch.complex64) for w in ws])
    bsz = 256                                                # blocks per batch (spectra: bsz*M*D*8 B = 109 MB)

    def polyphase():
        outs = []
        for b0 in range(0, nblk, bsz):
            nb = min(bsz, nblk - b0)
            # block b: rows [b*ld - vd, b*ld - vd + M) of the [rows, D] view x_p[m] = x[mD + p]; outputs b*ld + [0, ld)
            v = xr[b0 * ld * d:].as_strided((nb, m_fft, d), (ld * d, d, 1))
            spec = torch.fft.fft(v, dim=1)                                            # [nb, M, D]
            y = torch.einsum("cpk,bkp->bck", g_t, spec)                               # [nb, C, M]
            t = torch.fft.ifft(y, dim=2)[:, :, vd:]
            outs.append(t.permute(1, 0, 2).reshape(len(ws), nb * ld))
        return torch.cat(outs, dim=1) * rot[:, : nblk * ld]

    try:
        ms_poly, out_poly = timed(polyphase)
        err_poly = float((out_poly[:, vd:rows] - ref[:, vd:rows]).abs().max())
    except Exception as exc:                                   # keep the direct comparison even if this one fails
        ms_poly, err_poly = None, repr(exc)

    print(json.dumps({
        "workload": "cfg2 shape (10 MS/s int16, 5 NFM targets, D=104), channel samples only, input resident in HBM",
        "samples": n, "reps": a.reps,
        "product_ms": ms_prod, "product_Msamples_per_s": n / ms_prod / 1e3,
        "cufft_direct_ms": ms_direct, "cufft_direct_Msamples_per_s": n / ms_direct / 1e3,
        "cufft_direct_max_abs_diff": err_direct,
        "cufft_polyphase_ms": ms_poly, "cufft_polyphase_Msamples_per_s": (n / ms_poly / 1e3) if ms_poly else None,
        "cufft_polyphase_max_abs_diff": err_poly,
        "note": "library formulations are comparisons only; nothing in iq_to_audio_b200 calls cuFFT or cuBLAS"}))


if __name__ == "__main__":
    main()
