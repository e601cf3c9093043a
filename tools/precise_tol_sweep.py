#!/usr/bin/env python
"""How far from a float32 rounding boundary must a sample of the register-pass float64 filter (csrc/precise_fft.cu) be
for its rounding to be trusted?  Runs the bit-faithful SSB channels of a 20 MS/s capture with the direct form and with
the transform form at several repair tolerances (IQ2A_PRECISE_TOL) and counts the channel samples that differ.

    python tools/precise_tol_sweep.py [--seconds 2]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def child(seconds: float) -> None:
    import numpy as np
    import torch
    import bench
    from iq_to_audio_b200.bank import ChannelBank, Target
    from iq_to_audio_b200.processing import channel_decimation, design_channel_filter

    fs = 20e6
    dev = torch.device("cuda", 0)
    d, _ = channel_decimation(fs, 96_000.0)
    chunk = 8 << 20
    n = max(chunk, int(seconds * fs) // chunk * chunk)
    # a strong tone in one SSB channel, noise only in the other (the quiet channel is the hard case: small |s|)
    amp = float(os.environ.get("SWEEP_AMP", "0.2"))      # 0.2: the bench's level; 0.45: close to int16 full scale
    raw = bench.synth_capture_device(0, n + d, dev, 5, fs, [(-4.1e6, "usb", 900.0, amp), (2.3e6, "am", 700.0, amp)],
                                     noise=float(os.environ.get("SWEEP_NOISE", "0.02")))
    tg = [Target(-4.1e6, design_channel_filter(fs, 2_800.0, d), 1, "usb", 300.0, True),
          Target(6.2e6, design_channel_filter(fs, 2_800.0, d), 1, "lsb", 300.0, True)]
    bank = ChannelBank(fs, d, tg, codec="pcm_s16le", iq_order="iq", ref_chunk=chunk, device=0)
    rows = bank.rows_in(0, n)
    bb = torch.empty((2, rows), dtype=torch.complex64, device=dev)
    audio = torch.empty((2, rows), dtype=torch.float32, device=dev)
    go = lambda: bank.process_resident(raw.data_ptr(), 0, n + d, 0, n, dev_audio=audio.data_ptr(), dev_baseband=bb.data_ptr(), out_stride=rows)
    go()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); go(); e1.record()
    torch.cuda.synchronize()
    np.save(os.environ["SWEEP_OUT"], torch.view_as_real(bb).cpu().numpy())
    print(json.dumps({"ms": e0.elapsed_time(e1), "rows": rows, "rms": [float(bb[i].abs().pow(2).mean().sqrt()) for i in range(2)]}))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=2.0)
    ap.add_argument("--child", action="store_true")
    ap.add_argument("--noise", type=float, default=0.02, help="wideband noise per component (the quiet channel's level)")
    ap.add_argument("--tols", default="1e-12,1e-13,1e-14,3e-15,1e-15,0")
    ap.add_argument("--amp", type=float, default=0.2, help="amplitude of the two strong carriers (wideband level)")
    a = ap.parse_args()
    if not a.child:
        os.environ["SWEEP_NOISE"] = str(a.noise)
        os.environ["SWEEP_AMP"] = str(a.amp)
    if a.child:
        child(a.seconds)
        return
    import numpy as np
    out = {}
    ref = None
    for label, env in [("direct", {"IQ2A_PRECISE_FIR": "direct"})] + \
            [("default", {}) if t == "default" else (f"tol={t}", {"IQ2A_PRECISE_TOL": t, "IQ2A_PRECISE_TOL_REL": "0"})
             for t in a.tols.split(",")]:
        path = f"/tmp/sweep_{label.replace('=', '_')}.npy"
        r = subprocess.run([sys.executable, __file__, "--child", "--seconds", str(a.seconds)], capture_output=True, text=True,
                           env={**os.environ, **env, "SWEEP_OUT": path})
        if r.returncode != 0:
            out[label] = {"error": r.stderr[-300:]}
            continue
        info = json.loads(r.stdout.strip().splitlines()[-1])
        bb = np.load(path)
        if ref is None:
            ref = bb
        info["samples_differing_from_direct"] = [int(np.any(bb[i] != ref[i], axis=-1).sum()) for i in range(2)]
        out[label] = info
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
