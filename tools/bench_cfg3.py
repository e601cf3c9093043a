#!/usr/bin/env python
"""BASELINE configs[2]: 20 MS/s capture, AM + USB + LSB targets, AGC on -- resident throughput of the bank with the
bit-faithful SSB path (default) and with the float64-scan AGC (IQ2A_PRECISE_SSB=0).  One JSON line.

    python tools/bench_cfg3.py [--seconds 10]
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


def run(seconds: float) -> dict:
    import torch
    import bench
    from iq_to_audio_b200.bank import ChannelBank, Target
    from iq_to_audio_b200.processing import channel_decimation, design_channel_filter

    fs = 20e6
    dev = torch.device("cuda", 0)
    d, fs_ch = channel_decimation(fs, 96_000.0)
    chunk = 8 << 20                                           # tune_chunk_size(20 MS/s)
    n = int(seconds * fs) // chunk * chunk
    saved = bench.FS
    bench.FS = fs
    raw = bench.synth_capture_device(0, n + d, dev, seed=3)
    bench.FS = saved
    tg = [Target(2.3e6, design_channel_filter(fs, 10_000.0, d), 1, "am", 300.0, True),
          Target(-4.1e6, design_channel_filter(fs, 2_800.0, d), 1, "usb", 300.0, True),
          Target(6.2e6, design_channel_filter(fs, 2_800.0, d), 1, "lsb", 300.0, True)]
    bank = ChannelBank(fs, d, tg, codec="pcm_s16le", iq_order="iq", ref_chunk=chunk, device=0)
    rows = bank.rows_in(0, n)
    audio = torch.empty((3, rows), dtype=torch.float32, device=dev)
    go = lambda: bank.process_resident(raw.data_ptr(), 0, n + d, 0, n, dev_audio=audio.data_ptr(), out_stride=rows)
    go()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); go(); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return {"samples": n, "ms": best, "Msamples_per_s": n / best / 1e3, "x_realtime": n / fs / (best / 1e3),
            "decimation": d, "taps": [len(t.taps) for t in tg], "fft_size": bank.fft_size}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--child", action="store_true")
    a = ap.parse_args()
    if a.child:
        print(json.dumps(run(a.seconds)))
        return
    out = {"workload": "cfg3 shape: 20 MS/s int16, AM + USB + LSB targets, AGC on, resident in HBM"}
    for label, env in (("bit_faithful_ssb", {}), ("float64_scan_agc", {"IQ2A_PRECISE_SSB": "0"})):
        r = subprocess.run([sys.executable, __file__, "--child", "--seconds", str(a.seconds)], capture_output=True,
                           text=True, env={**os.environ, **env})
        out[label] = json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else {"error": r.stderr[-400:]}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
