#!/usr/bin/env python
"""BASELINE configs[4]: wideband channelizer sweep at 61.44 MS/s, 1 ... 256 NFM channels, against the CPU port.

For each channel count the capture (int16, resident in HBM) goes through ChannelBank.process_resident (audio out on
the device); time is CUDA-event time over `--reps` passes.  The CPU column is the oracle port of the reference's
per-target loop on one host core, measured once on a bounded sample and scaled by the channel count (the reference
runs targets as sequential passes, cli.py:683-710).  Prints one JSON object.

    python tools/sweep_channels.py [--seconds 1.0] [--counts 1,2,4,8,16,32,64,128,256]
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=1.0)
    ap.add_argument("--counts", default="1,2,4,8,16,32,64,128,256")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--cpu-samples", type=int, default=1 << 21)
    a = ap.parse_args()
    import torch
    import bench
    from iq_to_audio_b200.bank import ChannelBank, Target
    from iq_to_audio_b200.processing import channel_decimation, design_channel_filter
    from oracle import iq_oracle as orc

    fs = 61.44e6
    dev = torch.device("cuda", 0)
    d, fs_ch = channel_decimation(fs, 96_000.0)
    taps = design_channel_filter(fs, 12_500.0, d)
    chunk = 4 << 20
    n = int(a.seconds * fs) // chunk * chunk
    fs_saved = bench.FS
    bench.FS = fs                                              # device synthesis at this rate
    raw = bench.synth_capture_device(0, n + d, dev, seed=2)
    bench.FS = fs_saved
    rows = []
    for c in [int(v) for v in a.counts.split(",")]:
        offs = [(-29.0 + 58.0 * (i + 0.5) / c) * 1e6 for i in range(c)]
        bank = ChannelBank(fs, d, [Target(o, taps, 1, "nfm", 300.0, True) for o in offs], codec="pcm_s16le",
                           iq_order="iq", ref_chunk=chunk, device=0)
        nr = bank.rows_in(0, n)
        audio = torch.empty((c, nr), dtype=torch.float32, device=dev)
        run = lambda: bank.process_resident(raw.data_ptr(), 0, n + d, 0, n, dev_audio=audio.data_ptr(), out_stride=nr)
        run(); run()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(a.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        rows.append({"channels": c, "ms": best, "input_Msamples_per_s": n / best / 1e3,
                     "channel_Msamples_per_s": c * n / best / 1e3, "x_realtime": n / fs / (best / 1e3),
                     "kernel_generation": bank.kernel_generation, "fft_size": bank.fft_size})
        bank.close()
        del audio
    # CPU port, one channel, bounded sample
    m = a.cpu_samples
    x = orc.order_iq(orc.unpack_interleaved(raw[: 2 * m].cpu().numpy(), "pcm_s16le"), "iq")
    t0 = time.perf_counter()
    orc.run_target(x, orc.TargetPlan(sample_rate=fs, freq_offset=1.0e6, mix_sign=1), 1 << 20)
    cpu = m / (time.perf_counter() - t0) / 1e6
    for r in rows:
        r["cpu_port_input_Msamples_per_s_1core"] = cpu / r["channels"]
        r["speedup_vs_1core"] = r["input_Msamples_per_s"] / (cpu / r["channels"])
    print(json.dumps({"workload": f"cfg5 sweep: 61.44 MS/s int16, {n} samples resident, D={d}, {len(taps)} taps, NFM",
                      "cpu_port_Msamples_per_s_per_channel_1core": cpu, "cpu_sample": m, "rows": rows}))


if __name__ == "__main__":
    main()
