"""CPU-only: the committed profile that bench.py's `roofline.traffic` is scaled from must belong to the kernel the
cfg2 bank actually selects (round-1 verdict: the figure went stale silently when the kernel changed)."""
from __future__ import annotations

import json
from pathlib import Path

import bench
from iq_to_audio_b200 import plan as P
from iq_to_audio_b200.processing import channel_decimation, design_channel_filter

ROOT = Path(__file__).resolve().parents[1]


def test_traffic_profile_names_the_selected_kernel():
    prof = json.loads((ROOT / "profiles" / "channelize_traffic.json").read_text())
    d, _ = channel_decimation(bench.FS, 96_000.0)
    taps = design_channel_filter(bench.FS, bench.BW, d)
    # the selection rule of iq2a_bank_create for int16 input: symmetric taps of one length, D % 4 == 0 and
    # (ntaps - 1) % 4 == 0 -> the mirror-pair kernel (generation 5) with one instantiation per channel count
    P.pair_tiles(len(taps), d)                                   # raises if generation 5 could not take cfg2
    assert (taps == taps[::-1]).all()
    want = f"k_channelize5<{len(bench.OFFSETS)}>"
    assert prof["kernel"].startswith(want), (prof["kernel"], want)
    assert 4.0 < prof["dram_bytes_per_sample"] < 6.0 and prof["samples"] > 1e7
    src = ROOT / prof["source"].split(" ")[0]
    assert src.exists(), f"{src} (the capture the figure comes from) is not committed"
