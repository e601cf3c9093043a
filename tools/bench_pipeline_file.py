#!/usr/bin/env python
"""File -> 48 kHz WAV through the whole host pipeline (native reader with the pinned read ring, batched streaming,
native 48 kHz writer), capture and outputs on tmpfs: BASELINE configs[0] (2.5 MS/s, 5 s, one NFM target) and a prefix
of configs[1] (10 MS/s, 5 NFM targets in one pass).  One JSON line.

    python tools/bench_pipeline_file.py [--seconds2 12] [--dir /dev/shm/iq2a_bench]
"""
from __future__ import annotations

import argparse
import json
import shutil
import sys
import time
import wave
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def write_wav_from_device(path: Path, fs: float, seconds: float, carriers) -> int:
    import torch
    import bench
    n = int(fs * seconds)
    raw = bench.synth_capture_device(0, n, torch.device("cuda", 0), 77, fs, carriers).cpu().numpy()
    with wave.open(str(path), "wb") as w:
        w.setparams((2, 2, int(fs), 0, "NONE", "not compressed"))
        w.writeframes(raw.tobytes())
    return n


def run(cfg_kwargs: dict, n: int, fs: float, reps: int = 2) -> dict:
    from iq_to_audio_b200.pipeline import ProcessingConfig, ProcessingPipeline
    best, outs = 1e30, None
    for _ in range(reps):
        t0 = time.perf_counter()
        res = ProcessingPipeline(ProcessingConfig(**cfg_kwargs)).run_many()
        dt = time.perf_counter() - t0
        if dt < best:
            best, outs = dt, res
    sizes = [Path(r.output_path).stat().st_size for r in outs if getattr(r, "output_path", None)]
    return {"samples": n, "wall_s": best, "Msamples_per_s": n / best / 1e6, "x_realtime": n / fs / best,
            "outputs": len(outs), "output_bytes": sizes}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds2", type=float, default=12.0)
    ap.add_argument("--seconds3", type=float, default=60.0, help="a second, longer run of configs[1] (0 = skip)")
    ap.add_argument("--dir", default="/dev/shm/iq2a_bench")
    a = ap.parse_args()
    from iq_to_audio_b200.benchmark import write_synthetic_capture
    tmp = Path(a.dir)
    tmp.mkdir(parents=True, exist_ok=True)
    try:
        out = {}
        # ---- configs[0]: the repo's --benchmark capture ---------------------------------------------------
        fs1, center = 2.5e6, 100e6
        p1 = tmp / "cfg1.wav"
        n1 = write_synthetic_capture(p1, fs1, 5.0, 25_000.0)
        base = dict(center_freq=center, bandwidth=12_500.0, demod_mode="nfm", input_format="pcm_s16le", input_container="wav")
        run(dict(in_path=p1, target_freq=center + 25e3, output_path=tmp / "warm.wav", max_input_seconds=0.5, **base), n1, fs1, 1)
        out["cfg1"] = run(dict(in_path=p1, target_freq=center + 25e3, output_path=tmp / "cfg1_out.wav", **base), n1, fs1)
        # ---- configs[1], a prefix: 10 MS/s, five NFM targets in one pass ----------------------------------------
        import bench
        fs2 = bench.FS
        p2 = tmp / "cfg2.wav"
        n2 = write_wav_from_device(p2, fs2, a.seconds2, None)
        out["cfg2_prefix"] = run(dict(in_path=p2, target_freq=center + bench.OFFSETS[0],
                                      target_freqs=[center + o for o in bench.OFFSETS], output_path=tmp / "cfg2_out.wav", **base),
                                 n2, fs2)
        out["cfg2_prefix"]["seconds"] = a.seconds2
        if a.seconds3 > 0:
            p2.unlink()
            n3 = write_wav_from_device(p2, fs2, a.seconds3, None)
            out["cfg2_long"] = run(dict(in_path=p2, target_freq=center + bench.OFFSETS[0],
                                        target_freqs=[center + o for o in bench.OFFSETS], output_path=tmp / "cfg2_out.wav", **base),
                                   n3, fs2)
            out["cfg2_long"]["seconds"] = a.seconds3
        out["storage"] = str(tmp)
        print(json.dumps(out))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
