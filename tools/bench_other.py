#!/usr/bin/env python
"""One entry of bench.py's `other_workloads` block on its own (one GPU), e.g. under ncu:

    python tools/bench_other.py cfg5_c256 [--reps 3]
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("key")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--cpu", action="store_true")
    a = ap.parse_args()
    import torch
    import bench
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    peaks = ROOT / "MEASURED_PEAKS.json"
    peak = float(json.loads(peaks.read_text())["hbm_gbs"]) if peaks.exists() else 6650.0
    for key, desc, fs, secs, specs, _all in bench.other_workload_specs():
        if key == a.key:
            print(json.dumps(bench.run_other_workload(key, desc, fs, secs, specs, dev, 0, 1, peak, cpu=a.cpu, reps=a.reps)))
            return
    raise SystemExit(f"unknown workload {a.key}")


if __name__ == "__main__":
    main()
