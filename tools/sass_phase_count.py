#!/usr/bin/env python
"""Static instruction mix of the channel-bank kernel, per barrier-delimited phase, from the built object.

    python tools/sass_phase_count.py [--cg 5] [--bt 2] [--stg 0]

Runs `cuobjdump -sass` on `iq_to_audio_b200/csrc/_obj/channelizer2_<cg>.o` (no GPU needed), cuts the SASS of
`k_channelize2<cg, bt, stg>` at its `BAR.SYNC`s and prints the opcode histogram of every segment: what was read before
GPU time was spent on a variant (registers / spills come from the `.ptxas.log` next to the object).  Both branches of
a run-time `if` are counted (e.g. the two IQ orders of the unpack, the full-tile and the ragged multiply-accumulate).
"""
from __future__ import annotations

import argparse
import collections
import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
OBJ = ROOT / "iq_to_audio_b200" / "csrc" / "_obj"


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--cg", type=int, default=5)
    ap.add_argument("--bt", type=int, default=2)
    ap.add_argument("--stg", type=int, default=0)
    a = ap.parse_args()
    obj = OBJ / f"channelizer2_{a.cg}.o"
    sass = subprocess.run(["cuobjdump", "-sass", str(obj)], capture_output=True, text=True, check=True).stdout
    want = f"k_channelize2ILi{a.cg}ELi{a.bt}ELi{a.stg}E"
    inside, ops = False, []
    for line in sass.splitlines():
        if "Function :" in line:
            inside = want in line
            continue
        if not inside:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            ops.append(m.group(1))
    if not ops:
        raise SystemExit(f"{want} not found in {obj}")
    log = (OBJ / f"channelizer2_{a.cg}.ptxas.log").read_text().splitlines()
    for i, line in enumerate(log):
        if want in line and "Function properties" in line:
            print(log[i + 1].strip())
            print(log[i + 2].strip())
            break
    segs, cur = [], collections.Counter()
    for op in ops:
        cur[op.split(".")[0]] += 1
        if op.startswith("BAR"):
            segs.append(cur)
            cur = collections.Counter()
    segs.append(cur)
    print(f"{want}: {len(ops)} instructions, {len(segs)} segments")
    for i, c in enumerate(segs):
        top = ", ".join(f"{k} {v}" for k, v in c.most_common(12))
        print(f"  segment {i:2d}: {sum(c.values()):5d}   {top}")


if __name__ == "__main__":
    main()
