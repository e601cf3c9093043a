#!/usr/bin/env python
"""Build a variant of the library with extra -D flags on the mirror-pair kernel, for A/B timing on the GPU box.

    python tools/build_variant.py NAME -DIQ2A_SOMETHING [-D...]      ->  _ab/libiq2a_NAME.so
    IQ2A_LIB=_ab/libiq2a_NAME.so python bench.py ...

Only `channelizer5_inst.cu` (every group size) is recompiled with the flags; the other objects are the ones of the
regular build (run `python -m iq_to_audio_b200.build` first).  The variants are git-ignored (*.so) but travel with gpurun.
"""
from __future__ import annotations

import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from iq_to_audio_b200 import build as B  # noqa: E402


def main() -> None:
    name, flags = sys.argv[1], sys.argv[2:]
    B.build()
    out = ROOT / "_ab"
    out.mkdir(exist_ok=True)
    nvcc = B._nvcc()

    def one(cg: int) -> Path:
        o = out / f"c5_{name}_{cg}.o"
        cmd = [nvcc, *B.NVCC_FLAGS, *flags, f"-DIQ2A_CG={cg}", "-c", str(B.CSRC / "channelizer5_inst.cu"), "-o", str(o)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise SystemExit(r.stderr[-3000:])
        if cg == 5:
            for line in r.stderr.splitlines():
                if "spill" in line or "Used" in line:
                    print(line.strip())
        return o

    with ThreadPoolExecutor(max_workers=6) as ex:
        mine = list(ex.map(one, B.GROUP_SIZES))
    others = [o for o in sorted(B.OBJ.glob("*.o")) if not o.name.startswith("channelizer5_")]
    lib = out / f"libiq2a_{name}.so"
    r = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(lib), *map(str, mine), *map(str, others)],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise SystemExit(r.stderr[-3000:])
    print(lib)


if __name__ == "__main__":
    main()
