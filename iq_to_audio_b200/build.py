"""Build libiq2a_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension).

The .so is git-ignored but travels to the GPU box with the repo snapshot.  Objects are
compiled in parallel (one nvcc per translation unit) and only when a source is newer.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OBJ = CSRC / "_obj"
LIB = PKG / "libiq2a_b200.so"
SOURCES = ["cabi.cu", "channelizer.cu", "channelizer5s_inst.cu", "tail.cu", "stage.cu", "precise.cu", "precise_fft.cu", "resample.cu", "spectrum.cu"]
FFT_SIZES = (512, 1024)
GROUP_SIZES = (1, 2, 3, 4, 5, 6)
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(exe).exists():
        raise RuntimeError("nvcc not found; the B200 path cannot be built")
    return exe


def _stale(target: Path, deps: list[Path]) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    nvcc = _nvcc()
    OBJ.mkdir(exist_ok=True)
    headers = sorted(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "iq2a_b200.h"]
    jobs = []
    objs = []
    for src in SOURCES:
        s = CSRC / src
        o = OBJ / (s.stem + ".o")
        objs.append(o)
        if force or _stale(o, [s] + headers):
            jobs.append((s, o, []))
    inst = CSRC / "channelizer_inst.cu"
    for m in FFT_SIZES:
        for cg in GROUP_SIZES:
            o = OBJ / f"channelizer_{m}_{cg}.o"
            objs.append(o)
            if force or _stale(o, [inst] + headers):
                jobs.append((inst, o, [f"-DIQ2A_M={m}", f"-DIQ2A_CG={cg}"]))
    inst2 = CSRC / "channelizer2_inst.cu"
    for cg in GROUP_SIZES:
        o = OBJ / f"channelizer2_{cg}.o"
        objs.append(o)
        if force or _stale(o, [inst2] + headers):
            jobs.append((inst2, o, [f"-DIQ2A_CG={cg}"]))
    inst5 = CSRC / "channelizer5_inst.cu"
    for cg in GROUP_SIZES:
        o = OBJ / f"channelizer5_{cg}.o"
        objs.append(o)
        if force or _stale(o, [inst5] + headers):
            jobs.append((inst5, o, [f"-DIQ2A_CG={cg}"]))
    # biggest kernels first so the pool drains evenly
    jobs.sort(key=lambda j: -len(j[2]))

    def compile_one(job):
        s, o, defs = job
        cmd = [nvcc, *NVCC_FLAGS, *defs, "-c", str(s), "-o", str(o)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        (OBJ / (o.stem + ".ptxas.log")).write_text(r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {o.name}:\n{r.stderr[-4000:]}")
        return o.name

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
            for name in ex.map(compile_one, jobs):
                if verbose:
                    print(f"compiled {name}", file=sys.stderr)
    if force or jobs or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *map(str, objs)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stderr[-4000:]}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
