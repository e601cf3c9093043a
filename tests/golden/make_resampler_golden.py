#!/usr/bin/env python
"""Golden vectors for the 48 kHz output stage from a REAL libswresample (build container only).

The reference delegates this stage to ffmpeg (processing.py:399-418).  No ffmpeg binary exists in
the image, but opencv's wheel bundles libswresample 6.1.100 / libavutil 60.8.100 (FFmpeg 8 series);
this script drives it through ctypes exactly like ffmpeg's auto-inserted `aresample` does: one
context doing rate + format conversion, all-default options, several swr_convert calls, then a
flush.  Outputs go to tests/golden/resampler_vectors.npz.
"""
from __future__ import annotations

import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIBDIR = "/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs"
if LIBDIR not in os.environ.get("LD_LIBRARY_PATH", ""):
    os.environ["LD_LIBRARY_PATH"] = LIBDIR + ":" + os.environ.get("LD_LIBRARY_PATH", "")
    os.execv(sys.executable, [sys.executable] + sys.argv)

avutil = C.CDLL(LIBDIR + "/libavutil-ec54c519.so.60.8.100")
swr = C.CDLL(LIBDIR + "/libswresample-02b1114a.so.6.1.100")
FMT_S16, FMT_FLT = 1, 3


class AVChannelLayout(C.Structure):
    _fields_ = [("order", C.c_int), ("nb_channels", C.c_int), ("u", C.c_uint64), ("opaque", C.c_void_p)]


avutil.av_channel_layout_default.argtypes = [C.POINTER(AVChannelLayout), C.c_int]
swr.swr_alloc_set_opts2.argtypes = [C.POINTER(C.c_void_p), C.POINTER(AVChannelLayout), C.c_int, C.c_int,
                                    C.POINTER(AVChannelLayout), C.c_int, C.c_int, C.c_int, C.c_void_p]
swr.swr_init.argtypes = [C.c_void_p]
swr.swr_convert.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_void_p), C.c_int]
swr.swr_free.argtypes = [C.POINTER(C.c_void_p)]


def run_swr(x: np.ndarray, in_rate: int, out_rate: int, out_fmt: int, pieces: list[int]) -> list[np.ndarray]:
    lay = AVChannelLayout()
    avutil.av_channel_layout_default(C.byref(lay), 1)
    ctx = C.c_void_p()
    assert swr.swr_alloc_set_opts2(C.byref(ctx), C.byref(lay), out_fmt, out_rate, C.byref(lay), FMT_FLT, in_rate, 0, None) == 0
    assert swr.swr_init(ctx) == 0
    outs, pos = [], 0
    dt = np.int16 if out_fmt == FMT_S16 else np.float32
    for n in pieces + [0]:
        seg = np.ascontiguousarray(x[pos:pos + n], dtype=np.float32)
        pos += n
        cap = int(n * out_rate / in_rate) + 4096
        out = np.empty(cap, dtype=dt)
        op, ip = C.c_void_p(out.ctypes.data), C.c_void_p(seg.ctypes.data)
        got = swr.swr_convert(ctx, C.byref(op), cap, C.byref(ip) if n > 0 else None, n)
        assert got >= 0
        outs.append(out[:got].copy())
    swr.swr_free(C.byref(ctx))
    return outs


def main() -> None:
    rng = np.random.default_rng(77)
    vec = {}
    cases = []
    for in_rate, n, pieces in ((96_154, 40_330, [40_330]), (96_154, 100_991, [40_330, 40_330, 20_331]),
                               (96_000, 6_554 * 4 + 3, [6_554, 6_554, 6_554, 6_557]), (95_238, 20_000, [20_000]),
                               (100_000, 12_345, [12_345]), (96_154, 333, [333])):
        t = np.arange(n) / in_rate
        x = (0.4 * np.sin(2 * np.pi * 1_000.0 * t) + 0.25 * np.sin(2 * np.pi * 7_300.0 * t + 1.0)
             + 0.3 * rng.normal(size=n)).astype(np.float32)
        x = np.clip(x, -0.99, 0.99)                 # what AudioWriter.write hands to ffmpeg
        key = f"r{in_rate}_n{n}"
        f = run_swr(x, in_rate, 48_000, FMT_FLT, pieces)
        s = run_swr(x, in_rate, 48_000, FMT_S16, pieces)
        vec[key + "_in"] = x
        vec[key + "_flt"] = np.concatenate(f)
        vec[key + "_s16"] = np.concatenate(s)
        vec[key + "_counts"] = np.asarray([len(a) for a in s], dtype=np.int64)
        cases.append((in_rate, n, len(pieces)))
    # sample counts only, many lengths (the flush rule)
    rows = []
    for in_rate in (96_154, 95_238, 96_000, 89_286, 100_000):
        for n in list(range(1_000, 1_024)) + [5_000, 12_345, 20_000, 20_001, 20_002, 20_003, 40_330]:
            o = run_swr(np.zeros(n, dtype=np.float32), in_rate, 48_000, FMT_FLT, [n])
            rows.append((in_rate, n, len(o[0]), len(o[0]) + len(o[1])))
    vec["count_table"] = np.asarray(rows, dtype=np.int64)
    np.savez_compressed(HERE / "resampler_vectors.npz", **vec)
    print("cases", cases, "count rows", len(rows), "bytes", (HERE / "resampler_vectors.npz").stat().st_size)


if __name__ == "__main__":
    main()
