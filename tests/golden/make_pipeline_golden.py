#!/usr/bin/env python
"""Golden output of the UNMODIFIED reference pipeline: ``run_benchmark -> ProcessingPipeline.run``.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_pipeline_golden.py

The reference needs ``soundfile`` and an ``ffmpeg`` binary, neither of which is in this image.  Both are replaced
by shims that do what the real ones do on this path and nothing else:

* ``soundfile``: ``info()`` raises (the reference then reads the WAV header with the standard library,
  probe.py:85-101), ``write()`` stores PCM_16 like libsndfile does for float input (pcm.c ``f2s_array``: ``lrintf(x * 32767.0f)``) --
  used once, by ``benchmark._generate_synthetic_iq`` (benchmark.py:38).
* ``ffmpeg`` (``IQ_TO_AUDIO_FFMPEG``, utils.py:130): a script that implements the two invocations of the path --
  decode (processing.py:143-158: ``-ignore_length 1 -i capture.wav -f f32le -ac 2 -`` = every int16 of the data
  chunk divided by 32768, to stdout) and encode (processing.py:399-418: float32 mono on stdin): the encode shim
  writes the bytes it receives, unchanged, to ``$IQ2A_CAPTURE_DIR`` -- that stream is the exact float32 audio the
  reference hands to its encoder and the primary parity target.

Stored in ``tests/golden/pipeline_cfg1.npz``: the stream, the ``-ar`` rate the reference asked the encoder for and
the SHA-256 of the capture's PCM payload (the GPU test regenerates the capture and checks it is the same input).
Nothing from the reference is copied into the repository: only its outputs.
"""
from __future__ import annotations

import hashlib
import logging
import os
import stat
import sys
import tempfile
import types
import wave
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF_SRC = Path("/root/reference/src")

FAKE_FFMPEG = r'''#!{python}
import os, sys, wave
import numpy as np
a = sys.argv[1:]
if a[-1] == "-":                                   # decode: capture -> f32le stereo on stdout
    src = a[a.index("-i") + 1]
    with open(src, "rb") as f:
        blob = f.read()
    pos = blob.index(b"data") + 8                  # -ignore_length 1: everything after the data chunk header
    pcm = np.frombuffer(blob[pos:pos + (len(blob) - pos) // 4 * 4], dtype="<i2")
    out = sys.stdout.buffer
    for lo in range(0, pcm.size, 1 << 22):
        out.write((pcm[lo:lo + (1 << 22)].astype(np.float32) / np.float32(32768.0)).tobytes())
    out.flush()
else:                                              # encode: f32le mono on stdin -> keep the stream
    data = sys.stdin.buffer.read()
    dst = a[-1]
    rate = a[a.index("-ar") + 1]
    cap = os.environ["IQ2A_CAPTURE_DIR"]
    with open(os.path.join(cap, "encoder_stdin.f32"), "wb") as f:
        f.write(data)
    with open(os.path.join(cap, "encoder_args.txt"), "w") as f:
        f.write(rate + "\n" + " ".join(a) + "\n")
    with wave.open(dst, "wb") as w:                # something valid at the output path
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(48000)
        w.writeframes(b"")
'''


def _install_shims(tmp: Path) -> None:
    stub = types.ModuleType("soundfile")

    def info(*a, **k):
        raise RuntimeError("soundfile stub: no header probing")

    def write(path, data, samplerate, format=None, subtype=None):       # noqa: A002
        assert subtype == "PCM_16" and data.ndim == 2
        # libsndfile pcm.c f2s_array: lrintf(src[i] * normfact) with float normfact = 0x7FFF, i.e. float32 arithmetic
        pcm = np.clip(np.rint(np.asarray(data, dtype=np.float32) * np.float32(32767.0)), -32768, 32767).astype("<i2")
        with wave.open(str(path), "wb") as w:
            w.setnchannels(data.shape[1]); w.setsampwidth(2); w.setframerate(int(samplerate))
            w.writeframes(pcm.tobytes())

    stub.info = info
    stub.write = write
    stub.read = info
    stub.SoundFile = object
    sys.modules["soundfile"] = stub
    exe = tmp / "ffmpeg"
    exe.write_text(FAKE_FFMPEG.format(python=sys.executable))
    exe.chmod(exe.stat().st_mode | stat.S_IXUSR)
    os.environ["IQ_TO_AUDIO_FFMPEG"] = str(exe)
    os.environ["IQ2A_CAPTURE_DIR"] = str(tmp)


def main() -> None:
    if not REF_SRC.exists():
        raise SystemExit("reference tree not present; goldens can only be regenerated in the build container")
    logging.basicConfig(level=logging.INFO, format="%(message)s")
    with tempfile.TemporaryDirectory() as t:
        tmp = Path(t)
        _install_shims(tmp)
        sys.path.insert(0, str(REF_SRC))
        import iq_to_audio.benchmark as bench                     # the unmodified reference
        import iq_to_audio.processing as proc

        # the capture is deleted with the reference's scratch directory: hash its payload as the reader sees it
        seen = {}
        real_reader = proc.IQReader.__enter__

        def spy(self):
            blob = Path(self.path).read_bytes()
            pos = blob.index(b"data") + 8
            seen["sha256"] = hashlib.sha256(blob[pos:]).hexdigest()
            seen["frames"] = (len(blob) - pos) // 4
            return real_reader(self)

        proc.IQReader.__enter__ = spy                             # observation only: the call goes through unchanged
        seconds, fs, off = 5.0, 2.5e6, 25_000.0                   # repo --benchmark defaults (cli.py, BASELINE configs[0])
        rc = bench.run_benchmark(seconds=seconds, sample_rate=fs, freq_offset=off, center_freq=None, target_freq=None,
                                 base_kwargs={"input_format": "pcm_s16le", "input_container": "wav"})
        assert rc == 0
        stream = np.frombuffer((tmp / "encoder_stdin.f32").read_bytes(), dtype="<f4").copy()
        rate = int((tmp / "encoder_args.txt").read_text().splitlines()[0])
    out = HERE / "pipeline_cfg1.npz"
    np.savez_compressed(out, clipped=stream, ffmpeg_rate=np.int64(rate), seconds=np.float64(seconds),
                        sample_rate=np.float64(fs), freq_offset=np.float64(off),
                        capture_sha256=np.array(seen["sha256"]), capture_frames=np.int64(seen["frames"]))
    print(f"wrote {out.name}: {stream.size} float32 samples at -ar {rate}, peak {np.abs(stream).max():.4f}, "
          f"capture {seen['frames']} frames sha256 {seen['sha256'][:16]}...")


if __name__ == "__main__":
    main()
