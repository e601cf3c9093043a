"""CPU-only: the capture reader's block logic (no CUDA calls: the pinned ring is replaced by ordinary arrays)."""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

from iq_to_audio_b200.pipeline import InputFormat, IQReader


def _read_all(reader: IQReader, **kw) -> list[np.ndarray]:
    out = []
    while (blk := reader.read_raw_block(**kw)) is not None:
        out.append(blk.copy())
    return out


def test_parallel_positional_reads_equal_the_sequential_reader(tmp_path: Path):
    """Large blocks are filled by several preadv slices: same bytes, same block sizes, same handling of the trailing
    partial frame (reference: processing.py:253-256) as the plain readinto loop."""
    rng = np.random.default_rng(1)
    header = 44
    frames_per_block = (9 << 20) // 4
    data = rng.integers(0, 256, size=header + 3 * (9 << 20) + 12_345 * 4 + 2, dtype=np.uint8)   # ragged tail
    path = tmp_path / "capture.cs16"
    path.write_bytes(data.tobytes())
    fmt = InputFormat("raw", "pcm_s16le", 1e6, header)
    with IQReader(path, frames_per_block, "iq", fmt, sample_rate=1e6) as plain:
        want = _read_all(plain)
    assert [b.size for b in want] == [9 << 20] * 3 + [12_345 * 4]

    par = IQReader(path, frames_per_block, "iq", fmt, sample_rate=1e6)
    par._fh = path.open("rb", buffering=0)
    par._fh.seek(header)
    par._pos = header
    par._ring = [np.empty(9 << 20, dtype=np.uint8) for _ in range(IQReader.RING_DEPTH)]
    par._pool = ThreadPoolExecutor(IQReader.READ_THREADS)
    try:
        got = _read_all(par)
        # a block limited by max_frames (the --max-input-seconds path) keeps the position in step as well
        par._fh.seek(header)
        par._pos = header
        first = par.read_raw_block(max_frames=frames_per_block - 7).copy()
        second = par.read_raw_block().copy()
    finally:
        par._pool.shutdown()
        par._fh.close()
    assert len(got) == len(want) and all((a == b).all() for a, b in zip(got, want))
    payload = data[header:]
    assert (first == payload[:first.size]).all() and first.size == (frames_per_block - 7) * 4
    assert (second == payload[first.size:first.size + second.size]).all()
