"""Time-segment sharding of a capture across the GPUs of one box.

The path shards by time: rank r owns a contiguous run of reference chunks.  A shard needs
  * a halo of `bank.halo` input samples before its start (the channel filter's history), and
  * a recurrence warm-up of W channel-rate samples started from zero state (de-emphasis and
    DC blocker are contractions: 0.966^600 and 0.995^4200 are below 1e-9), which costs another
    (W+1)*D input samples of read-ahead;
everything else is closed form in the global sample index (NCO phase table, decimation phase,
AGC restart points, statistics windows), so there is NO data-path collective.  Only the audio
is gathered (NCCL on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

from dataclasses import dataclass

WARMUP_ROWS = {"nfm": 600, "fm": 600, "am": 4200, "usb": 4200, "ssb": 4200, "lsb": 4200,
               "iq": 0, "none": 0, "pass": 0}


@dataclass(frozen=True)
class Segment:
    rank: int
    begin: int          # first input sample owned (multiple of the reference chunk)
    end: int            # one past the last input sample owned
    first_frame: int    # first input sample that must be resident (halo + warm-up, 4-frame aligned)
    warmup_rows: int    # channel-rate rows recomputed from zero state and discarded
    row_begin: int      # first channel-rate row produced (ceil(begin / D))
    row_end: int

    @property
    def rows(self) -> int:
        return self.row_end - self.row_begin

    @property
    def frames(self) -> int:
        return self.end - self.first_frame


def warmup_rows_for(modes) -> int:
    return max((WARMUP_ROWS[m.lower()] for m in modes), default=0)


def plan_segments(n_total: int, world: int, chunk: int, decimation: int, halo: int, modes) -> list[Segment]:
    """Split [0, n_total) into `world` contiguous segments whose starts are multiples of the
    reference chunk (AGC restarts and NCO phase wraps then fall where the single-stream run
    has them).  Trailing ranks may get empty segments when there are fewer chunks than ranks."""
    if world < 1 or chunk < 1 or decimation < 1:
        raise ValueError("world, chunk and decimation must be positive")
    n_chunks = (n_total + chunk - 1) // chunk
    warm = warmup_rows_for(modes)
    d = decimation
    segs = []
    for r in range(world):
        c0 = (n_chunks * r) // world
        c1 = (n_chunks * (r + 1)) // world
        begin = min(n_total, c0 * chunk)
        end = min(n_total, c1 * chunk)
        w = warm if begin > 0 else 0
        row_begin = (begin + d - 1) // d
        row_end = (end + d - 1) // d
        first = max(0, (row_begin - w) * d - halo) if begin > 0 else 0
        first -= first % 4
        segs.append(Segment(r, begin, end, first, min(w, row_begin), row_begin, row_end))
    return segs


def gather_rows(local, segs: list[Segment], dst: int = 0):
    """Gather per-rank audio [C, rows_r] (torch tensors, same dtype/device kind on every rank) to
    `dst` in segment order; returns the concatenated [C, total_rows] tensor on `dst`, None elsewhere.
    Ranks own different row counts, so each rank pads to the maximum before the collective."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    rank = dist.get_rank()
    width = max(s.rows for s in segs)
    padded = torch.zeros((local.shape[0], width), dtype=local.dtype, device=local.device)
    padded[:, : local.shape[1]] = local
    bucket = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, bucket, dst=dst)
    if rank != dst:
        return None
    return torch.cat([bucket[s.rank][:, : s.rows] for s in segs], dim=1)


def reduce_peak(local_peaks, dst: int = 0):
    """AudioWriter.peak of the whole run = max over shards (ref: processing.py:449-451)."""
    import torch.distributed as dist
    dist.reduce(local_peaks, dst=dst, op=dist.ReduceOp.MAX)
    return local_peaks
