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


class PeerGather:
    """Audio gather over NVLink peer memory, for pipelined steps on the GPUs of one box.

    Every rank writes its rows straight into a slot of a symmetric (peer-mapped) buffer; rank `dst` pulls the
    slots of the other ranks with device-to-device copies on a side stream -- copy engines over NVSwitch, no SMs,
    so the pull of step k runs under the persistent channel-bank kernel of step k+1 (an NCCL gather needs SMs that
    the kernel occupies and ends up serialised behind it).  A signal-pad barrier per step on the compute stream
    orders producers and the consumer:

        compute(k) -> [dst: wait until pull(k-1) finished] -> barrier(k) -> compute(k+1) ...
                                                                 \\-> dst side stream: pull(k)

    With two slots, slot k%2 is rewritten by compute(k+2), which every rank enqueues behind barrier(k+1), and dst
    enters barrier(k+1) only after pull(k) has finished.  `torch.distributed._symmetric_memory` provides the
    allocation, the peer views and the barrier; NCCL is not involved in the data path.
    """

    def __init__(self, shape, dtype, device, *, dst: int = 0, depth: int = 2):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.torch = torch
        self.rank, self.world, self.dst, self.depth = dist.get_rank(), dist.get_world_size(), dst, depth
        self.shape, self.dtype = tuple(shape), dtype
        self.numel = 1
        for v in self.shape:
            self.numel *= int(v)
        self.buf = symm.empty((depth, *self.shape), dtype=dtype, device=device)
        self.hdl = symm.rendezvous(self.buf, dist.group.WORLD)
        self.pull_stream = torch.cuda.Stream(device=device) if self.rank == dst else None
        self.pulled = [None] * depth            # dst: event "pull of this slot finished"
        self.out = None
        if self.rank == dst:
            self.out = [[torch.empty(self.shape, dtype=dtype, device=device) for _ in range(self.world)]
                        for _ in range(depth)]
            self.peers = [[self.hdl.get_buffer(r, self.shape, dtype, s * self.numel) for r in range(self.world)]
                          for s in range(depth)]

    def slot(self, k: int):
        return self.buf[k % self.depth]

    def publish(self, k: int, stream) -> None:
        """Call once the kernels producing slot k have been enqueued on `stream` (the compute stream)."""
        torch = self.torch
        s = k % self.depth
        with torch.cuda.stream(stream):
            if self.rank == self.dst:
                prev = self.pulled[(k + 1) % self.depth]
                if prev is not None:
                    stream.wait_event(prev)
            self.hdl.barrier(channel=0)
            if self.rank == self.dst:
                ready = torch.cuda.Event()
                ready.record(stream)
        if self.rank == self.dst:
            self.pull_stream.wait_event(ready)
            with torch.cuda.stream(self.pull_stream):
                for r in range(self.world):
                    if r != self.dst:
                        self.out[s][r].copy_(self.peers[s][r], non_blocking=True)
                done = torch.cuda.Event()
                done.record(self.pull_stream)
            self.pulled[s] = done

    def result(self, k: int):
        """dst only: list of per-rank tensors of step k (own slot included), valid after `drain()`."""
        s = k % self.depth
        self.out[s][self.dst] = self.buf[s]
        return self.out[s]

    def drain(self) -> None:
        if self.pull_stream is not None:
            self.pull_stream.synchronize()
