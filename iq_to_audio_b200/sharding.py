"""Time-segment sharding of a capture across the GPUs of one box.

The path shards by time: rank r owns a contiguous run of reference chunks.  A shard needs
  * a halo of `bank.halo` input samples before its start (the channel filter's history), and
  * a recurrence warm-up of W channel-rate samples started from zero state (de-emphasis and
    DC blocker are contractions; W = ceil(log(1e-9) / log(pole)) from each target's actual pole), which costs another
    (W+1)*D input samples of read-ahead;
everything else is closed form in the global sample index (NCO phase table, decimation phase,
AGC restart points, statistics windows), so there is NO data-path collective.  Only the audio
is gathered (NCCL on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

from dataclasses import dataclass

import math

#: the recurrences a shard restarts from zero state must have forgotten it to this level when its own rows begin
WARMUP_EPS = 1e-9
DC_BLOCKER_POLE = 0.995          # ref: decoders/common.py:9
_NO_STATE = ("iq", "none", "pass")


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


def _mode_of(target) -> str:
    return (target if isinstance(target, str) else getattr(target, "mode", "nfm")).lower()


def warmup_rows_for(targets, *, fs_channel: float = 96_153.846, deemph_us: float = 300.0, eps: float = WARMUP_EPS) -> int:
    """Channel-rate rows a shard recomputes from zero state before its first own row.

    A first-order recurrence with pole a has forgotten its start after ceil(log(eps) / log(a)) rows, so the count
    follows the ACTUAL pole of every target: the de-emphasis pole exp(-1 / (fs_channel * tau))
    (ref: decoders/nfm.py:40-47; 598 rows for 300 us at 96 154 Hz, 1 495 for 750 us, 1 195 at fs_channel 192 kHz)
    and the DC blocker's 0.995 for AM / SSB (4 135 rows).  `targets` are mode names or objects with `.mode` and
    optionally `.deemph_us` (bank.Target); the largest requirement wins."""
    need = 0
    for t in targets:
        mode = _mode_of(t)
        if mode in _NO_STATE:
            continue
        if mode in ("nfm", "fm"):
            tau = max(float(getattr(t, "deemph_us", deemph_us) if not isinstance(t, str) else deemph_us) * 1e-6, 1e-6)
            pole = math.exp(-1.0 / (float(fs_channel) * tau))
        elif mode in ("am", "usb", "ssb", "lsb"):
            pole = DC_BLOCKER_POLE
        else:
            raise KeyError(mode)
        need = max(need, int(math.ceil(math.log(eps) / math.log(pole))))
    return need


def plan_segments(n_total: int, world: int, chunk: int, decimation: int, halo: int, modes, *,
                  sample_rate: float | None = None, deemph_us: float = 300.0) -> list[Segment]:
    """Split [0, n_total) into `world` contiguous segments whose starts are multiples of the
    reference chunk (AGC restarts and NCO phase wraps then fall where the single-stream run
    has them).  Trailing ranks may get empty segments when there are fewer chunks than ranks.

    `modes`: mode names or Target objects; `sample_rate` (input rate, so fs_channel = sample_rate / decimation)
    and `deemph_us` size the recurrence warm-up from the targets' real poles (warmup_rows_for); without a
    sample rate the reference's default channel rate (96 153.846 Hz) is assumed."""
    if world < 1 or chunk < 1 or decimation < 1:
        raise ValueError("world, chunk and decimation must be positive")
    n_chunks = (n_total + chunk - 1) // chunk
    fs_ch = float(sample_rate) / decimation if sample_rate else 96_153.846
    warm = warmup_rows_for(modes, fs_channel=fs_ch, deemph_us=deemph_us)
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


class WriterExchange:
    """Per-target writers: the audio of target c from every time shard is assembled on rank c % world.

    One output file per target means one writer per target; spreading the writers over the ranks turns the
    many-to-one gather (rank 0 would take in (world-1)/world of ALL audio, every step) into a balanced exchange in
    which a writer receives only its own targets.  Channels are dealt round-robin, so the rows a rank sends in
    round j -- targets j*world .. j*world+world-1, one per destination in rank order -- are contiguous in the
    [C, rows] audio layout and go out as one `all_to_all_single` with split sizes (NCCL on GPUs, gloo in the CPU
    tests); no packing kernel, no copy.
    """

    def __init__(self, n_channels: int, rows: int, dtype, device, *, depth: int = 2):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.C, self.rows = int(n_channels), int(rows)
        self.rounds = (self.C + self.world - 1) // self.world
        self.owned = [c for c in range(self.C) if c % self.world == self.rank]
        self.empty = torch.empty(0, dtype=dtype, device=device)
        # recv[slot][j]: [world, rows] -- target j*world + rank as produced by shard 0 .. world-1
        self.recv = [[torch.empty((self.world, self.rows), dtype=dtype, device=device) for _ in self.owned]
                     for _ in range(depth)]

    def exchange(self, slot: int, audio, *, async_op: bool = True):
        """audio: [C, rows] of this rank's shard.  Returns the list of work handles (one per round)."""
        works = []
        w = self.world
        for j in range(self.rounds):
            lo, hi = j * w, min(self.C, (j + 1) * w)
            send = audio[lo:hi].reshape(-1)
            in_split = [self.rows if lo + r < self.C else 0 for r in range(w)]
            mine = lo + self.rank < self.C
            out = self.recv[slot][j].reshape(-1) if mine else self.empty
            out_split = [self.rows if mine else 0] * w
            works.append(self.dist.all_to_all_single(out, send, out_split, in_split, async_op=async_op))
        return works

    def result(self, slot: int) -> dict:
        """{target: [world, rows]} for the targets this rank writes (rows of shard s in row s)."""
        return {c: self.recv[slot][j] for j, c in enumerate(self.owned)}


class PeerGather:
    """Audio gather over NVLink peer memory, for pipelined steps on the GPUs of one box.

    Rank `dst` owns a receive area that every rank can address (symmetric memory).  After the kernels of step k,
    each other rank pushes its rows into its place in that area with one device-to-device copy on a side stream:
    world-1 copy engines write over NVSwitch in parallel (measured 740 GB/s per engine, unaffected by SM load,
    tools/peer_copy_probe.py), no SMs are used, so the transfer of step k runs under the persistent channel-bank
    kernel of step k+1.  (An NCCL gather needs SMs that the kernel occupies and ends up serialised behind it.)

        compute stream :  compute(k) -> wait push(k-1) -> [dst: wait consume(k-2)] -> barrier(k) -> compute(k+1)
        side stream    :  wait compute(k) -> push(k)            dst: wait barrier(k) -> consume(k-1)

    The signal-pad barrier sits on the compute stream, between two steps, where SMs are free (next to the
    persistent kernel it could not be scheduled).  Once barrier(k) has passed on `dst`, every push(k-1) has landed.
    Three slots make reuse strict: push(k+3) overwrites slot k%3 after the sender's barrier(k+2), which `dst` enters
    only after consume(k).  `torch.distributed._symmetric_memory` provides the allocation, the peer views and the
    barrier; NCCL is not in the data path.
    """

    DEPTH = 3

    def __init__(self, shape, dtype, device, *, dst: int = 0):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.torch, self.dist = torch, dist
        self.rank, self.world, self.dst = dist.get_rank(), dist.get_world_size(), dst
        self.shape, self.dtype = tuple(shape), dtype
        numel = 1
        for v in self.shape:
            numel *= int(v)
        depth = self.DEPTH
        self.local = [torch.empty(self.shape, dtype=dtype, device=device) for _ in range(depth)]
        self.recv = symm.empty((depth, self.world, *self.shape), dtype=dtype, device=device)
        self.hdl = symm.rendezvous(self.recv, dist.group.WORLD)
        self.side = torch.cuda.Stream(device=device)
        self.pushed = {}                        # step -> event: push has left the local slot and landed on dst
        self.consumed = {}                      # dst: step -> event: consumer work finished
        self.consumer = None                    # dst: callable(step, tensors) run on the side stream
        self.mine_on_dst = [self.hdl.get_buffer(dst, self.shape, dtype, (s * self.world + self.rank) * numel)
                            for s in range(depth)]

    def slot(self, k: int):
        return self.local[k % self.DEPTH]

    def before_compute(self, k: int, stream) -> None:
        """Order the kernels that will overwrite slot k behind the push that last read it."""
        ev = self.pushed.pop(k - self.DEPTH, None)
        if ev is not None:
            stream.wait_event(ev)

    def publish(self, k: int, stream) -> None:
        """Call once the kernels producing slot k have been enqueued on `stream` (the compute stream)."""
        torch = self.torch
        ready = torch.cuda.Event()
        ready.record(stream)
        self.side.wait_event(ready)
        with torch.cuda.stream(self.side):
            if self.rank != self.dst:
                self.mine_on_dst[k % self.DEPTH].copy_(self.local[k % self.DEPTH], non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.side)
        self.pushed[k] = done
        prev = self.pushed.get(k - 1)
        if prev is not None:
            stream.wait_event(prev)
        if self.rank == self.dst:
            old = self.consumed.pop(k - 2, None)
            if old is not None:
                stream.wait_event(old)
        with torch.cuda.stream(stream):
            self.hdl.barrier(channel=0)
        if self.rank == self.dst and k >= 1:
            landed = torch.cuda.Event()
            landed.record(stream)
            self._consume(k - 1, landed)

    def _consume(self, k: int, after) -> None:
        torch = self.torch
        self.side.wait_event(after)
        with torch.cuda.stream(self.side):
            if self.consumer is not None:
                self.consumer(k, self.result(k))
            ev = torch.cuda.Event()
            ev.record(self.side)
        self.consumed[k] = ev

    def result(self, k: int):
        """dst only: per-rank tensors of step k (valid for the consumer, or after `flush`)."""
        s = k % self.DEPTH
        return [self.local[s] if r == self.dst else self.recv[s, r] for r in range(self.world)]

    def flush(self, last_step: int, stream) -> None:
        """After the last publish: wait until every push has landed and the consumer has seen the last step."""
        stream.synchronize()
        self.side.synchronize()
        self.dist.barrier()
        if self.rank == self.dst and last_step >= 0 and last_step not in self.consumed:
            self._consume(last_step, self.torch.cuda.Event())
            self.side.synchronize()


def writer_place(c: int, shard: int, world: int, per_rank: int, slot: int, rows: int) -> tuple[int, int, int]:
    """Where shard `shard`'s rows of target c go: (writer rank, index among the writer's targets, element offset into
    the writer's receive area laid out [slot][per_rank][world][rows])."""
    owner, j = c % world, c // world
    return owner, j, ((slot * per_rank + j) * world + shard) * rows


class PeerWriters:
    """Per-target writers fed over NVLink peer memory by copy engines (no NCCL, no SMs in the data path).

    The ownership rule of `WriterExchange` (target c is assembled on rank c % world) with the transport of
    `PeerGather`: every rank owns a receive area that all ranks can address (symmetric memory) and, after the
    kernels of step k, pushes each target's rows straight into their place on that target's writer with one
    device-to-device copy per target on a side stream.  A writer takes in only its own targets -- (world-1)/world
    of ONE target per shard instead of everything -- so the many-to-one congestion that sank the push to rank 0 at
    N = 8 does not arise, and because copy engines do the moving the persistent channel-bank kernel keeps every SM
    (the NCCL exchange needs 16 of them reserved).  Nothing on the COMPUTE stream waits for another rank:

        compute stream :  wait push(k-3) done (own slot free) -> compute(k)
        side stream    :  wait compute(k) -> push(k) -> barrier(k) -> [owner: consume(k)]

    barrier(k) (signal pads of the symmetric allocation) runs on the side stream: once it has passed, every push(k)
    has landed everywhere and the owner may read slot k % 3; a pusher reaches push(k+3) only after barrier(k+2), by
    which time every owner has consumed step k, so slot reuse is safe without any cross-rank wait in the compute
    path.  (Round 1 had the barrier on the compute stream between two steps: every step then took the maximum over
    the ranks plus the barrier's latency, 2.41 -> 2.64 ms at N = 8.)  The ranks may drift up to three steps apart.
    """

    DEPTH = 3

    def __init__(self, n_channels: int, rows: int, dtype, device):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.torch, self.dist = torch, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.C, self.rows, self.dtype = int(n_channels), int(rows), dtype
        self.per_rank = (self.C + self.world - 1) // self.world          # targets a writer can own
        self.owned = [c for c in range(self.C) if c % self.world == self.rank]
        depth = self.DEPTH
        self.local = [torch.empty((self.C, self.rows), dtype=dtype, device=device) for _ in range(depth)]
        # recv[slot][j][s]: target j * world + rank as produced by shard s (same shape on every rank: symmetric)
        self.recv = symm.empty((depth, self.per_rank, self.world, self.rows), dtype=dtype, device=device)
        self.hdl = symm.rendezvous(self.recv, dist.group.WORLD)
        self.side = torch.cuda.Stream(device=device)
        self.pushed = {}
        self.consumed = {}
        self.consumer = None                    # owners: callable(step, {target: [world, rows]}) on the side stream
        # where this rank's rows of target c live on the writer of c, per slot
        self.dest = [[self._place(slot, c) for c in range(self.C)] for slot in range(depth)]

    def _place(self, slot: int, c: int):
        owner, j, off = writer_place(c, self.rank, self.world, self.per_rank, slot, self.rows)
        if owner == self.rank:
            return self.recv[slot, j, self.rank]
        return self.hdl.get_buffer(owner, (self.rows,), self.dtype, off)

    def slot(self, k: int):
        return self.local[k % self.DEPTH]

    def before_compute(self, k: int, stream) -> None:
        ev = self.pushed.pop(k - self.DEPTH, None)
        if ev is not None:
            stream.wait_event(ev)

    def publish(self, k: int, stream) -> None:
        torch = self.torch
        s = k % self.DEPTH
        ready = torch.cuda.Event()
        ready.record(stream)
        self.side.wait_event(ready)
        with torch.cuda.stream(self.side):
            # start with the neighbour's targets so that the writers are not all hit in the same order
            for i in range(self.C):
                c = (self.rank + 1 + i) % self.C
                self.dest[s][c].copy_(self.local[s][c], non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.side)
            self.hdl.barrier(channel=0)                   # every push of step k has landed on every writer
            if self.owned:
                if self.consumer is not None:
                    self.consumer(k, self.result(k))
                ev = torch.cuda.Event()
                ev.record(self.side)
                self.consumed[k] = ev
                self.consumed.pop(k - self.DEPTH, None)
        self.pushed[k] = done

    def result(self, k: int) -> dict:
        """{target: [world, rows]} for the targets this rank writes (rows of shard s in row s)."""
        s = k % self.DEPTH
        return {c: self.recv[s, j] for j, c in enumerate(self.owned)}

    def flush(self, last_step: int, stream) -> None:
        """After the last publish: returns when every rank's pushes of every step have landed on their writers.  The
        last thing on the side stream is barrier(last_step) across the ranks (signal pads), so waiting for the side
        stream IS the rendezvous; a host-side collective barrier on top of it only added its own latency."""
        stream.synchronize()
        self.side.synchronize()
