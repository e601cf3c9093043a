"""CPU-only, world_size 2 over gloo: the time-segment sharding logic (segment plan, halo,
warm-up, gather order) with the CPU oracle standing in for the per-rank engine."""
from __future__ import annotations

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from iq_to_audio_b200 import sharding
from oracle import iq_oracle as orc
from tests import _cases


def test_segment_plan_covers_capture_on_the_chunk_grid():
    segs = sharding.plan_segments(10_000_000, 4, 262_144, 104, 6_552, ["nfm"] * 5)
    assert segs[0].begin == 0 and segs[-1].end == 10_000_000
    for a, b in zip(segs, segs[1:]):
        assert a.end == b.begin and b.begin % 262_144 == 0 and a.row_end == b.row_begin
    assert segs[0].warmup_rows == 0 and segs[0].first_frame == 0
    for s in segs[1:]:
        assert s.warmup_rows == 598 and s.first_frame % 4 == 0          # ceil(log(1e-9) / log(pole)), 300 us at 96 154 Hz
        assert s.first_frame <= (s.row_begin - 598) * 104 - 6_552
    assert sum(s.rows for s in segs) == orc.decimated_count(0, 10_000_000, 104)
    # fewer chunks than ranks -> empty trailing segments, still a partition
    segs = sharding.plan_segments(300_000, 8, 262_144, 104, 6_552, ["am"])
    assert sum(s.rows for s in segs) == orc.decimated_count(0, 300_000, 104)
    assert sharding.warmup_rows_for(["nfm", "usb"]) == 4135             # the DC blocker's 0.995 pole
    # the warm-up follows the real de-emphasis pole: longer time constants and higher channel rates need more rows
    assert sharding.warmup_rows_for(["nfm"], deemph_us=750.0) == 1495
    assert sharding.warmup_rows_for(["nfm"], fs_channel=192_000.0) == 1194
    assert sharding.warmup_rows_for(["iq"]) == 0

    class _T:
        mode, deemph_us = "nfm", 750.0
    assert sharding.warmup_rows_for([_T(), "nfm"]) == 1495
    assert sharding.plan_segments(10_000_000, 2, 262_144, 104, 6_552, ["nfm"], sample_rate=10e6,
                                  deemph_us=750.0)[1].warmup_rows == 1495
    with pytest.raises(ValueError):
        sharding.plan_segments(10, 0, 1, 1, 0, ["nfm"])


def _oracle_shard(x, seg, plan, chunk):
    """What a rank computes: the reference loop started `warmup_rows` early from zero state,
    with the NCO phase and decimator offset taken from the global index."""
    d = plan.decimation
    nt = len(plan.taps)
    if seg.begin == 0:
        return orc.run_target(x[: seg.end], plan, chunk).audio
    h0 = (seg.row_begin - seg.warmup_rows) * d - (nt - 1)
    nco = orc.NcoState.for_offset(plan.freq_offset, plan.sample_rate)
    # exact phase at h0: replay the reference's per-chunk wrap up to the chunk holding h0
    k0 = h0 // chunk
    for _ in range(k0):
        nco.phase = orc.nco_advance(nco.phase, nco.increment, plan.mix_sign, chunk)
    nco.phase = nco.phase + plan.mix_sign * nco.increment * (h0 - k0 * chunk)
    fir = orc.FirState(plan.taps, plan.filter_block)
    dec = orc.DecimState(d)
    dec.offset = h0 % d
    dem = orc.DemodState.create(plan.mode, plan.fs_channel, deemph_us=plan.deemph_us)
    # one call per reference chunk piece so that phase wraps happen where the single stream has them
    audio = []
    pos = h0
    while pos < seg.end:
        nxt = min(seg.end, (pos // chunk + 1) * chunk)
        mixed = orc.nco_mix(nco, x[pos:nxt], plan.mix_sign)
        audio.append(orc.demodulate(dem, orc.decimate(dec, orc.fir_overlap_save(fir, mixed)))[0])
        pos = nxt
    audio = np.concatenate(audio)
    return audio[audio.size - seg.rows:]


def _worker(rank, world, port, ret, deemph_us=300.0):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m = _cases.manifest()["case_b_nfm_10M"]
        x = _cases.complex_input("case_b_nfm_10M")
        chunk = m["chunk"]
        plan = orc.TargetPlan(sample_rate=m["fs"], freq_offset=m["targets"][1]["f_off"], mix_sign=1, deemph_us=deemph_us)
        halo = ((len(plan.taps) - 1 + plan.decimation - 1) // plan.decimation + 1) * plan.decimation
        segs = sharding.plan_segments(x.size, world, chunk, plan.decimation, halo, ["nfm"], sample_rate=m["fs"],
                                      deemph_us=deemph_us)
        mine = _oracle_shard(x, segs[rank], plan, chunk)
        local = torch.from_numpy(np.ascontiguousarray(mine)).reshape(1, -1)
        full = sharding.gather_rows(local, segs, dst=0)
        peak = sharding.reduce_peak(torch.tensor([float(np.abs(mine).max())]), dst=0)
        if rank == 0:
            ret["audio"] = full.numpy()[0].copy()
            ret["peak"] = float(peak[0])
    finally:
        dist.destroy_process_group()


def test_two_rank_time_sharding_reproduces_single_stream():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    g = _cases.load("case_b_nfm_10M_t1")
    audio = ret["audio"]
    assert audio.size == g["audio"].size                      # exact row partition, in order
    assert np.abs(audio - g["audio"]).max() <= 1e-6           # halo + 600-row warm-up: < 1e-8 expected
    assert abs(ret["peak"] - float(g["peak"])) <= 1e-6


def test_time_sharding_with_a_long_deemphasis_time_constant():
    """750 us de-emphasis (pole 0.986): the shard warm-up has to follow the pole (1 495 rows, not the 598 of the
    300 us default), or every shard after the first starts ~1e-4 off the single-stream run."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret, 750.0), nprocs=2, join=True)
    m = _cases.manifest()["case_b_nfm_10M"]
    x = _cases.complex_input("case_b_nfm_10M")
    plan = orc.TargetPlan(sample_rate=m["fs"], freq_offset=m["targets"][1]["f_off"], mix_sign=1, deemph_us=750.0)
    want = orc.run_target(x, plan, m["chunk"]).audio
    assert ret["audio"].size == want.size
    assert np.abs(ret["audio"] - want).max() <= 1e-7
    # the old fixed table (600 rows) misses by orders of magnitude more: 0.98675^600 = 3e-4 of the state survives
    assert 0.98675 ** 600 > 1e-4 and 0.98675 ** 1495 < 3e-9


def _xchg_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        c_total, rows = 5, 37
        x = sharding.WriterExchange(c_total, rows, torch.float32, torch.device("cpu"))
        # shard `rank` of target c carries the value 100 c + rank (+ the row index / 1000)
        audio = torch.stack([100.0 * c + rank + torch.arange(rows) / 1000.0 for c in range(c_total)]).float()
        for w in x.exchange(0, audio):
            w.wait()
        for w in x.exchange(1, audio + 0.5, async_op=False) or []:
            pass
        ret[rank] = ({c: t.clone().numpy() for c, t in x.result(0).items()},
                     {c: t.clone().numpy() for c, t in x.result(1).items()}, x.owned, x.rounds)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_per_target_writers_exchange(world):
    """Target c ends up on rank c % world with the shards of all ranks in rank order (5 targets on 2 and 3 ranks:
    uneven ownership, a last round in which only some ranks receive)."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_xchg_worker, args=(world, port, ret), nprocs=world, join=True)
    seen = set()
    for rank in range(world):
        r0, r1, owned, rounds = ret[rank]
        assert owned == [c for c in range(5) if c % world == rank] and rounds == -(-5 // world)
        assert sorted(r0) == owned
        for c in owned:
            seen.add(c)
            want = np.stack([100.0 * c + src + np.arange(37) / 1000.0 for src in range(world)]).astype(np.float32)
            np.testing.assert_array_equal(r0[c], want)
            np.testing.assert_array_equal(r1[c], (want.astype(np.float64) + 0.5).astype(np.float32))
    assert seen == set(range(5))


def test_writer_place_covers_every_receive_area_exactly_once():
    """PeerWriters' addressing: for every (slot, target, shard) one place, on the target's writer, no two alike, and
    all inside the symmetric allocation [depth][per_rank][world][rows] -- for target counts below, equal to and above
    the number of ranks (writers that own nothing, one, several targets)."""
    from iq_to_audio_b200.sharding import writer_place
    rows, depth = 7, 3
    for world in (2, 3, 4, 8):
        for C in (1, 3, 5, 8, 11):
            per_rank = (C + world - 1) // world
            seen = {r: set() for r in range(world)}
            for slot in range(depth):
                for c in range(C):
                    for shard in range(world):
                        owner, j, off = writer_place(c, shard, world, per_rank, slot, rows)
                        assert owner == c % world and j == c // world and j < per_rank
                        assert 0 <= off and off + rows <= depth * per_rank * world * rows
                        assert off % rows == 0 and off not in seen[owner]
                        seen[owner].add(off)
            for r in range(world):
                owned = len([c for c in range(C) if c % world == r])
                assert len(seen[r]) == depth * owned * world


def test_numa_helpers_are_safe_without_a_gpu():
    """numa.bind_to_device_numa is a no-op (and says so) when the topology is unknown -- e.g. in this CPU container."""
    from iq_to_audio_b200 import numa
    assert numa._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert numa._parse_cpulist("") == set()
    before = os.sched_getaffinity(0)
    info = numa.bind_to_device_numa(0)
    assert info["device"] == 0 and (info["bound"] or os.sched_getaffinity(0) == before)
