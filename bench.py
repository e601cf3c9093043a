#!/usr/bin/env python
"""bench.py -- throughput of the channelize-and-demodulate hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--seconds S]

Workload (BASELINE.json configs[1]): SDR++-style int16 baseband capture, 10 MS/s, 60 s,
5 NFM targets batched, de-emphasis 300 us, reference chunk 4 194 304 (tune_chunk_size).
One "step" = one pass of the whole capture through the hot path:
  value : capture already resident in HBM, `ChannelBank.process_resident` (device timed)
  e2e   : capture in pinned HOST memory, streamed chunk by chunk through
          `ChannelBank.process_chunk` (the reference loop body), H2D of every chunk and
          D2H of every audio block inside the timed region
N > 1 (torchrun): the capture is N x 60 s, time-sharded -- rank r owns seconds [60 r, 60 (r+1))
with a filter-length halo and a recurrence warm-up; only audio is gathered to rank 0 (NCCL).

`--impl reference` times the CPU restatement of the reference path (oracle/iq_oracle.py; the
reference is pure Python and cannot travel to the GPU box) on a bounded sample, one process
per target in parallel.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

FS = 10e6
SECONDS = 60.0
OFFSETS = [-3.2e6, -1.1e6, 0.4e6, 2.3e6, 4.1e6]
# test knob (not the benchmark): fewer targets than ranks exercises writers that own nothing
if os.environ.get("IQ2A_BENCH_TARGETS"):
    OFFSETS = OFFSETS[:max(1, int(os.environ["IQ2A_BENCH_TARGETS"]))]
BW = 12_500.0
DEEMPH_US = 300.0
SM_RESERVE = 16           # multi-GPU runs: SMs kept free of the persistent kernel for the NCCL exchange (of 148)
REQ_CHUNK = 1_048_576
METRIC = "input complex Msamples/s"
UNIT = "Msamples/s"


def workload_name(seconds: float) -> str:
    return f"cfg2: int16 IQ, 10 MS/s, {seconds:g} s, 5 NFM targets batched, deemph 300 us"


# ------------------------------------------------------------------------------------------
# synthetic capture: 5 FM carriers + AWGN, int16 interleaved
# ------------------------------------------------------------------------------------------
def synth_capture_device(n0: int, n: int, device, seed: int, fs: float | None = None, carriers=None, noise: float = 0.02):
    """int16 [2*n] on `device` for global sample indices [n0, n0+n): carriers are functions of the
    global index (float64 phase), noise is seeded per call.  `carriers`: (offset_hz, kind, tone_hz, amplitude) with
    kind fm / am / usb / lsb; default: the five FM carriers of cfg2 at the module's FS."""
    import torch
    fs = FS if fs is None else fs
    if carriers is None:
        carriers = [(off, "fm", 700.0 + 150.0 * i, 0.12) for i, off in enumerate(OFFSETS)]
    out = torch.empty(2 * n, dtype=torch.int16, device=device)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    step = 1 << 23
    two_pi = 2.0 * np.pi
    for s in range(0, n, step):
        m = min(step, n - s)
        t = (torch.arange(m, device=device, dtype=torch.float64) + float(n0 + s)) / fs
        re = torch.zeros(m, device=device, dtype=torch.float64)
        im = torch.zeros(m, device=device, dtype=torch.float64)
        for off, kind, tone, amp in carriers:
            if kind == "fm":
                ph, env = two_pi * off * t + (2500.0 / tone) * torch.sin(two_pi * tone * t), amp
            elif kind == "am":
                ph, env = two_pi * off * t, amp * (1.0 + 0.8 * torch.cos(two_pi * tone * t))
            else:                                               # an analytic tone above (usb) / below (lsb) the carrier
                ph, env = two_pi * (off + (tone if kind == "usb" else -tone)) * t, amp
            ph = torch.remainder(ph, two_pi)
            re += env * torch.cos(ph)
            im += env * torch.sin(ph)
        awgn = torch.randn((m, 2), device=device, dtype=torch.float32, generator=gen) * noise
        iq = torch.stack((re.float() + awgn[:, 0], im.float() + awgn[:, 1]), dim=1).clamp_(-0.999, 0.999)
        out[2 * s:2 * (s + m)] = torch.round(iq * 32767.0).to(torch.int16).reshape(-1)
    return out


# ------------------------------------------------------------------------------------------
# the other BASELINE configurations, measured in the same run (the headline stays cfg2)
# ------------------------------------------------------------------------------------------
def other_workload_specs():
    """(key, description, fs, seconds per GPU, [(offset, mode, bandwidth, agc)], all_ranks)"""
    nfm = lambda offs: [(o, "nfm", 12_500.0, True) for o in offs]
    grid = lambda c: [(-29.0 + 58.0 * (i + 0.5) / c) * 1e6 for i in range(c)]
    return [
        ("cfg1", "repo --benchmark shape: 2.5 MS/s, 5 s, 1 NFM target at +25 kHz", 2.5e6, 5.0, nfm([25e3]), False),
        ("cfg3", "20 MS/s, AM + USB + LSB targets, AGC on (SSB on the bit-faithful path)", 20e6, 5.0,
         [(2.3e6, "am", 10_000.0, True), (-4.1e6, "usb", 2_800.0, True), (6.2e6, "lsb", 2_800.0, True)], False),
        ("cfg4", "61.44 MS/s, 5 NFM targets, one 10 s time shard per GPU (halo + recurrence warm-up), audio to "
                 "per-target writers", 61.44e6, 10.0, nfm([-21.3e6, -9.7e6, 1.9e6, 12.4e6, 25.1e6]), True),
        # (BASELINE configs[4] also sweeps the reference's filter block 16k-256k: this path has no such parameter -- the
        # hop is the M - Vd rows of the polyphase bank's 512-point transforms -- so only the channel axis exists here)
        ("cfg5_c16", "wideband sweep point: 61.44 MS/s, 16 NFM channels on a uniform grid (filter_block n/a: hop = M - Vd rows)",
         61.44e6, 2.0, nfm(grid(16)), False),
        ("cfg5_c256", "wideband sweep point: 61.44 MS/s, 256 NFM channels on a uniform grid (filter_block n/a: hop = M - Vd rows)",
         61.44e6, 1.0, nfm(grid(256)), False),
    ]


def cpu_port_rate(fs: float, spec, n: int) -> float:
    """Msamples/s of the oracle port on one host core for ONE target of the workload (bounded sample of n samples)."""
    from oracle import iq_oracle as orc
    off, mode, bw, agc = spec
    kind = {"nfm": "fm", "am": "am", "usb": "usb", "lsb": "lsb"}[mode]
    car = dict(offset=off, amp=0.2, kind=kind, tone=900.0)
    if kind == "fm":
        car["dev"] = 2500.0
    if kind == "am":
        car["depth"] = 0.8
    x = orc.order_iq(orc.unpack_interleaved(orc.to_s16(orc.multi_carrier_capture(fs, n, [car], noise_std=0.02, seed=7)),
                                            "pcm_s16le"), "iq")
    plan = orc.TargetPlan(sample_rate=fs, freq_offset=off, bandwidth=bw, mode=mode, deemph_us=DEEMPH_US,
                          agc_enabled=agc, filter_block=65_536, mix_sign=1)
    from iq_to_audio_b200.processing import tune_chunk_size
    t0 = time.perf_counter()
    orc.run_target(x, plan, tune_chunk_size(fs, REQ_CHUNK))
    return n / (time.perf_counter() - t0) / 1e6


def run_other_workload(key, desc, fs, seconds, specs, dev, rank, world, peak, cpu: bool, reps: int = 3) -> dict:
    """One BASELINE configuration resident in HBM through ChannelBank.process_resident_async; at world > 1 every
    rank takes one time shard (sharding.plan_segments) and the audio goes to per-target writers after each pass."""
    import torch
    import torch.distributed as dist
    from iq_to_audio_b200 import sharding
    from iq_to_audio_b200.bank import ChannelBank, Target
    from iq_to_audio_b200.processing import channel_decimation, design_channel_filter, tune_chunk_size
    d, fs_ch = channel_decimation(fs, 96_000.0)
    chunk = tune_chunk_size(fs, REQ_CHUNK)
    n_seg = max(chunk, int(round(fs * seconds)) // chunk * chunk)
    taps = {}
    targets = []
    for off, mode, bw, agc in specs:
        if bw not in taps:
            taps[bw] = design_channel_filter(fs, bw, d)
        targets.append(Target(off, taps[bw], 1, mode, DEEMPH_US, agc))
    bank = ChannelBank(fs, d, targets, codec="pcm_s16le", iq_order="iq", ref_chunk=chunk, device=dev.index)
    seg = sharding.plan_segments(world * n_seg, world, chunk, d, bank.halo, targets, sample_rate=fs)[rank]
    first = seg.first_frame
    kinds = {"nfm": "fm", "am": "am", "usb": "usb", "lsb": "lsb"}
    amp = min(0.12, 0.9 / max(1, len(specs)))
    carriers = [(off, kinds[mode], 600.0 + 7.0 * i, amp) for i, (off, mode, bw, agc) in enumerate(specs)]
    capture = synth_capture_device(first, seg.end - first + d, dev, 4321 + rank, fs, carriers)
    rows = bank.rows_in(seg.begin, seg.end)
    audio = torch.empty((bank.n_channels, rows), dtype=torch.float32, device=dev)
    comp = torch.cuda.Stream(device=dev)
    # multi-GPU: audio to the per-target writers, by copy-engine pushes over peer memory under the next pass's kernels
    # (sharding.PeerWriters, as in the headline arm) when symmetric memory is available on every rank, else NCCL
    peer = xchg = None
    if world > 1:
        try:
            peer = sharding.PeerWriters(bank.n_channels, rows, torch.float32, dev)
        except Exception as exc:
            print(f"[bench] {key}: peer-memory transport unavailable ({exc!r}); using NCCL", file=sys.stderr)
        ok = torch.tensor([1 if peer is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            peer = None
            xchg = sharding.WriterExchange(bank.n_channels, rows, torch.float32, dev)
    passes = [0]

    def step():
        k = passes[0]
        passes[0] += 1
        if peer is not None:
            out_buf = peer.slot(k)
            peer.before_compute(k, comp)
            bank.process_resident_async(capture.data_ptr(), first, seg.end - first + d, seg.begin, seg.end,
                                        warmup_rows=seg.warmup_rows, dev_audio=out_buf.data_ptr(), out_stride=rows,
                                        stream=comp.cuda_stream)
            peer.publish(k, comp)
            return
        with torch.cuda.stream(comp):
            bank.process_resident_async(capture.data_ptr(), first, seg.end - first + d, seg.begin, seg.end,
                                        warmup_rows=seg.warmup_rows, dev_audio=audio.data_ptr(), out_stride=rows,
                                        stream=comp.cuda_stream)
            if xchg is not None:
                for wk in xchg.exchange(0, audio, async_op=True):
                    wk.wait()

    def finish():
        if peer is not None and passes[0] > 0:
            peer.flush(passes[0] - 1, comp)
        comp.synchronize()
        torch.cuda.synchronize()

    def sync():
        finish()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(2):
        step()
    sync()
    bank.set_timing(True)
    launches0 = bank.launches
    t0 = time.perf_counter()
    for _ in range(reps):
        step()
    finish()
    ms = (time.perf_counter() - t0) * 1e3 / reps
    timing = bank.get_timing()
    launches = (bank.launches - launches0) // reps
    bank.set_timing(False)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    n_in = seg.end - max(0, seg.begin - seg.warmup_rows * d)
    b_alg = 4.0 + len(specs) * 4.0 / d
    chan_ms = timing["channelize_ms"] / max(timing["calls"], 1)
    out = {"workload": desc, "samples_per_gpu": n_seg, "n_gpus": world, "ms": ms, "value": world * n_seg / (ms * 1e-3) / 1e6,
           "unit": UNIT, "x_realtime": world * n_seg / fs / (ms * 1e-3), "decimation": d, "taps": sorted({len(t) for t in taps.values()}),
           "fft_size": bank.fft_size, "kernel_generation": bank.kernel_generation, "channels": len(specs),
           "channelize_ms": chan_ms, "tail_ms": timing["tail_ms"] / max(timing["calls"], 1), "gpu_launches": launches,
           "roofline": {"algorithmic_bytes_per_sample": b_alg, "achieved": n_in * b_alg / (chan_ms * 1e-3) / 1e9 if chan_ms > 0 else None,
                        "step_achieved": n_seg * b_alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": n_in * b_alg / (chan_ms * 1e-3) / 1e9 / peak if chan_ms > 0 else None,
                        "step_frac": n_seg * b_alg / (ms * 1e-3) / 1e9 / peak}}
    if key == "cfg4":
        # the same shard streamed from pinned host memory through the public API (H2D of every chunk inside the timed
        # region): several reference chunks per GPU call, or a 4 Mi-sample chunk would be 7 block sets for 296 CTA slots
        host = torch.empty(2 * n_seg, dtype=torch.int16).pin_memory()
        host.copy_(capture[2 * (seg.begin - first):2 * (seg.begin - first) + 2 * n_seg])
        host_np = host.numpy()
        batch = max(1, min(8, -(-49_152 * d // chunk), (128 << 20) // (4 * chunk)))

        def stream_pass() -> int:
            bank.reset()
            span = batch * chunk
            nbytes = 0
            for r in bank.stream((host_np[2 * s:2 * min(s + span, n_seg)] for s in range(0, n_seg, span)), chunk_frames=chunk):
                nbytes += r.audio.nbytes + r.clipped.nbytes
            return nbytes
        stream_pass()
        sync()
        t0 = time.perf_counter()
        d2h = stream_pass()
        torch.cuda.synchronize()
        s_ms = (time.perf_counter() - t0) * 1e3
        mine = s_ms
        if world > 1:
            t = torch.tensor([s_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            s_ms = float(t.item())
        out["stream"] = {"value": world * n_seg / (s_ms * 1e-3) / 1e6, "unit": UNIT, "ms": s_ms, "chunks_per_call": batch,
                         "h2d_bytes": int(4 * n_seg), "d2h_bytes": int(d2h), "this_rank_h2d_GBps": 4 * n_seg / (mine * 1e-3) / 1e9,
                         "api": f"ChannelBank.stream(chunk_frames={chunk}) from pinned host memory, 2 calls in flight"}
        del host, host_np
    bank.close()
    del capture, audio
    torch.cuda.empty_cache()
    if cpu and rank == 0:
        n_cpu = 1 << 21
        per_target = float(np.mean([cpu_port_rate(fs, sp, n_cpu) for sp in specs[:3:2] or specs[:1]]))
        out["cpu_port"] = {"value": per_target / len(specs), "unit": UNIT, "cores": 1, "kind": "port",
                           "per_target": per_target,
                           "sample": f"{n_cpu} samples per target, one target of each filter length timed, targets sequential as cli.py:683 runs them"}
        out["speedup_vs_cpu_port_1core"] = out["value"] / out["cpu_port"]["value"]
    return out


def make_targets():
    from iq_to_audio_b200.bank import Target
    from iq_to_audio_b200.processing import channel_decimation, design_channel_filter
    d, fs_ch = channel_decimation(FS, 96_000.0)
    taps = design_channel_filter(FS, BW, d)
    return d, fs_ch, [Target(o, taps, 1, "nfm", DEEMPH_US, True) for o in OFFSETS]


# ------------------------------------------------------------------------------------------
# clocks / throttle reasons during the timed region
# ------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index: int, period: float = 0.001):
        self.period = period
        self.samples: list[int] = []
        self.reasons: set[str] = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in names.items():
                    if bit and (mask & bit):
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        # re-entrant: the resident and the end-to-end timed regions both add samples (one NVML query takes several
        # milliseconds on these boxes, so the 20 ms resident region alone yields one or two)
        if self.nv is not None:
            self._stop.clear()
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=2)

    def summary(self) -> dict:
        med = int(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference loop (processing.py:1070-1154)
# ------------------------------------------------------------------------------------------
def _cpu_one_target(args):
    raw, off, chunk = args
    from oracle import iq_oracle as orc
    x = orc.order_iq(orc.unpack_interleaved(raw, "pcm_s16le"), "iq")
    plan = orc.TargetPlan(sample_rate=FS, freq_offset=off, bandwidth=BW, mode="nfm", deemph_us=DEEMPH_US,
                          filter_block=65_536, mix_sign=1)
    t0 = time.perf_counter()
    res = orc.run_target(x, plan, chunk)
    return time.perf_counter() - t0, int(res.audio.size)


def cpu_sample_raw(n: int) -> np.ndarray:
    from oracle import iq_oracle as orc
    carriers = [dict(offset=o, amp=0.12, kind="fm", tone=700.0 + 150.0 * i, dev=2500.0) for i, o in enumerate(OFFSETS)]
    return orc.to_s16(orc.multi_carrier_capture(FS, n, carriers))


def cpu_run(n_frames: int, chunk: int, parallel: bool) -> tuple[float, int]:
    """(seconds, processes used) for all 5 targets over n_frames of the capture."""
    raw = cpu_sample_raw(n_frames)
    jobs = [(raw, off, chunk) for off in OFFSETS]
    if parallel:
        import multiprocessing as mp
        procs = max(1, min(len(jobs), os.cpu_count() or 1))
        with mp.get_context("fork").Pool(procs) as pool:
            t0 = time.perf_counter()
            pool.map(_cpu_one_target, jobs)
            return time.perf_counter() - t0, procs
    total = 0.0
    for j in jobs:       # as shipped: targets one after another (cli.py:683-710)
        dt, _ = _cpu_one_target(j)
        total += dt
    return total, 1


def _emit(line: dict) -> None:
    """The one JSON line goes to the REAL stdout; everything else that libraries print to fd 1 during the
    run (e.g. NCCL's version banner) was diverted to stderr by `_divert_stdout`."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def _divert_stdout() -> None:
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def main() -> None:
    _divert_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--seconds", type=float, default=SECONDS, help="capture length per GPU (default 60 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip the other_workloads block (cfg1, cfg3, cfg4, cfg5 points)")
    ap.add_argument("--peer-gather", action="store_true",
                    help="multi-GPU: push the audio into rank 0's memory with copy engines (sharding.PeerGather) "
                         "instead of the NCCL gather")
    ap.add_argument("--peer-writers", action="store_true",
                    help="multi-GPU: per-target writers fed by copy-engine pushes over peer memory (sharding.PeerWriters)")
    ap.add_argument("--nccl-writers", action="store_true",
                    help="multi-GPU: per-target writers over NCCL all_to_all_single (sharding.WriterExchange)")
    ap.add_argument("--serial-exchange", dest="overlap_exchange", action="store_false",
                    help="multi-GPU: run the exchange of step k before the kernels of step k+1 instead of under them")
    ap.add_argument("--sm-reserve", type=int, default=int(os.environ.get("IQ2A_BENCH_SM_RESERVE", SM_RESERVE)),
                    help="multi-GPU: SMs kept free of the persistent kernel for the NCCL gather (= NCCL channels)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 0)
    steps = max(args.steps, 1)

    from iq_to_audio_b200.processing import tune_chunk_size
    chunk = tune_chunk_size(FS, REQ_CHUNK)

    # ------------------------------------------------------------------ reference (CPU) arm
    if args.impl == "reference":
        if rank != 0:
            return
        n_chunks = 2
        n = n_chunks * chunk
        times = []
        procs = 1
        for _ in range(warmup):
            cpu_run(chunk // 4, chunk, True)
        for _ in range(steps):
            dt, procs = cpu_run(n, chunk, True)
            times.append(dt)
        dt = float(np.mean(times))
        val = n / dt / 1e6
        sample = f"first {n_chunks} chunks ({n} samples) of the workload per step, one process per target"
        line = {
            "metric": METRIC, "value": val, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "c128 transforms / f32 audio (numpy+scipy)", "data": "synthetic",
            "config": {"workload": workload_name(args.seconds), "chunk": chunk, "filter_block": 65_536},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        _emit(line)
        return

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist
    from iq_to_audio_b200.bank import ChannelBank

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the B200 arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # the streaming (e2e) leg is bound by host->device copies: keep this rank's threads and the pinned buffers it
    # allocates from here on next to the GPU's PCIe root
    from iq_to_audio_b200.numa import bind_to_device_numa
    placement = bind_to_device_numa(local_rank)
    if world > 1:
        # keep stdout to the single JSON line: NCCL's version/info banner goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        # the exchange of step k runs under the kernels of step k+1: small NCCL CTAs (256 threads) fit next to the
        # persistent channel-bank kernel on the SMs it leaves free (--sm-reserve); measured on 2/4/8 B200:
        # 2.71 / 2.75 / 3.16 ms per step, against 2.80 / 2.80 / 3.58 with the exchange serialised behind each step
        # and 3.46 / 3.41 / 3.22 with NCCL's defaults (profiles/r01_multi_gpu_exchange.md)
        if args.overlap_exchange:
            os.environ.setdefault("NCCL_NTHREADS", "256")
            os.environ.setdefault("NCCL_MAX_NCHANNELS", "32")
        dist.init_process_group("nccl", device_id=dev)

    n_seg = int(round(FS * args.seconds))
    n_seg = (n_seg // chunk) * chunk if n_seg >= chunk else n_seg     # segments start on the reference chunk grid
    d, fs_ch, targets = make_targets()
    bank = ChannelBank(FS, d, targets, codec="pcm_s16le", iq_order="iq", ref_chunk=chunk, device=local_rank)
    from iq_to_audio_b200 import sharding
    seg = sharding.plan_segments(world * n_seg, world, chunk, d, bank.halo, ["nfm"] * len(OFFSETS),
                                 sample_rate=FS, deemph_us=300.0)[rank]
    warm_rows, seg_begin, seg_end, first = seg.warmup_rows, seg.begin, seg.end, seg.first_frame
    capture = synth_capture_device(first, seg_end - first + d, dev, seed=1234 + rank)   # + one row of slack
    rows = bank.rows_in(seg_begin, seg_end)
    # two audio slots: the gather of step k (NVLink, rank 0's ingress) runs while step k+1 computes.  Preferred:
    # peer-memory push by copy engines (sharding.PeerGather); fallback: NCCL gather.
    comp = torch.cuda.Stream(device=dev)
    # Multi-GPU audio collection.  Default: per-target writers (target c is assembled on rank c % N), by whichever
    # transport is faster on this box, decided in the warm-up by timing a few steps of each (identical decision on
    # every rank): copy-engine pushes over peer memory (sharding.PeerWriters: no SMs, no NCCL in the data path) or
    # NCCL all_to_all_single under the next step's kernels on reserved SMs (sharding.WriterExchange).
    # --peer-writers / --nccl-writers / --peer-gather force one.
    peer = None
    auto = world > 1 and not (args.peer_gather or args.peer_writers or args.nccl_writers)
    if world > 1 and not args.nccl_writers:
        try:
            peer = sharding.PeerGather((bank.n_channels, rows), torch.float32, dev) if args.peer_gather else \
                sharding.PeerWriters(bank.n_channels, rows, torch.float32, dev)
        except Exception as exc:                      # symmetric memory unavailable on this box / build
            print(f"[bench] peer-memory transport unavailable ({exc!r}); using NCCL", file=sys.stderr)
            peer = None
        ok = torch.tensor([1 if peer is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            peer = None
    use_nccl = world > 1 and (peer is None or auto)
    audio_bufs = [torch.empty((bank.n_channels, rows), dtype=torch.float32, device=dev)
                  for _ in range(2 if world > 1 else 1)] if (world == 1 or use_nccl) else None
    xchg = sharding.WriterExchange(bank.n_channels, rows, torch.float32, dev) if use_nccl else None
    scheme = {"peer": peer is not None}

    def select(use_peer: bool) -> None:
        scheme["peer"] = use_peer
        # room for NCCL's CTAs next to the persistent kernel; the copy engines need none
        bank.set_sm_reserve(args.sm_reserve if (world > 1 and not use_peer and args.overlap_exchange) else 0)

    select(scheme["peer"])
    peer_no, flushed = [0], [0]
    step_no = [0]
    pending = [None, None]
    host_publish = []

    def resident_step():
        kk = step_no[0]
        step_no[0] += 1
        if scheme["peer"]:
            kk = peer_no[0]
            peer_no[0] += 1
            audio = peer.slot(kk)
            peer.before_compute(kk, comp)
            bank.process_resident_async(capture.data_ptr(), first, seg_end - first + d, seg_begin, seg_end,
                                        warmup_rows=warm_rows, dev_audio=audio.data_ptr(), out_stride=rows,
                                        stream=comp.cuda_stream)
            t_h = time.perf_counter()
            peer.publish(kk, comp)
            host_publish.append(time.perf_counter() - t_h)
            return
        k = kk % len(audio_bufs)
        audio = audio_bufs[k]
        if world == 1:
            # device-side ordering only, like the multi-GPU arm: the host enqueues ahead and the timed region ends with a
            # synchronize (the synchronous call left the GPU idle for the launch latency of every step)
            bank.process_resident_async(capture.data_ptr(), first, seg_end - first + d, seg_begin, seg_end,
                                        warmup_rows=warm_rows, dev_audio=audio.data_ptr(), out_stride=rows,
                                        stream=comp.cuda_stream)
            return
        # multi-GPU: device-side ordering only, the host runs ahead
        with torch.cuda.stream(comp):
            if args.overlap_exchange and pending[k] is not None:
                for wk in pending[k]:        # the exchange that last read this buffer (two steps ago)
                    wk.wait()
            bank.process_resident_async(capture.data_ptr(), first, seg_end - first + d, seg_begin, seg_end,
                                        warmup_rows=warm_rows, dev_audio=audio.data_ptr(), out_stride=rows,
                                        stream=comp.cuda_stream)
            works = xchg.exchange(k, audio)
            if args.overlap_exchange:
                pending[k] = works           # runs on NCCL's stream under the next step's kernels
            else:
                for wk in works:
                    wk.wait()                # orders `comp` behind the collective; does not block the host

    def drain():
        if peer is not None and peer_no[0] > flushed[0]:
            peer.flush(peer_no[0] - 1, comp)
            flushed[0] = peer_no[0]
        with torch.cuda.stream(comp):
            for k in range(len(pending)):
                for wk in pending[k] or ():
                    wk.wait()
                pending[k] = None
        comp.synchronize()

    def sync_all():
        drain()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    calibration = None
    if auto and peer is not None:
        calibration = {}
        for name, use_peer in (("peer_writers_ms", True), ("nccl_writers_ms", False)):
            select(use_peer)
            for _ in range(2):
                resident_step()
            sync_all()
            t_c = time.perf_counter()
            for _ in range(4):
                resident_step()
            drain()
            torch.cuda.synchronize()
            t = torch.tensor([(time.perf_counter() - t_c) * 1e3 / 4], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            calibration[name] = float(t.item())
            sync_all()
        select(calibration["peer_writers_ms"] <= calibration["nccl_writers_ms"])
    for _ in range(warmup):
        resident_step()
    sync_all()
    # what arrived at the writers is what the ranks produced: per-(rank, target) checksums of the last warm-up step
    gather_ok = None
    via_peer = scheme["peer"]
    if world > 1:
        last = (peer_no[0] if via_peer else step_no[0]) - 1
        src = peer.slot(last) if via_peer else audio_bufs[last % 2]
        mine = src.double().abs().sum(dim=1)                         # [C]
        sums = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(sums, mine)
        sums = torch.stack(sums)                                     # [world, C]
        if via_peer and isinstance(peer, sharding.PeerWriters):
            ok = all(float((blk.double().abs().sum(dim=1) - sums[:, c]).abs().max()) <= 1e-9 * max(1.0, float(sums[:, c].max()))
                     for c, blk in peer.result(last).items())
        elif via_peer:
            ok = rank != 0 or all(abs(float(g.double().abs().sum()) - float(sums[r].sum())) <= 1e-9 * max(1.0, float(sums[r].sum()))
                                  for r, g in enumerate(peer.result(last)))
        else:
            ok = all(float((blk.double().abs().sum(dim=1) - sums[:, c]).abs().max()) <= 1e-9 * max(1.0, float(sums[:, c].max()))
                     for c, blk in xchg.result(last % 2).items())
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        gather_ok = bool(int(flag.item()))
        if not gather_ok:
            raise SystemExit("exchanged audio does not match what the ranks produced")
    bank.set_timing(True)
    launches0 = bank.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local_rank)               # NVML initialisation takes a rank-dependent time: not in the bracket
    with sampler as clocks:
        if world > 1:
            # the bracket of the timed region: the ranks enter together (any skew here is charged to the early rank,
            # which then waits for the others at the writers' barrier: 2.04-2.56 ms per rank with five steps)
            dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        t0 = time.perf_counter()
        for _ in range(steps):
            resident_step()
        drain()
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
    dev_ms = e0.elapsed_time(e1)
    if world > 1 and xchg is not None and os.environ.get("IQ2A_BENCH_DEBUG"):
        # the exchange alone, nothing else running: what has to hide under a step
        sync_all()
        tx = []
        for _ in range(3):
            dist.barrier(); torch.cuda.synchronize()
            t_a = time.perf_counter()
            for wk in xchg.exchange(0, audio_bufs[0]):
                wk.wait()
            torch.cuda.synchronize()
            tx.append((time.perf_counter() - t_a) * 1e3)
        print(f"[bench] rank {rank}: exchange alone ms {[round(v, 3) for v in tx]}; step wall {wall * 1e3 / steps:.3f}", file=sys.stderr)
    if host_publish and os.environ.get("IQ2A_BENCH_DEBUG"):
        print(f"[bench] rank {rank}: host ms in publish(): {[round(v * 1e3, 3) for v in host_publish[-steps:]]} wall {wall * 1e3:.3f} dev {dev_ms:.3f}", file=sys.stderr)
    # the bank launches on its own stream: the wall clock bracketed by synchronize is the step time;
    # the CUDA events on torch's stream only bound the gather
    step_ms = max(dev_ms, wall * 1e3) / steps
    timing = bank.get_timing()
    launches = bank.launches - launches0
    bank.set_timing(False)
    sync_all()
    per_rank = None
    if world > 1:
        mine = torch.tensor([step_ms, timing["channelize_ms"] / max(timing["calls"], 1), float(placement.get("numa_node") if placement.get("numa_node") is not None else -1)],
                            dtype=torch.float64, device=dev)
        every = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        per_rank = {"step_ms": [round(float(v[0]), 4) for v in every], "kernel_ms": [round(float(v[1]), 4) for v in every],
                    "numa_node": [int(v[2]) for v in every]}
        step_ms = max(per_rank["step_ms"])

    # ---- e2e: pinned host capture streamed through process_chunk -----------------------------
    e2e = None
    if not args.no_e2e:
        host = torch.empty(2 * n_seg, dtype=torch.int16).pin_memory()
        host.copy_(capture[2 * (seg_begin - first):2 * (seg_begin - first) + 2 * n_seg])
        host_np = host.numpy()
        nchunks = (n_seg + chunk - 1) // chunk
        bytes_out = 0

        batch = max(1, min(8, -(-49_152 * d // chunk)))          # reference chunks per GPU call (pipeline.py's rule)

        def e2e_step():
            nonlocal bytes_out
            bank.reset()
            bytes_out = 0
            span = batch * chunk
            views = (host_np[2 * s:2 * min(s + span, n_seg)] for s in range(0, n_seg, span))
            for r in bank.stream(views, chunk_frames=chunk):
                bytes_out += r.audio.nbytes + r.clipped.nbytes
        e2e_step()
        sync_all()
        reps = max(1, min(steps, 3))
        with sampler:
            t0 = time.perf_counter()
            for _ in range(reps):
                e2e_step()
            torch.cuda.synchronize()
            e2e_ms = (time.perf_counter() - t0) * 1e3 / reps
        e2e_rank_ms = [e2e_ms]
        if world > 1:
            t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
            every = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(every, t)
            e2e_rank_ms = [float(v.item()) for v in every]
            e2e_ms = max(e2e_rank_ms)
        # what the host can feed: the same pinned buffer copied to the device by every rank at once, nothing else
        # running (the ceiling of the e2e number on this box: PCIe link per GPU x what the host memory system sustains)
        sink = torch.empty(1 << 27, dtype=torch.int16, device=dev)                    # 256 MiB
        piece = host[:min(host.numel(), sink.numel())]
        reps_c = 8
        sink[:piece.numel()].copy_(piece, non_blocking=True)
        sync_all()
        t0 = time.perf_counter()
        for _ in range(reps_c):
            sink[:piece.numel()].copy_(piece, non_blocking=True)
        torch.cuda.synchronize()
        copy_gbps = reps_c * piece.numel() * 2 / (time.perf_counter() - t0) / 1e9
        copy_rank = [copy_gbps]
        if world > 1:
            t = torch.tensor([copy_gbps], dtype=torch.float64, device=dev)
            every = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(every, t)
            copy_rank = [float(v.item()) for v in every]
        del sink
        e2e = {"value": world * n_seg / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(4 * n_seg),
               "h2d_copy_only_GBps": [round(v, 2) for v in copy_rank],
               "d2h_bytes_per_step": int(bytes_out), "ms_per_step": e2e_ms,
               "per_rank_h2d_GBps": [round(4 * n_seg / (m * 1e-3) / 1e9, 2) for m in e2e_rank_ms],
               "api": f"ChannelBank.stream(chunk_frames={chunk}): submit/collect of {batch} reference chunks per call, 2 calls in flight, pinned host input"}

    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peak = float(json.loads(peaks_path.read_text())["hbm_gbs"])
        peak_src = "MEASURED_PEAKS.json hbm_gbs (burst copy)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"

    # ---- the other BASELINE configurations (cfg1, cfg3, cfg4 time-sharded over all ranks, two cfg5 points) ----
    others = None
    if not args.no_others:
        del capture
        torch.cuda.empty_cache()
        others = {}
        for key, desc, fs_w, secs, specs, all_ranks in other_workload_specs():
            if not all_ranks and rank != 0:
                continue
            try:
                others[key] = run_other_workload(key, desc, fs_w, secs, specs, dev, rank if all_ranks else 0,
                                                 world if all_ranks else 1, peak, cpu=not args.no_cpu_baseline)
            except Exception as exc:                       # a failed extra must not take the headline number with it
                others[key] = {"workload": desc, "error": repr(exc)[:300]}
        if world > 1:
            dist.barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (k_channelize) ---------------------------------------
    n_in = seg_end - max(0, seg_begin - warm_rows * d)                 # samples one launch turns into channel rows
    b_alg = 4.0 + sum(4.0 / d for _ in OFFSETS)                        # int16 in once + f32 audio out
    chan_ms = timing["channelize_ms"] / max(timing["calls"], 1)
    achieved = n_in * b_alg / (chan_ms * 1e-3) / 1e9 if chan_ms > 0 else None
    traffic = None
    prof = ROOT / "profiles" / "channelize_traffic.json"
    if prof.exists():
        try:
            traffic = json.loads(prof.read_text()).get("dram_bytes_per_sample", None)
            traffic = traffic * n_in if traffic is not None else None
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": {5: "k_channelize5<5>", 4: "k_channelize2<5,2>"}.get(bank.kernel_generation, "k_channelize<512,5,s16>"), "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                "algorithmic_bytes_per_sample": b_alg, "kernel_ms": chan_ms, "tail_ms": timing["tail_ms"] / max(timing["calls"], 1),
                "peak_source": peak_src}
    # second reading of the same kernel: it is FP32-issue bound, not HBM bound.  Flops per input
    # sample of the exact algorithm: (5 log2 M forward + 8 C multiply-accumulate) per transform point, M / Ld
    # transform points per new sample; peak = SMs x 128 FP32 lanes x 2 x SM clock measured during the run.
    m_fft, ld = bank.fft_size, bank.rows_per_block
    flop_per_sample = (5.0 * np.log2(m_fft) + 8.0 * len(OFFSETS)) * m_fft / ld
    clk = clocks.summary().get("sm_mhz") or 1965
    fp32_peak = 148 * 128 * 2 * clk * 1e6 / 1e12
    if chan_ms > 0:
        tf = n_in * flop_per_sample / (chan_ms * 1e-3) / 1e12
        roofline["fp32"] = {"flop_per_sample": flop_per_sample, "achieved": tf, "peak": fp32_peak, "unit": "TFLOP/s",
                            "frac": tf / fp32_peak,
                            "note": "all-FMA peak; the transform butterflies are adds (1 flop per lane-slot), so ~50 % is the ceiling for them"}

    cpu = None
    if not args.no_cpu_baseline:
        n_cpu = 2 * chunk
        dt, _ = cpu_run(n_cpu, chunk, False)
        cpu = {"value": n_cpu / dt / 1e6, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"first 2 chunks ({n_cpu} samples), 5 targets sequential as cli.py:683 runs them; host has {os.cpu_count()} cores"}

    value = world * n_seg / (step_ms * 1e-3) / 1e6
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (f64 NCO phase and audio recurrences)", "data": "synthetic",
        "config": {"workload": workload_name(args.seconds), "samples_per_gpu": n_seg, "chunk": chunk,
                   "fft_size": bank.fft_size, "kernel_generation": bank.kernel_generation, "hop": bank.hop, "decimation": d, "parallelism": f"time-shard x{world}", "sm_reserved_for_exchange": (args.sm_reserve if (not via_peer and args.overlap_exchange) else 0) if world > 1 else 0,
                   "exchange_overlapped": bool(world > 1 and (via_peer or args.overlap_exchange)),
                   "gather": "none" if world == 1 else ("per-target writers: copy-engine push over peer memory, target c on rank c % N" if (via_peer and isinstance(peer, sharding.PeerWriters)) else "peer-memory push to rank 0 (copy engines)" if via_peer else "per-target writers: all_to_all_single over NCCL, target c on rank c % N"),
                   "gather_calibration": calibration,
                   "gather_verified": gather_ok,
                   "l2": f"input {4 * n_seg / 1e9:.2f} GB per GPU >> 126 MB L2, read once per step"},
        "x_realtime": value * 1e6 / FS,
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "other_workloads": others,
        "per_rank": per_rank, "host_placement": placement,
        "clocks": clocks.summary(),
    }
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
