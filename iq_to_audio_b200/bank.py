"""ChannelBank: all targets of a capture in one pass on the GPU.

The reference runs one `ProcessingPipeline` per target, sequentially
(``src/iq_to_audio/cli.py:683-710``); each re-reads and re-mixes the capture.  A
`ChannelBank` owns C targets and performs, per chunk, what C iterations of the
reference loop body would (``processing.py:1088-1147``): unpack, mix, channel
filter, decimate, demodulate, peak/clip -- through ``libiq2a_b200.so``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib


@dataclass
class Target:
    """One channel of a bank (the per-target part of the reference's ProcessingConfig)."""
    freq_offset: float                 # target_freq - center_freq, Hz
    taps: np.ndarray                   # design_channel_filter(...) output
    mix_sign: int = 1
    mode: str = "nfm"
    deemph_us: float = 300.0
    agc_enabled: bool = True


@dataclass
class ChunkResult:
    audio: np.ndarray                  # float32 [C, n] decoder output (pre-clip)
    clipped: np.ndarray                # float32 [C, n] what the writer pipes to the encoder
    baseband: np.ndarray | None        # complex64 [C, n] decimated channel samples (if requested)
    rms_dbfs: np.ndarray               # float64 [C] (one reference chunk per call) or [C, W] (stream(..., chunk_frames=))
    count: int                         # n
    window_counts: np.ndarray | None = None   # int64 [W]: rows of every reference chunk the call covered


class ChannelBank:
    def __init__(self, sample_rate: float, decimation: int, targets: list[Target], *,
                 codec: str = "pcm_s16le", iq_order: str = "iq", ref_chunk: int = 1_048_576,
                 fft_size: int = 0, device: int = 0):
        if codec not in _lib.CODEC_IDS:
            raise ValueError(f"Unsupported input codec '{codec}'")
        if iq_order not in _lib.ORDER_IDS:
            raise ValueError(f"Unsupported iq_order '{iq_order}'")      # processing.py:269-270
        if not targets:
            raise ValueError("at least one target is required")
        self._lib = _lib.load()
        self.sample_rate = float(sample_rate)
        self.decimation = max(1, int(decimation))
        self.targets = list(targets)
        self.codec = codec
        self.codec_id = _lib.CODEC_IDS[codec]
        self.iq_order = iq_order
        self.ref_chunk = int(ref_chunk)
        self.device = int(device)
        descs = (_lib.ChannelDesc * len(targets))()
        self._taps_keepalive = []
        for d, t in zip(descs, targets):
            mode = t.mode.lower()
            if mode not in _lib.MODE_IDS:
                raise ValueError(f"Unsupported demod mode '{mode}'.")       # decoders/__init__.py:24
            taps = np.ascontiguousarray(t.taps, dtype=np.float64)
            self._taps_keepalive.append(taps)
            d.freq_offset_hz = float(t.freq_offset)
            d.mix_sign = int(t.mix_sign)
            d.mode = _lib.MODE_IDS[mode]
            d.taps = taps.ctypes.data_as(C.POINTER(C.c_double))
            d.ntaps = int(taps.size)
            d.agc_enabled = 1 if t.agc_enabled else 0
            d.deemph_us = float(t.deemph_us)
        cfg = _lib.BankConfig(self.sample_rate, self.decimation, self.codec_id, _lib.ORDER_IDS[iq_order],
                              len(targets), int(fft_size), self.ref_chunk, self.device, 0)
        handle = C.c_void_p()
        _lib.check(self._lib.iq2a_bank_create(C.byref(cfg), descs, C.byref(handle)))
        self._h = handle
        info = _lib.BankInfo()
        _lib.check(self._lib.iq2a_bank_info_get(self._h, C.byref(info)))
        self.fft_size = info.fft_size
        self.overlap_rows = info.overlap_rows
        self.rows_per_block = info.rows_per_block
        self.hop = info.hop
        self.halo = info.halo
        self.fs_channel = info.fs_channel
        self.kernel_generation = info.kernel_generation
        self.n_channels = len(targets)

    # -- lifetime ---------------------------------------------------------------------------
    def close(self) -> None:
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.iq2a_bank_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self) -> "ChannelBank":
        return self

    def __exit__(self, *exc) -> None:
        self.close()

    def reset(self) -> None:
        _lib.check(self._lib.iq2a_bank_reset(self._h))

    # -- carried state ------------------------------------------------------------------------
    def get_state(self) -> tuple[list[dict], int]:
        arr = (_lib.ChannelState * self.n_channels)()
        n = C.c_int64(0)
        _lib.check(self._lib.iq2a_bank_get_state(self._h, arr, C.byref(n)))
        keys = ("prev_re", "prev_im", "dc_x", "dc_y", "deemph_z", "peak")
        return [{k: getattr(s, k) for k in keys} for s in arr], int(n.value)

    def set_state(self, states: list[dict]) -> None:
        arr = (_lib.ChannelState * self.n_channels)()
        for s, d in zip(arr, states):
            for k, v in d.items():
                setattr(s, k, v)
        _lib.check(self._lib.iq2a_bank_set_state(self._h, arr))

    @property
    def peaks(self) -> list[float]:
        return [s["peak"] for s in self.get_state()[0]]

    @property
    def launches(self) -> int:
        n = C.c_int64(0)
        _lib.check(self._lib.iq2a_bank_launch_count(self._h, C.byref(n)))
        return int(n.value)

    def set_sm_reserve(self, n_sm: int) -> None:
        """Keep `n_sm` SMs free of the persistent channel-bank kernel (room for a concurrent NCCL gather)."""
        _lib.check(self._lib.iq2a_bank_set_sm_reserve(self._h, int(n_sm)))

    def set_timing(self, enable: bool = True) -> None:
        _lib.check(self._lib.iq2a_bank_set_timing(self._h, 1 if enable else 0))

    def get_timing(self) -> dict:
        c, h, t, n = C.c_double(0), C.c_double(0), C.c_double(0), C.c_int64(0)
        _lib.check(self._lib.iq2a_bank_get_timing(self._h, C.byref(c), C.byref(h), C.byref(t), C.byref(n)))
        return {"channelize_ms": c.value, "head_ms": h.value, "tail_ms": t.value, "calls": int(n.value)}

    def g_table(self) -> np.ndarray:
        out = np.empty((self.n_channels, self.decimation, self.fft_size), dtype=np.complex64)
        _lib.check(self._lib.iq2a_bank_copy_gtable(self._h, out.ctypes.data, out.size))
        return out

    # -- streaming: one reference chunk per call -------------------------------------------------
    def frames_in(self, raw) -> tuple[int, int]:
        """(address, n_frames) of a raw PCM buffer in the bank's codec; a trailing partial
        frame is dropped (processing.py:253-256)."""
        fb = _lib.FRAME_BYTES[self.codec_id]
        if isinstance(raw, np.ndarray):
            buf = raw if raw.flags.c_contiguous else np.ascontiguousarray(raw)
            self._keep = buf
            return buf.ctypes.data, buf.nbytes // fb
        mv = memoryview(raw)
        arr = np.frombuffer(mv, dtype=np.uint8)
        self._keep = arr
        return arr.ctypes.data, arr.nbytes // fb

    def process_chunk(self, raw, *, want_baseband: bool = False, n_frames: int | None = None) -> ChunkResult:
        addr, n = self.frames_in(raw)
        if n_frames is not None:
            n = min(n, int(n_frames))
        cap = max(1, (n + self.decimation - 1) // self.decimation + 1)
        cc = self.n_channels
        audio = np.empty((cc, cap), dtype=np.float32)
        clipped = np.empty((cc, cap), dtype=np.float32)
        bb = np.empty((cc, cap), dtype=np.complex64) if want_baseband else None
        rms = np.zeros(cc, dtype=np.float64)
        n_out = C.c_int64(0)
        _lib.check(self._lib.iq2a_bank_process_chunk(
            self._h, addr, n, audio.ctypes.data, clipped.ctypes.data, _lib.ptr(bb), cap, C.byref(n_out),
            rms.ctypes.data_as(C.POINTER(C.c_double))))
        k = int(n_out.value)
        return ChunkResult(audio[:, :k], clipped[:, :k], None if bb is None else bb[:, :k], rms, k)

    def stream(self, chunks, *, want_baseband: bool = False, want_audio: bool = True, want_clipped: bool = True,
               chunk_frames: int | None = None):
        """Pipelined form of `process_chunk` over an iterable of raw PCM chunks: chunk k+1 is
        submitted (H2D copy queued) before chunk k's results are collected, so the copy of the next
        chunk overlaps the kernels of the current one.  Yields one `ChunkResult` per chunk, in order.

        `chunk_frames`: every item of `chunks` may hold SEVERAL consecutive reference chunks of that many frames
        (the last one shorter); phase wraps, AGC restarts and statistics stay per reference chunk
        (iq2a_bank_submit_chunks) while the kernels see the whole item -- at high decimation one reference chunk
        is too few rows to fill the GPU.  `rms_dbfs` is then [C, W] and `window_counts` [W]."""
        want = (1 if want_audio else 0) | (2 if want_clipped else 0) | (4 if want_baseband else 0)
        pending = []          # [(keepalive, n_frames)]
        cf = int(chunk_frames) if chunk_frames else 0

        def collect(n):
            cap = max(1, (n + self.decimation - 1) // self.decimation + 1)
            cc = self.n_channels
            audio = np.empty((cc, cap), dtype=np.float32) if want_audio else None
            clipped = np.empty((cc, cap), dtype=np.float32) if want_clipped else None
            bb = np.empty((cc, cap), dtype=np.complex64) if want_baseband else None
            wcap = max(1, (n + cf - 1) // cf) if cf else 1
            rms = np.zeros((cc, wcap), dtype=np.float64)
            wrows = np.zeros(wcap, dtype=np.int64)
            n_out, n_win = C.c_int64(0), C.c_int64(0)
            _lib.check(self._lib.iq2a_bank_collect_chunks(
                self._h, _lib.ptr(audio), _lib.ptr(clipped), _lib.ptr(bb), cap, C.byref(n_out),
                rms.ctypes.data_as(C.POINTER(C.c_double)), wcap, wrows.ctypes.data_as(C.POINTER(C.c_int64)),
                C.byref(n_win)))
            k, w = int(n_out.value), int(n_win.value)
            return ChunkResult(None if audio is None else audio[:, :k], None if clipped is None else clipped[:, :k],
                               None if bb is None else bb[:, :k], rms[:, :w] if cf else rms[:, 0], k,
                               wrows[:w] if cf else None)

        for raw in chunks:
            addr, n = self.frames_in(raw)
            keep = self._keep
            _lib.check(self._lib.iq2a_bank_submit_chunks(self._h, addr, n, cf, want))
            pending.append((keep, n))
            if len(pending) == 2:
                _, n0 = pending.pop(0)
                yield collect(n0)
        while pending:
            _, n0 = pending.pop(0)
            yield collect(n0)

    # -- resident: a whole segment already in HBM --------------------------------------------------
    def process_resident(self, dev_ptr: int, first_frame: int, n_frames: int, seg_begin: int, seg_end: int, *,
                         warmup_rows: int = 0, dev_audio: int | None = None, dev_clipped: int | None = None,
                         dev_baseband: int | None = None, out_stride: int = 0, want_rms: bool = False):
        n_out = C.c_int64(0)
        rms = None
        rms_cap = 0
        if want_rms:
            rms_cap = (seg_end - 1) // self.ref_chunk - seg_begin // self.ref_chunk + 1 if seg_end > seg_begin else 0
            rms = np.zeros((self.n_channels, max(rms_cap, 1)), dtype=np.float64)
        _lib.check(self._lib.iq2a_bank_process_resident(
            self._h, dev_ptr, first_frame, n_frames, seg_begin, seg_end, warmup_rows, dev_audio, dev_clipped,
            dev_baseband, out_stride, C.byref(n_out),
            None if rms is None else rms.ctypes.data_as(C.POINTER(C.c_double)), rms_cap))
        return int(n_out.value), rms

    def process_resident_async(self, dev_ptr: int, first_frame: int, n_frames: int, seg_begin: int, seg_end: int,
                               *, warmup_rows: int = 0, dev_audio: int | None = None,
                               dev_clipped: int | None = None, dev_baseband: int | None = None,
                               out_stride: int = 0, stream: int = 0) -> None:
        _lib.check(self._lib.iq2a_bank_process_resident_async(
            self._h, dev_ptr, first_frame, n_frames, seg_begin, seg_end, warmup_rows, dev_audio, dev_clipped,
            dev_baseband, out_stride, stream))

    def rows_in(self, seg_begin: int, seg_end: int) -> int:
        d = self.decimation
        return (seg_end + d - 1) // d - (seg_begin + d - 1) // d
