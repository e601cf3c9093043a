"""CPU oracle for the channelize-and-demodulate path.  TEST INFRASTRUCTURE ONLY.

This module restates, in plain numpy/scipy, the arithmetic of the reference's
streaming hot path so that the CUDA implementation can be checked against it on
machines where ``/root/reference`` does not exist (the GPU box).  Nothing under
``iq_to_audio_b200/`` may import it; only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs do.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the unmodified
reference classes from ``/root/reference/src`` (stub ``soundfile`` only) and
records their outputs; ``tests/test_oracle_golden.py`` checks every function
below against those fixtures, bit for bit where the arithmetic is float32 /
integer, and to 0 ulp where the same numpy/scipy calls are used.

Third-party arithmetic that is part of the reference's result (not vendored in
the reference tree): scipy.fft (pocketfft, complex128), scipy.signal.firwin /
kaiser_beta / lfilter, numpy ufuncs -- the oracle calls the very same functions
(scipy >= 1.10, numpy >= 2), see SURVEY.md section 8(c).

All ``ref:`` citations are relative to ``/root/reference/src/iq_to_audio/``.

The restatement is organised as free functions over small explicit state
records (one record per carried quantity) rather than as stage objects, so a
time-sharded or chunk-split run can hand the state across explicitly.
"""
from __future__ import annotations

import ctypes
import math
import os
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np
from scipy import fft as _sfft
from scipy import signal as _ssig

TWO_PI = 2.0 * np.pi

# --------------------------------------------------------------------------
# K1 + K2: sample unpack and IQ order / polarity
# --------------------------------------------------------------------------

#: bytes per complex frame on disk for each codec (ref: input_formats.py:45-94)
FRAME_BYTES = {"pcm_u8": 2, "pcm_s16le": 4, "pcm_f32le": 8}
IQ_ORDERS = ("iq", "qi", "iq_inv", "qi_inv")


def unpack_interleaved(raw: bytes | np.ndarray, codec: str) -> np.ndarray:
    """Raw PCM bytes -> interleaved float32, as the ffmpeg decode leg does.

    ref: processing.py:143-158 asks ffmpeg for ``-f f32le -ac 2``; the sample
    format conversion is libswresample's: s16 -> x/32768, u8 -> (x-128)/128,
    f32 passthrough (verified against libswresample 6.1.100).
    A trailing partial frame is dropped (ref: processing.py:253-256).
    """
    buf = np.frombuffer(raw, dtype=np.uint8) if not isinstance(raw, np.ndarray) else raw.view(np.uint8).ravel()
    fb = FRAME_BYTES[codec]
    usable = (buf.size // fb) * fb
    buf = buf[:usable]
    if codec == "pcm_s16le":
        return buf.view("<i2").astype(np.float32) * np.float32(1.0 / 32768.0)
    if codec == "pcm_u8":
        return (buf.astype(np.float32) - np.float32(128.0)) * np.float32(1.0 / 128.0)
    if codec == "pcm_f32le":
        return buf.view("<f4").astype(np.float32, copy=True)
    raise ValueError(f"unsupported codec {codec!r}")


def order_iq(interleaved: np.ndarray, iq_order: str) -> np.ndarray:
    """Interleaved float32 pairs -> complex64 with order/polarity applied.

    ref: processing.py:261-279 (``_extract_iq`` and the complex build).
    """
    if iq_order not in IQ_ORDERS:
        raise ValueError(f"Unsupported iq_order '{iq_order}'")
    first = interleaved[0::2]
    second = interleaved[1::2]
    if iq_order.startswith("iq"):
        i_part, q_part = first, second
    else:
        i_part, q_part = second, first
    if iq_order.endswith("_inv"):
        q_part = -q_part
    out = np.empty(i_part.size, dtype=np.complex64)
    out.real = i_part
    out.imag = q_part
    return out


# --------------------------------------------------------------------------
# K3: NCO mixer
# --------------------------------------------------------------------------

@dataclass
class NcoState:
    """ref: processing.py:285-287 -- phase starts at 0, inc = -2*pi*f_off/fs."""
    increment: float
    phase: float = 0.0

    @classmethod
    def for_offset(cls, freq_offset_hz: float, sample_rate: float) -> "NcoState":
        return cls(increment=-2.0 * np.pi * freq_offset_hz / sample_rate)


def nco_advance(phase: float, increment: float, sign: int, count: int) -> float:
    """Phase carried to the next chunk (ref: processing.py:295)."""
    return (phase + sign * increment * count) % TWO_PI


def nco_mix(st: NcoState, x: np.ndarray, sign: int) -> np.ndarray:
    """ref: processing.py:289-297.  float64 phase ramp, LO rounded to complex64,
    complex64 product; phase wrapped modulo 2*pi once per call."""
    if x.size == 0:
        return x
    ramp = np.arange(x.size, dtype=np.float64)
    arg = st.phase + sign * st.increment * ramp
    lo = np.exp(1j * arg).astype(np.complex64)
    st.phase = nco_advance(st.phase, st.increment, sign, x.size)
    return np.asarray(x.astype(np.complex64, copy=False) * lo, dtype=np.complex64)


# --------------------------------------------------------------------------
# K4: overlap-save FIR (complex128 transforms, complex64 in/out)
# --------------------------------------------------------------------------

@dataclass
class FirState:
    taps: np.ndarray            # float64 [ntaps]
    block: int
    nfft: int = 0
    spec: np.ndarray = field(default=None, repr=False)   # complex128 [nfft]
    hist: np.ndarray = field(default=None, repr=False)   # complex64 [ntaps-1]

    def __post_init__(self) -> None:
        if self.block <= 0:
            raise ValueError("block_size must be positive")
        nt = len(self.taps)
        # ref: processing.py:310 -- next power of two >= block + ntaps - 1
        self.nfft = 1 << math.ceil(math.log2(self.block + nt - 1))
        padded = np.zeros(self.nfft, dtype=np.complex128)
        padded[:nt] = self.taps
        self.spec = np.asarray(_sfft.fft(padded))
        self.hist = np.zeros(nt - 1, dtype=np.complex64)


def fir_overlap_save(st: FirState, x: np.ndarray) -> np.ndarray:
    """ref: processing.py:325-346.  Output length == input length; history is
    the last ntaps-1 *input* samples; segmentation restarts at every call."""
    if x.size == 0:
        return x
    x = x.astype(np.complex64)
    ov = len(st.taps) - 1
    pieces = []
    pos = 0
    while pos < x.size:
        seg = x[pos:pos + st.block]
        pos += seg.size
        buf = np.concatenate([st.hist, seg]).astype(np.complex128)
        if buf.size < st.nfft:
            buf = np.pad(buf, (0, st.nfft - buf.size))
        prod = np.asarray(_sfft.ifft(np.asarray(_sfft.fft(buf)) * st.spec))
        pieces.append(prod[ov:ov + seg.size].astype(np.complex64))
        if ov:
            if seg.size >= ov:
                st.hist = seg[-ov:].copy()
            else:
                st.hist = np.concatenate([st.hist[seg.size:], seg]).astype(np.complex64)
    return np.concatenate(pieces)


def fir_direct_f64(taps: np.ndarray, x: np.ndarray, at: np.ndarray) -> np.ndarray:
    """Exact causal FIR evaluated only at sample indices ``at`` (zero history),
    float64/complex128 dot products.  Independent cross-check of K4 (SURVEY 7.3:
    matches the overlap-save result to ~3.5e-8)."""
    xs = x.astype(np.complex128)
    nt = len(taps)
    out = np.empty(len(at), dtype=np.complex128)
    for i, n in enumerate(at):
        lo = max(0, n - nt + 1)
        seg = xs[lo:n + 1][::-1]
        out[i] = np.dot(taps[:seg.size], seg)
    return out


# --------------------------------------------------------------------------
# K5: decimator (global phase)
# --------------------------------------------------------------------------

@dataclass
class DecimState:
    factor: int
    offset: int = 0

    def __post_init__(self) -> None:
        self.factor = max(1, self.factor)   # ref: processing.py:351


def decimate(st: DecimState, x: np.ndarray) -> np.ndarray:
    """ref: processing.py:354-360 -- keep global indices == 0 (mod factor)."""
    if st.factor == 1 or x.size == 0:
        return x
    first = (-st.offset) % st.factor
    kept = x[first::st.factor]
    st.offset = (st.offset + x.size) % st.factor
    return kept


def decimated_count(n_start: int, n_end: int, factor: int) -> int:
    """Number of global indices that are multiples of ``factor`` in [n_start, n_end)."""
    f = max(1, factor)
    return (n_end + f - 1) // f - (n_start + f - 1) // f


# --------------------------------------------------------------------------
# K6 / K7: NFM discriminator and de-emphasis
# --------------------------------------------------------------------------

@dataclass
class DiscState:
    prev: np.complex64 = np.complex64(1 + 0j)     # ref: decoders/nfm.py:15


def fm_discriminate(st: DiscState, s: np.ndarray) -> np.ndarray:
    """ref: decoders/nfm.py:17-24 -- angle(s[n] * conj(s[n-1])), float32, unscaled."""
    if s.size == 0:
        return np.empty(0, dtype=np.float32)
    lagged = np.concatenate(([st.prev], s[:-1]))
    out = np.asarray(np.angle(s * np.conj(lagged)), dtype=np.float32)
    st.prev = s[-1]
    return out


@dataclass
class DeemphState:
    alpha: float
    beta: float
    z: float = 0.0

    @classmethod
    def design(cls, tau_us: float, fs_channel: float) -> "DeemphState":
        # ref: decoders/nfm.py:40-47
        tau = max(tau_us * 1e-6, 1e-6)
        a = math.exp(-1.0 / (fs_channel * tau))
        return cls(alpha=a, beta=1.0 - a, z=0.0)


def deemphasis(st: DeemphState, x: np.ndarray) -> np.ndarray:
    """ref: decoders/nfm.py:49-62 -- y[n] = beta*x[n] + alpha*y[n-1] via lfilter
    (direct form II transposed, float64), output cast to float32, carried state
    is lfilter's zf[0] (= alpha * y_last)."""
    if x.size == 0:
        return x
    y, zf = _ssig.lfilter(np.array([st.beta]), np.array([1.0, -st.alpha]),
                          x.astype(np.float32, copy=False), zi=np.array([st.z]))
    st.z = float(np.asarray(zf, dtype=np.float64)[0])
    return np.asarray(y, dtype=np.float32)


# --------------------------------------------------------------------------
# K9 / K11: DC blocker and AGC (float32 sequential loops, C helper)
# --------------------------------------------------------------------------

_SEQ = None


def _seq_lib():
    """Load oracle/_build/libseqf32.so (built by oracle/Makefile or __graft_entry__.build())."""
    global _SEQ
    if _SEQ is None:
        here = Path(__file__).resolve().parent
        so = here / "_build" / "libseqf32.so"
        if not so.exists():
            import subprocess
            subprocess.run(["make", "-C", str(here)], check=True, capture_output=True)
        lib = ctypes.CDLL(os.fspath(so))
        fp = ctypes.POINTER(ctypes.c_float)
        dp = ctypes.POINTER(ctypes.c_double)
        lib.dc_block_f32.argtypes = [fp, fp, ctypes.c_size_t, ctypes.c_double, dp, dp]
        lib.dc_block_f32.restype = None
        lib.agc_f32.argtypes = [fp, fp, ctypes.c_size_t, ctypes.c_double, ctypes.c_double]
        lib.agc_f32.restype = None
        _SEQ = lib
    return _SEQ


@dataclass
class DcBlockState:
    radius: float = 0.995        # ref: decoders/common.py:9
    x_prev: float = 0.0
    y_prev: float = 0.0


def dc_block(st: DcBlockState, x: np.ndarray) -> np.ndarray:
    """ref: decoders/common.py:16-30 (float32 recurrence y = x - x1 + r*y1)."""
    if x.size == 0:
        return x
    xin = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(xin)
    xp = ctypes.c_double(st.x_prev)
    yp = ctypes.c_double(st.y_prev)
    fp = ctypes.POINTER(ctypes.c_float)
    _seq_lib().dc_block_f32(xin.ctypes.data_as(fp), out.ctypes.data_as(fp), xin.size,
                            st.radius, ctypes.byref(xp), ctypes.byref(yp))
    st.x_prev, st.y_prev = xp.value, yp.value
    return out


def dc_block_pyloop(st: DcBlockState, x: np.ndarray) -> np.ndarray:
    """Same recurrence as a literal per-sample Python loop over np.float32
    scalars (small inputs only) -- used to pin the C helper."""
    out = np.empty(x.size, dtype=np.float32)
    xp, yp, r = st.x_prev, st.y_prev, st.radius
    for k, v in enumerate(x.astype(np.float32, copy=False)):
        cur = v - xp + r * yp
        out[k] = cur
        xp, yp = v, cur
    st.x_prev, st.y_prev = float(xp), float(yp)
    return out


AGC_TARGET = 10.0 ** (-12.0 / 20.0)      # ref: decoders/ssb.py:21,33
AGC_DECAY = 0.001                        # ref: decoders/ssb.py:22
AGC_FLOOR = 1e-6                         # ref: decoders/ssb.py:76


def agc(x: np.ndarray, target: float = AGC_TARGET, decay: float = AGC_DECAY) -> np.ndarray:
    """ref: decoders/ssb.py:67-80 -- gain restarts at 1.0 on every call."""
    if x.size == 0:
        return x
    xin = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(xin)
    fp = ctypes.POINTER(ctypes.c_float)
    _seq_lib().agc_f32(xin.ctypes.data_as(fp), out.ctypes.data_as(fp), xin.size, target, decay)
    return out


def agc_pyloop(x: np.ndarray, target: float = AGC_TARGET, decay: float = AGC_DECAY) -> np.ndarray:
    """Literal Python-loop form of :func:`agc` (small inputs; pins the C helper)."""
    g = 1.0
    out = np.empty(x.size, dtype=np.float32)
    for k, v in enumerate(x):
        m = abs(v)
        if m > AGC_FLOOR:
            g += decay * (target / m - g)
        out[k] = v * g
    return out


# --------------------------------------------------------------------------
# K12: per-chunk statistics
# --------------------------------------------------------------------------

def rms_dbfs(audio: np.ndarray) -> float:
    """ref: decoders/nfm.py:87-88 (same expression in am.py:30-31, ssb.py:46-47)."""
    rms = math.sqrt(float(np.mean(audio.astype(np.float64) ** 2)) + 1e-18)
    return 20.0 * math.log10(rms + 1e-12)


# --------------------------------------------------------------------------
# decoder plug-ins (K6-K11 chained), one state record per channel
# --------------------------------------------------------------------------

NFM_MODES = {"nfm", "fm"}
SSB_MODES = {"usb", "ssb", "lsb"}


@dataclass
class DemodState:
    mode: str
    fs_channel: float
    agc_on: bool = True
    disc: DiscState = field(default_factory=DiscState)
    deemph: DeemphState | None = None
    dc: DcBlockState = field(default_factory=DcBlockState)

    @classmethod
    def create(cls, mode: str, fs_channel: float, *, deemph_us: float = 300.0,
               agc_enabled: bool = True) -> "DemodState":
        m = mode.lower()
        if m not in NFM_MODES | SSB_MODES | {"am"}:
            raise ValueError(f"Unsupported demod mode '{m}'.")   # ref: decoders/__init__.py:24
        st = cls(mode=m, fs_channel=fs_channel, agc_on=agc_enabled)
        if m in NFM_MODES:
            st.deemph = DeemphState.design(deemph_us, fs_channel)
        return st


def demodulate(st: DemodState, s: np.ndarray) -> tuple[np.ndarray, float]:
    """One decoder.process call: complex64 channel samples -> (float32 audio, rms dBFS).

    ref: decoders/nfm.py:82-97, decoders/am.py:25-41, decoders/ssb.py:38-61.
    NFM and AM ignore ``agc_on`` (ref: decoders/__init__.py:16-19).
    """
    if st.mode in NFM_MODES:
        audio = deemphasis(st.deemph, fm_discriminate(st.disc, s))
    elif st.mode == "am":
        audio = dc_block(st.dc, np.abs(s).astype(np.float32, copy=False))
    else:
        # ref: decoders/ssb.py:42-43 -- real(conj(s)) == real(s): USB and LSB identical
        base = (np.conj(s) if st.mode == "lsb" else s).real.astype(np.float32, copy=False)
        audio = dc_block(st.dc, base)
        if st.agc_on:
            audio = agc(audio)
    return audio, rms_dbfs(audio)


# --------------------------------------------------------------------------
# K13: writer-side peak and clip
# --------------------------------------------------------------------------

CLIP_LEVEL = 0.99       # ref: processing.py:452


def peak_and_clip(audio: np.ndarray, running_peak: float) -> tuple[np.ndarray, float]:
    """ref: processing.py:449-453 -- pre-clip running peak, clip to +-0.99, float32."""
    if audio.size == 0:
        return audio, running_peak
    pk = float(np.max(np.abs(audio)))
    return (np.clip(audio, -CLIP_LEVEL, CLIP_LEVEL).astype(np.float32, copy=False),
            max(running_peak, pk))


# --------------------------------------------------------------------------
# setup-time: K15 filter design, K16 mixer-sign probe, chunk/decimation planning
# --------------------------------------------------------------------------

def plan_decimation(sample_rate: float, fs_ch_target: float) -> tuple[int, float]:
    """ref: processing.py:885-890."""
    d = max(1, int(round(sample_rate / fs_ch_target)))
    fs_ch = sample_rate / d
    if fs_ch > fs_ch_target * 1.5:
        d = max(int(math.floor(sample_rate / fs_ch_target)), 1)
        fs_ch = sample_rate / d
    return d, fs_ch


def plan_chunk(sample_rate: float, requested: int) -> int:
    """ref: processing.py:65-81 (``tune_chunk_size``)."""
    base = max(1, requested)
    if sample_rate <= 0:
        return base
    secs = 0.25
    if sample_rate >= 2_000_000.0:
        secs = 0.40
    if sample_rate >= 5_000_000.0:
        secs = 0.50
    want = int(round(sample_rate * secs))
    if want <= base:
        return base
    cap = 4_194_304
    want = min(cap, max(base, want))
    return int(min(max(1 << math.ceil(math.log2(want)), base), cap))


def channel_taps(sample_rate: float, bandwidth: float, decimation: int) -> np.ndarray:
    """ref: processing.py:599-620 -- 80 dB Kaiser firwin, odd length in [1025, 32769]."""
    trans = max(1_000.0, bandwidth * 0.5)
    cutoff = min(bandwidth * 0.5 * 1.05, (sample_rate / (2.0 * max(decimation, 1))) * 0.9)
    if cutoff <= 0:
        raise ValueError("Invalid cutoff frequency for channel filter.")
    n = int(np.clip(4.0 / max(trans / sample_rate, 1e-8), 1024, 32768))
    n += (n % 2 == 0)
    return np.asarray(_ssig.firwin(n, cutoff=cutoff, window=("kaiser", _ssig.kaiser_beta(80.0)),
                                   fs=sample_rate), dtype=np.float64)


def pick_mix_sign(warmup: np.ndarray, sample_rate: float, freq_offset: float,
                  taps: np.ndarray, decimation: int) -> int:
    """ref: processing.py:623-663 -- try sign=+1 then -1, keep the strictly larger
    mean channel power after the filter's settling samples."""
    if warmup.size == 0:
        return 1
    limit = max(int(sample_rate * 0.05), len(taps) * 4, 131_072)
    take = min(warmup.size, limit)
    if take < len(taps):
        take = min(warmup.size, len(taps) * 2)
    snip = warmup[:take].astype(np.complex64, copy=False)
    idx = np.arange(snip.size, dtype=np.float64)
    d = max(decimation, 1)
    blk = min(snip.size, max(len(taps), 16_384))
    winner, best = 1, -np.inf
    for sgn in (1, -1):
        lo = np.exp(-1j * sgn * 2.0 * np.pi * freq_offset * idx / sample_rate).astype(np.complex64, copy=False)
        dec = fir_overlap_save(FirState(taps, blk), snip * lo)[::d]
        if dec.size == 0:
            pw = -np.inf
        else:
            drop = min(len(taps), dec.size // 4)
            keep = dec[drop:]
            if keep.size == 0:
                keep = dec
            pw = float(np.mean(np.abs(keep) ** 2))
        if pw > best:
            best, winner = pw, sgn
    return winner


# --------------------------------------------------------------------------
# the hot loop (ref: processing.py:1070-1154) for one target, in memory
# --------------------------------------------------------------------------

@dataclass
class TargetPlan:
    """Everything `ProcessingPipeline.run` derives before the loop (ref: processing.py:885-1002)."""
    sample_rate: float
    freq_offset: float
    bandwidth: float = 12_500.0
    mode: str = "nfm"
    fs_ch_target: float = 96_000.0
    deemph_us: float = 300.0
    agc_enabled: bool = True
    filter_block: int = 65_536
    mix_sign: int | None = None        # None -> probe on the first chunk

    decimation: int = 0
    fs_channel: float = 0.0
    taps: np.ndarray = field(default=None, repr=False)

    def __post_init__(self) -> None:
        self.decimation, self.fs_channel = plan_decimation(self.sample_rate, self.fs_ch_target)
        self.taps = channel_taps(self.sample_rate, self.bandwidth, self.decimation)


@dataclass
class StreamResult:
    audio: np.ndarray                  # float32, pre-clip (decoder output, concatenated)
    clipped: np.ndarray                # float32, what the writer pipes to the encoder
    baseband: np.ndarray               # complex64 decimated channel samples
    counts: list[int]                  # decimated samples per chunk
    rms_dbfs: list[float]              # per-chunk decoder statistic
    peak: float
    mix_sign: int
    final: dict                        # carried state after the last chunk


def run_target(x: np.ndarray, plan: TargetPlan, chunk: int,
               max_input_samples: int | None = None) -> StreamResult:
    """Replay of the reference loop body on an in-memory complex64 capture.

    ref: processing.py:1028-1042 (warm-up chunk + sign probe), :1070-1154 (loop,
    including the ``max_input_samples`` truncation at :1072-1079 and :1152-1154).
    """
    nco = NcoState.for_offset(plan.freq_offset, plan.sample_rate)
    fir = FirState(plan.taps, plan.filter_block)
    dec = DecimState(plan.decimation)
    dem = DemodState.create(plan.mode, plan.fs_channel, deemph_us=plan.deemph_us,
                            agc_enabled=plan.agc_enabled)
    warm = x[:chunk]
    if max_input_samples is not None and warm.size > max_input_samples:
        warm = warm[:max_input_samples]
    sign = plan.mix_sign if plan.mix_sign in (1, -1) else pick_mix_sign(
        warm, plan.sample_rate, plan.freq_offset, plan.taps, plan.decimation)

    audio_parts, clip_parts, bb_parts, counts, stats = [], [], [], [], []
    peak = 0.0
    done = 0
    for start in range(0, x.size, chunk):
        blk = x[start:start + chunk]
        if max_input_samples is not None:
            left = max_input_samples - done
            if left <= 0:
                break
            blk = blk[:left]
        if blk.size == 0:
            continue
        done += blk.size
        chan = decimate(dec, fir_overlap_save(fir, nco_mix(nco, blk, sign)))
        audio, db = demodulate(dem, chan)
        safe, peak = peak_and_clip(audio, peak)
        bb_parts.append(chan)
        audio_parts.append(audio)
        clip_parts.append(safe)
        counts.append(int(chan.size))
        stats.append(db)
        if max_input_samples is not None and done >= max_input_samples:
            break

    def _cat(parts, dt):
        return np.concatenate(parts) if parts else np.empty(0, dtype=dt)

    final = {
        "phase": nco.phase, "offset": dec.offset,
        "disc_prev": complex(dem.disc.prev),
        "deemph_z": dem.deemph.z if dem.deemph else 0.0,
        "dc_x": dem.dc.x_prev, "dc_y": dem.dc.y_prev,
    }
    return StreamResult(_cat(audio_parts, np.float32), _cat(clip_parts, np.float32),
                        _cat(bb_parts, np.complex64), counts, stats, peak, sign, final)


# --------------------------------------------------------------------------
# synthetic captures -- deterministic, shared by tests and bench
# --------------------------------------------------------------------------

def benchmark_capture_s16(sample_rate: float, seconds: float, freq_offset: float,
                          amplitude: float = 0.7, noise_std: float = 0.02) -> np.ndarray:
    """The ``--benchmark`` input: tone + AWGN (seed 42), clipped, PCM_16.

    ref: benchmark.py:19-38.  libsndfile's float -> PCM_16 is taken as
    lrint(x * 32767)-style scaling... the exact rule is not verifiable here
     and does not matter for parity: both arms read the same int16.
    Returns interleaved int16 [2*N].
    """
    n = int(round(sample_rate * seconds))
    if n <= 0:
        raise ValueError("Benchmark duration is too short to generate samples.")
    t = np.arange(n, dtype=np.float64) / sample_rate
    tone = np.exp(1j * 2.0 * math.pi * freq_offset * t)
    noise = np.random.default_rng(42).normal(scale=noise_std, size=(n, 2))
    iq = np.clip(np.column_stack((amplitude * tone.real + noise[:, 0],
                                  amplitude * tone.imag + noise[:, 1])).astype(np.float32),
                 -0.999, 0.999)
    return np.round(iq * 32767.0).astype(np.int16).ravel()


def multi_carrier_capture(sample_rate: float, n: int, carriers: list[dict], *,
                          noise_std: float = 0.02, seed: int = 42,
                          n0: int = 0) -> np.ndarray:
    """Sum of modulated carriers + AWGN as float64 I/Q columns [n, 2].

    Each carrier dict: ``offset`` (Hz), ``amp``, ``kind`` in {"fm","am","usb","lsb","cw"},
    ``tone`` (Hz), and ``dev`` (FM deviation, Hz) or ``depth`` (AM).
    ``n0`` offsets the time axis (segment generation).  Noise is drawn for the
    whole [0, n0+n) range only when n0 == 0; segment generators use their own seed.
    """
    t = (np.arange(n, dtype=np.float64) + n0) / sample_rate
    acc = np.zeros(n, dtype=np.complex128)
    for c in carriers:
        w = 2.0 * math.pi * c["offset"] * t
        kind = c.get("kind", "fm")
        tone = c.get("tone", 1000.0)
        if kind == "fm":
            beta = c.get("dev", 2500.0) / tone
            acc += c["amp"] * np.exp(1j * (w + beta * np.sin(2.0 * math.pi * tone * t)))
        elif kind == "am":
            env = 1.0 + c.get("depth", 0.8) * np.sin(2.0 * math.pi * tone * t)
            acc += c["amp"] * env * np.exp(1j * w)
        elif kind in ("usb", "lsb"):
            sgn = 1.0 if kind == "usb" else -1.0
            # suppressed-carrier tone plus a pilot so Re(s) is never identically ~0
            acc += c["amp"] * (np.exp(1j * (w + sgn * 2.0 * math.pi * tone * t)) + 0.5 * np.exp(1j * w))
        else:
            acc += c["amp"] * np.exp(1j * w)
    noise = np.random.default_rng(seed).normal(scale=noise_std, size=(n, 2))
    return np.column_stack((acc.real + noise[:, 0], acc.imag + noise[:, 1]))


def to_s16(iq_cols: np.ndarray) -> np.ndarray:
    """float I/Q columns -> interleaved int16 (clip +-0.999, round(x*32767))."""
    return np.round(np.clip(iq_cols, -0.999, 0.999) * 32767.0).astype(np.int16).ravel()


def to_u8(iq_cols: np.ndarray) -> np.ndarray:
    return np.clip(np.round(iq_cols * 127.0 + 128.0), 0, 255).astype(np.uint8).ravel()


def to_f32(iq_cols: np.ndarray) -> np.ndarray:
    return iq_cols.astype(np.float32).ravel()
