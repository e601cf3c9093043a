"""Host-side planning for the polyphase overlap-save channel bank.

The reference computes, per target, ``decimate(fir(mix(x)))`` at the full input
rate (``src/iq_to_audio/processing.py:1088-1096``).  The B200 path computes the
same decimated samples directly:

    s_c[m] = e^{j phi_c(mD)} * sum_k h_c[k] e^{-j w_c k} x[mD - k]            (1)

(mix-then-filter == filter-with-modulated-taps-then-rotate, exact because the
reference's phase ramp is linear inside a chunk, processing.py:292-293).  With
k = qD - p, p in [0, D), the inner sum is a sum over the D polyphase branches of
the *raw* input, x_p[m] = x[mD + p], each convolved at the channel rate with
g_{c,p}[q] = h_c[qD - p] e^{-j w_c (qD - p)}:

    s_c[m] = e^{j phi_c(mD)} * sum_p (g_{c,p} * x_p)[m]                        (2)

Each branch convolution is done by overlap-save with M-point transforms:
per block, D forward M-point FFTs of the raw branches (shared by every
channel), a multiply-accumulate over p against the precomputed spectra
G[c, p, :] = FFT_M(g_{c,p}) / M, and ONE inverse M-point FFT per channel.  This
is the decimation-in-time factorisation of the reference's N = D*M point
overlap-save with the last radix-D stage, the tap multiply and the spectral
fold for "keep every D-th sample" collapsed into the G table -- so the
decimated output is exact (all aliases summed), and no transform is ever longer
than M <= 2048 points, which fits a CTA's shared memory.

This module only builds tables (numpy, float64 -> float32); all per-sample work
is in ``csrc/``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

MODE_IDS = {"nfm": 0, "fm": 0, "am": 1, "usb": 2, "ssb": 2, "lsb": 3, "iq": 4, "none": 4, "pass": 4}
SUPPORTED_M = (256, 512, 1024, 2048)
#: first-pass radix of the in-register transforms for each M (csrc/fft_regs.cuh)
RADIX1 = {256: 16, 512: 32, 1024: 32, 2048: 32}


def spectrum_slot_to_bin(m_fft: int) -> np.ndarray:
    """Storage slot j' -> DFT bin k for the kernel's two-pass transform.

    Pass 1 is an R1-point DIF step over m1 (m = M2*m1 + m2), pass 2 an M2-point
    step over m2; the result for bin k = k1 + R1*k2 is left in slot
    j' = k1*M2 + k2 (csrc/channelizer.cu).  The inverse transform consumes the
    same layout, so no bit-reversal pass exists anywhere.
    """
    r1 = RADIX1[m_fft]
    m2 = m_fft // r1
    slots = np.arange(m_fft)
    k1, k2 = slots // m2, slots % m2
    return k1 + r1 * k2


def overlap_rows(ntaps_max: int, decimation: int) -> int:
    """Rows (channel-rate samples) of history a block needs: the longest branch
    filter has q in [0, Qmax]; one more row yields s[m-1] for the discriminator."""
    qmax = (ntaps_max - 1 + decimation - 1) // decimation
    return qmax + 1


def choose_fft_size(ntaps_max: int, decimation: int, n_channels: int, filter_block: int | None = None) -> int:
    """Pick M.  Cost model per new input sample: (5*log2(M) + 8*C) / (1 - Vd/M)
    flops (forward branch transforms + per-channel multiply-accumulate, divided
    by the overlap-save efficiency).  ``filter_block`` (the reference's hop,
    processing.py:309-310) is honoured as an upper bound on the hop D*(M - Vd)
    when it is small enough to matter."""
    vd = overlap_rows(ntaps_max, decimation)
    best, best_cost = None, math.inf
    for m in SUPPORTED_M:
        if m <= vd + 8:
            continue
        cost = (5.0 * math.log2(m) + 8.0 * n_channels) / (1.0 - vd / m)
        if cost < best_cost:
            best, best_cost = m, cost
    if best is None:
        raise ValueError(f"channel filter too long for the supported transform sizes (needs {vd} history rows)")
    return best


@dataclass
class ChannelSpec:
    """One target (what the reference derives per ProcessingConfig, processing.py:882-1002)."""
    freq_offset: float
    taps: np.ndarray                  # float64 [ntaps], real
    mix_sign: int = 1
    mode: str = "nfm"
    deemph_us: float = 300.0
    agc_enabled: bool = True


@dataclass
class BankPlan:
    sample_rate: float
    decimation: int
    channels: list[ChannelSpec]
    m_fft: int
    vd: int = 0                       # overlap rows
    ld: int = 0                       # new channel-rate samples per block
    increments: np.ndarray = field(default=None, repr=False)   # float64 [C]: sign * (-2 pi f_off / fs)
    g_table: np.ndarray = field(default=None, repr=False)      # complex64 [C, D, M] in slot order

    @property
    def n_channels(self) -> int:
        return len(self.channels)

    @property
    def hop(self) -> int:
        return self.ld * self.decimation

    @property
    def fs_channel(self) -> float:
        return self.sample_rate / self.decimation


def branch_filters(taps: np.ndarray, w: float, decimation: int, m_fft: int) -> np.ndarray:
    """g[p, q] = h[qD - p] * exp(-j*w*(qD - p)), zero elsewhere; complex128 [D, M]."""
    d = decimation
    g = np.zeros((d, m_fft), dtype=np.complex128)
    k = np.arange(len(taps))
    q = (k + d - 1) // d
    p = q * d - k
    if q.max() >= m_fft:
        raise ValueError("filter longer than the transform")
    g[p, q] = taps * np.exp(-1j * w * k)
    return g


def build_plan(sample_rate: float, decimation: int, channels: list[ChannelSpec], *,
               m_fft: int | None = None, filter_block: int | None = None) -> BankPlan:
    if not channels:
        raise ValueError("at least one channel is required")
    d = max(1, int(decimation))
    nt_max = max(len(c.taps) for c in channels)
    if m_fft is None:
        m_fft = choose_fft_size(nt_max, d, len(channels), filter_block)
    if m_fft not in SUPPORTED_M:
        raise ValueError(f"unsupported transform size {m_fft}")
    vd = overlap_rows(nt_max, d)
    if vd + 1 > m_fft:
        raise ValueError("filter longer than the transform")
    plan = BankPlan(sample_rate=sample_rate, decimation=d, channels=channels, m_fft=m_fft, vd=vd, ld=m_fft - vd)
    perm = spectrum_slot_to_bin(m_fft)
    incs = np.empty(len(channels), dtype=np.float64)
    tab = np.empty((len(channels), d, m_fft), dtype=np.complex64)
    for ci, ch in enumerate(channels):
        # ref: processing.py:287 (increment), :293 (sign * increment * n)
        inc = -2.0 * np.pi * ch.freq_offset / sample_rate
        w = ch.mix_sign * inc
        incs[ci] = w
        # e^{+j w n} applied to x[n-k] -> taps carry e^{-j w k} relative to the output sample
        g = branch_filters(np.asarray(ch.taps, dtype=np.float64), w, d, m_fft)
        spec = np.fft.fft(g, axis=1) / m_fft
        tab[ci] = spec[:, perm].astype(np.complex64)
    plan.increments = incs
    plan.g_table = tab
    return plan


def phase_table(increment_signed: float, chunk: int, n_chunks: int, phase0: float = 0.0) -> np.ndarray:
    """Per-chunk NCO start phases exactly as the reference carries them.

    ref: processing.py:295 -- ``phase = (phase + sign*inc*size) % (2*pi)`` once per
    chunk, in Python floats.  ``increment_signed`` is ``sign * increment`` (the
    product the reference forms first, processing.py:293).
    """
    out = np.empty(max(1, n_chunks), dtype=np.float64)
    ph = float(phase0)
    two_pi = 2.0 * np.pi
    for k in range(out.size):
        out[k] = ph
        ph = (ph + increment_signed * chunk) % two_pi
    return out


def emulate_block_math(plan: BankPlan, x: np.ndarray, n_out: int, phases: np.ndarray, chunk: int) -> np.ndarray:
    """numpy model of what csrc/channelizer.cu computes (float64), for tests:
    complex128 [C, n_out] channel samples s_c[m], m = 0..n_out-1 (stream start at 0)."""
    d, m_fft, vd, ld = plan.decimation, plan.m_fft, plan.vd, plan.ld
    perm = spectrum_slot_to_bin(m_fft)
    inv_perm = np.argsort(perm)
    out = np.zeros((plan.n_channels, n_out), dtype=np.complex128)
    xs = np.asarray(x, dtype=np.complex128)
    g_nat = plan.g_table.astype(np.complex128)[:, :, inv_perm]        # back to natural bin order
    nblocks = (n_out + ld - 1) // ld
    for b in range(nblocks):
        row0 = b * ld - vd
        rows = np.zeros((m_fft, d), dtype=np.complex128)
        lo, hi = row0 * d, (row0 + m_fft) * d
        src_lo, src_hi = max(lo, 0), min(hi, xs.size)
        if src_hi > src_lo:
            rows.reshape(-1)[src_lo - lo:src_hi - lo] = xs[src_lo:src_hi]
        spec = np.fft.fft(rows, axis=0)                                 # [M, D]: X_p[j]
        for ci in range(plan.n_channels):
            y = np.fft.ifft((g_nat[ci].T * spec).sum(axis=1)) * m_fft   # G already carries 1/M
            m_glob = b * ld + np.arange(ld)
            keep = m_glob < n_out
            n_glob = m_glob[keep] * d
            seg = n_glob // chunk
            ph = phases[ci][seg] + plan.increments[ci] * (n_glob - seg * chunk)
            out[ci, m_glob[keep]] = y[vd:][keep] * np.exp(1j * ph)
    return out


# ---- mirror-pair form (csrc/channelizer5.cuh) -------------------------------------------------------------

@dataclass
class PairTile:
    """One tile of the mirror-pair kernel: forward column group ``u`` (columns 4u..4u+3) and mirror group ``w``
    (columns 4w..4w+3, block window rotated circularly by ``rot`` rows).  Column 4u+i mirrors column 4w+4-i
    (i = 1, 2, 3); column 4u mirrors column 4(w+1), which the PREVIOUS tile of the class carried over."""
    u: int
    w: int
    rot: int
    cls: int
    first: bool                       # first tile of its class: nothing carried
    last: bool                        # last tile of its class: column 4w is left over (its own mirror)


def pair_tiles(ntaps: int, decimation: int) -> tuple[list[PairTile], int, int]:
    """Tiles of the mirror-pair kernel for symmetric taps h[n] = h[ntaps-1-n] (firwin's, ref processing.py:613).

    With A = ceil((ntaps-1)/D) and r = A*D - (ntaps-1): branch p <= r mirrors r - p (window rotation A, class 0),
    branch p > r mirrors r + D - p (rotation A + 1, class 1).  The tensor copy moves aligned groups of four
    columns, hence D % 4 == 0 and (ntaps - 1) % 4 == 0 (then r % 4 == 0).  Returns (tiles, A, r)."""
    d = decimation
    if d % 4 or (ntaps - 1) % 4:
        raise ValueError("the pair kernel needs D % 4 == 0 and (ntaps - 1) % 4 == 0")
    a = -(-(ntaps - 1) // d)
    r = a * d - (ntaps - 1)
    g1, g2 = r // 4, d // 4 - r // 4
    tiles: list[PairTile] = []
    for cls, (gl, g) in enumerate(((0, g1), (g1, g2))):
        n = (g + 1) // 2
        for j in range(n):
            tiles.append(PairTile(u=gl + j, w=gl + g - 1 - j, rot=a + cls, cls=cls, first=j == 0, last=j == n - 1))
    return tiles, a, r


def emulate_block_math_paired(plan: BankPlan, x: np.ndarray, n_out: int, phases: np.ndarray, chunk: int) -> np.ndarray:
    """numpy model of csrc/channelizer5.cuh (float64): same result as :func:`emulate_block_math`, computed from
    one REAL-pair table entry (a, b) per branch pair:  Y = kappa^(1/2) * sum_pairs a*(X_p + X~_p') + j*b*(X_p - X~_p')
    with  a + j*b = kappa^(-1/2) * G[c, p],  kappa = exp(-j*w*(L-1)),  X~ the transform of the rotated window;
    the tile / carry / end-of-class schedule is the kernel's."""
    d, m_fft, vd, ld = plan.decimation, plan.m_fft, plan.vd, plan.ld
    ntaps = len(plan.channels[0].taps)
    for ch in plan.channels:
        h = np.asarray(ch.taps)
        if len(h) != ntaps or not np.array_equal(h, h[::-1]):
            raise ValueError("the pair form needs symmetric taps of one length")
    tiles, _, _ = pair_tiles(ntaps, d)
    out = np.zeros((plan.n_channels, n_out), dtype=np.complex128)
    xs = np.asarray(x, dtype=np.complex128)
    tabs, kap_half = [], []
    k = np.arange(m_fft)
    for ci, ch in enumerate(plan.channels):
        w = plan.increments[ci]
        g = branch_filters(np.asarray(ch.taps, dtype=np.float64), w, d, m_fft)
        tabs.append(np.fft.fft(g, axis=1) / m_fft * np.exp(1j * w * (ntaps - 1) / 2.0))
        kap_half.append(np.exp(-1j * w * (ntaps - 1) / 2.0))
    nblocks = (n_out + ld - 1) // ld
    for b in range(nblocks):
        row0 = b * ld - vd
        rows = np.zeros((m_fft, d), dtype=np.complex128)
        lo, hi = row0 * d, (row0 + m_fft) * d
        src_lo, src_hi = max(lo, 0), min(hi, xs.size)
        if src_hi > src_lo:
            rows.reshape(-1)[src_lo - lo:src_hi - lo] = xs[src_lo:src_hi]
        y_spec = np.zeros((plan.n_channels, m_fft), dtype=np.complex128)

        def pair(entry_col, weight, xf, xm):
            s, dd = xf + xm, xf - xm
            for ci in range(plan.n_channels):
                e = tabs[ci][entry_col] * weight
                y_spec[ci] += e.real * s + 1j * e.imag * dd

        carried = None
        for t in tiles:
            xf = [np.fft.fft(rows[:, 4 * t.u + i]) for i in range(4)]
            xm = [np.fft.fft(np.roll(rows[:, 4 * t.w + i], t.rot)) for i in range(4)]
            for i in (1, 2, 3):
                f, m = 4 * t.u + i, 4 * t.w + 4 - i
                pair(f, 1.0 if f < m else (0.5 if f == m else 0.0), xf[i], xm[4 - i])
            pair(4 * t.u, 1.0, xf[0], np.zeros(m_fft) if t.first else carried)
            if t.last:
                if t.w != t.u:                                        # an even number of groups: column 4w is left
                    rot_back = np.exp(2j * np.pi * t.rot * k / m_fft)
                    for ci in range(plan.n_channels):
                        y_spec[ci] += tabs[ci][4 * t.w] * rot_back * xm[0]
                carried = None
            else:
                carried = xm[0]
        for ci in range(plan.n_channels):
            y = np.fft.ifft(y_spec[ci]) * m_fft * kap_half[ci]
            m_glob = b * ld + np.arange(ld)
            keep = m_glob < n_out
            n_glob = m_glob[keep] * d
            seg = n_glob // chunk
            ph = phases[ci][seg] + plan.increments[ci] * (n_glob - seg * chunk)
            out[ci, m_glob[keep]] = y[vd:][keep] * np.exp(1j * ph)
    return out
