"""Oracle against the LIVE reference on seeded random configurations (build container only).

`tests/test_oracle_golden.py` pins the oracle on the recorded cases; this file widens that to configurations nobody
chose by hand -- sample rates with odd decimations, ragged chunk sizes, filter blocks shorter than the filter, every
demodulator, both mixer signs, preview truncation -- by importing the unmodified reference from `/root/reference/src`
(the way `tests/golden/make_golden.py` does) and comparing bit for bit.  `/root/reference` does not exist on the GPU
box: everything here is skipped there.  CPU only.
"""
from __future__ import annotations

import importlib.util
from pathlib import Path

import numpy as np
import pytest

from oracle import iq_oracle as orc

REF_SRC = Path("/root/reference/src")
pytestmark = pytest.mark.skipif(not REF_SRC.exists(), reason="reference tree only exists in the build container")


@pytest.fixture(scope="module")
def ref():
    spec = importlib.util.spec_from_file_location("make_golden", Path(__file__).parent / "golden" / "make_golden.py")
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    proc, dec, *_ = mg._import_reference()
    return mg, proc, dec


def _capture(rng, fs, n, f_off, mode):
    kind = {"nfm": "fm", "am": "am", "usb": "usb", "lsb": "lsb"}[mode]
    carrier = dict(offset=f_off, amp=float(rng.uniform(0.1, 0.4)), kind=kind, tone=float(rng.uniform(400.0, 2500.0)))
    if kind == "fm":
        carrier["dev"] = 2500.0
    if kind == "am":
        carrier["depth"] = 0.7
    other = dict(offset=-f_off * 0.6 + 0.05 * fs, amp=0.15, kind="fm", tone=900.0, dev=2500.0)
    return orc.multi_carrier_capture(fs, n, [carrier, other], seed=int(rng.integers(1 << 30)))


# (seed, fs, mode, bandwidth): decimations 13, 25, 52, 10, 31 (non powers of two on purpose)
CONFIGS = [
    (1, 1.25e6, "nfm", 12_500.0),
    (2, 2.4e6, "am", 10_000.0),
    (3, 5.0e6, "usb", 2_800.0),
    (4, 0.96e6, "lsb", 2_800.0),
    (5, 3.0e6, "nfm", 25_000.0),
]


@pytest.mark.parametrize("seed,fs,mode,bw", CONFIGS)
def test_oracle_equals_live_reference(ref, seed, fs, mode, bw):
    mg, proc, dec = ref
    rng = np.random.default_rng(seed)
    n = int(rng.integers(150_000, 260_000))
    f_off = float(rng.uniform(0.08, 0.38) * fs * rng.choice([-1.0, 1.0]))
    chunk = int(rng.integers(30_000, 90_000))                    # ragged: not a power of two, not a multiple of D
    filter_block = int(rng.choice([1_024, 4_096, 16_384]))       # 1 024 < ntaps - 1 for every case: the state path
    agc = bool(rng.integers(2))
    sign = [None, 1, -1][int(rng.integers(3))]
    limit = None if seed % 2 else int(n * 0.7) + 1               # preview truncation on the even seeds
    codec, packer = [("pcm_s16le", orc.to_s16), ("pcm_u8", orc.to_u8), ("pcm_f32le", orc.to_f32)][seed % 3]
    order = ["iq", "qi", "iq_inv", "qi_inv"][seed % 4]
    x = orc.order_iq(orc.unpack_interleaved(packer(_capture(rng, fs, n, f_off, mode)), codec), order)

    g = mg.run_reference_loop(proc, dec, x, fs=fs, f_off=f_off, bw=bw, mode=mode, chunk=chunk,
                              filter_block=filter_block, agc=agc, mix_sign=sign, max_input_samples=limit)
    plan = orc.TargetPlan(sample_rate=fs, freq_offset=f_off, bandwidth=bw, mode=mode, agc_enabled=agc,
                          filter_block=filter_block, mix_sign=sign)
    assert plan.decimation == int(g["decimation"]) and len(plan.taps) == int(g["ntaps"])
    res = orc.run_target(x, plan, chunk, max_input_samples=limit)

    assert res.counts == list(g["counts"])                       # chunk boundaries and decimation phase: exact
    assert res.mix_sign == int(g["mix_sign"])
    np.testing.assert_array_equal(res.baseband, g["baseband"])
    np.testing.assert_array_equal(res.audio, g["audio"])
    np.testing.assert_array_equal(res.clipped, g["clipped"])
    assert res.peak == float(g["peak"])
    np.testing.assert_array_equal(np.asarray(res.rms_dbfs), g["rms_dbfs"])
    assert res.final["phase"] == float(g["final_phase"]) and res.final["offset"] == int(g["final_offset"])


def test_product_filter_design_is_bit_identical_to_the_live_reference(ref):
    """`iq_to_audio_b200.processing.design_channel_filter` against the reference's own function over a grid of
    rates, decimations and bandwidths, including the (fs, D) pairs for which (fs/(2D))*0.9 and 0.9*fs/(2D) differ by
    one ulp and the Nyquist-bound cutoff is the one selected (wide bandwidths) -- ADVICE round 1."""
    from iq_to_audio_b200 import processing as gp
    _, proc, _ = ref
    n_cases = n_bound = 0
    for fs in (250e3, 1.0e6, 2.4e6, 2.5e6, 5.0e6, 10e6, 20e6, 61.44e6):
        for d in (1, 2, 3, 13, 26, 104, 208, 640):
            if fs / d < 8_000.0:
                continue
            for bw in (2_800.0, 10_000.0, 12_500.0, 25_000.0, 200_000.0):
                if (fs / (2.0 * d)) * 0.9 < bw * 0.5 * 1.05:
                    n_bound += 1
                want = proc.design_channel_filter(fs, bw, d)
                np.testing.assert_array_equal(gp.design_channel_filter(fs, bw, d), want)
                np.testing.assert_array_equal(orc.channel_taps(fs, bw, d), want)
                n_cases += 1
    assert n_cases > 100 and n_bound > 20
