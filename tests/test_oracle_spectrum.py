"""The spectrum oracle against outputs of the reference's own spectrum.py (tests/golden/spectrum_vectors.npz)."""
from pathlib import Path

import numpy as np
import pytest

from oracle import spectrum_oracle as so
from tests._spectrum_cases import PSD_CASES, WATERFALL_CASES, psd_input, waterfall_chunks

GOLD = np.load(Path(__file__).parent / "golden" / "spectrum_vectors.npz")
# float64 FFTs of the same data through two pocketfft builds (scipy.fft in the reference, numpy.fft here):
# the difference is rounding noise relative to the strongest bin.  In dB it is largest on the weakest bins.
DB_TOL = 1e-6


@pytest.mark.parametrize("name", list(PSD_CASES))
def test_compute_psd(name):
    case = PSD_CASES[name]
    got = so.psd_one(psd_input(case), case["fs"], case["nfft"])
    np.testing.assert_allclose(got, GOLD[f"psd_{name}_db"], rtol=0, atol=DB_TOL)
    f = so.freq_axis(case["nfft"], case["fs"])
    np.testing.assert_array_equal(np.concatenate([f[:4], f[-4:]]), GOLD[f"psd_{name}_freqs_edge"])


@pytest.mark.parametrize("name", list(WATERFALL_CASES))
def test_window_starts(name):
    case = WATERFALL_CASES[name]
    hop = max(1, case["hop"] or case["nfft"] // 4)
    sizes = [0 if c is None else c.size for c in waterfall_chunks(case)]
    got = so.window_starts(sizes, case["nfft"], hop)
    np.testing.assert_array_equal(np.asarray([s for s, _ in got], dtype=np.int64), GOLD[f"wf_{name}_starts"])


@pytest.mark.parametrize("name", list(WATERFALL_CASES))
def test_streaming_waterfall(name):
    case = WATERFALL_CASES[name]
    avg, times, matrix, frames = so.waterfall(waterfall_chunks(case), case["fs"], case["nfft"], case["hop"],
                                              case["max_slices"])
    assert frames == int(GOLD[f"wf_{name}_frames"])
    np.testing.assert_array_equal(times, GOLD[f"wf_{name}_times"])
    assert matrix.shape == GOLD[f"wf_{name}_matrix"].shape and matrix.dtype == np.float32
    np.testing.assert_allclose(avg, GOLD[f"wf_{name}_avg"], rtol=0, atol=DB_TOL)
    np.testing.assert_allclose(matrix, GOLD[f"wf_{name}_matrix"], rtol=0, atol=2e-5)   # float32 rows near -100 dB


def test_errors():
    with pytest.raises(ValueError, match="empty"):
        so.psd_one(np.empty(0, np.complex64), 1e6, 64)
    with pytest.raises(ValueError, match="enough samples"):
        so.waterfall([np.zeros(100, np.complex64)], 1e6, 256, None, 10)


def test_start_index_slips_after_a_yield():
    # documents the reference behaviour the device code reproduces: the second chunk's windows are reported
    # `pending` samples early (spectrum.py:108-110 subtracts the pending length again after :125 already did)
    got = so.window_starts([1000, 1000], 512, 512)
    assert got == [(0, 0), (24, 512), (536, 1024)]      # 488 samples were pending after the first chunk
