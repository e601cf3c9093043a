"""CPU-only: the numpy model of libswresample's default resampler (oracle/swr_model.py) against
vectors produced by the real library (tests/golden/make_resampler_golden.py)."""
from __future__ import annotations

import numpy as np
import pytest

from oracle.swr_model import SwrModel
from tests import _cases


@pytest.fixture(scope="module")
def rv():
    return _cases.load("resampler_vectors")


@pytest.mark.parametrize("key", ["r96154_n40330", "r96154_n100991", "r96000_n26219", "r95238_n20000",
                                 "r100000_n12345", "r96154_n333"])
def test_model_matches_libswresample(rv, key):
    in_rate = int(key.split("_")[0][1:])
    x = rv[key + "_in"]
    m = SwrModel(in_rate)
    y = m.resample(x)
    assert y.size == rv[key + "_flt"].size == int(rv[key + "_counts"].sum())
    assert np.abs(y - rv[key + "_flt"]).max() <= 2e-7
    s = m.resample_s16(x)
    d = np.abs(s.astype(np.int64) - rv[key + "_s16"].astype(np.int64))
    assert d.max() <= 1                                   # +-1 LSB (north star); in fact almost always 0
    assert np.mean(d == 0) > 0.999


def test_output_count_rule(rv):
    for in_rate, n, n_main, n_total in rv["count_table"]:
        assert SwrModel(int(in_rate)).output_count(int(n)) == (int(n_total), int(n_main))


def test_design_parameters():
    m = SwrModel(96_154)
    assert (m.filter_length, m.phase_count, m.center, m.src_incr, m.dst_incr) == (68, 1024, 33, 750, 1_538_464)
    m = SwrModel(96_000)                                   # cfg4: exact 2:1, a single phase
    assert (m.filter_length, m.phase_count) == (66, 1)
    assert SwrModel(48_000).passthrough
