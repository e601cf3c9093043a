"""CPU-only: the C-ABI library builds/loads and exports every symbol include/iq2a_b200.h
declares; compute entry points fail loudly without a GPU (no CPU fallback)."""
from __future__ import annotations

import re
from pathlib import Path

import numpy as np
import pytest

from iq_to_audio_b200 import _lib

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols() -> set[str]:
    text = (ROOT / "include" / "iq2a_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(iq2a_[a-z0-9_]+)\s*\(", text))


def test_header_and_binding_agree():
    assert declared_symbols() == set(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.iq2a_version() >= 100


def test_structs_match_header_layout():
    import ctypes as C
    assert C.sizeof(_lib.ChannelState) == 32
    assert C.sizeof(_lib.BankConfig) == 48
    assert C.sizeof(_lib.ChannelDesc) == 40
    assert C.sizeof(_lib.BankInfo) == 48


def test_no_cpu_fallback_without_gpu():
    if _lib.device_count() > 0:
        pytest.skip("a GPU is present")
    from iq_to_audio_b200.bank import ChannelBank, Target
    from iq_to_audio_b200.processing import ComplexOscillator, Decimator, OverlapSaveFIR
    with pytest.raises(RuntimeError, match="no CUDA device"):
        ComplexOscillator(1e3, 1e6).mix(np.ones(8, dtype=np.complex64), 1)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        OverlapSaveFIR(np.ones(5), 16).process(np.ones(8, dtype=np.complex64))
    with pytest.raises(RuntimeError, match="no CUDA device"):
        Decimator(3).process(np.ones(8, dtype=np.complex64))
    with pytest.raises(RuntimeError, match="no CUDA device"):
        ChannelBank(1e6, 10, [Target(1e3, np.ones(33))])


def test_argument_validation_mirrors_reference_errors():
    from iq_to_audio_b200.bank import ChannelBank, Target
    from iq_to_audio_b200.decoders import create_decoder
    from iq_to_audio_b200.processing import OverlapSaveFIR
    with pytest.raises(ValueError, match="block_size must be positive"):       # ref processing.py:304-305
        OverlapSaveFIR(np.ones(5), 0)
    with pytest.raises(ValueError, match="Unsupported demod mode"):             # ref decoders/__init__.py:24
        create_decoder("cw", deemph_us=300.0, agc_enabled=True)
    with pytest.raises(ValueError, match="Unsupported iq_order"):               # ref processing.py:269-270
        ChannelBank(1e6, 10, [Target(1e3, np.ones(33))], iq_order="xy")
    with pytest.raises(RuntimeError, match="setup"):                            # ref decoders/nfm.py:83-84
        create_decoder("nfm", deemph_us=300.0, agc_enabled=True).process(np.ones(4, dtype=np.complex64))
    # empty inputs return empty outputs without touching the device (ref processing.py:290-291, :326-327)
    from iq_to_audio_b200.processing import ComplexOscillator, Decimator
    e = np.empty(0, dtype=np.complex64)
    assert ComplexOscillator(1e3, 1e6).mix(e, 1).size == 0
    assert OverlapSaveFIR(np.ones(5), 16).process(e).size == 0
    assert Decimator(4).process(e).size == 0
