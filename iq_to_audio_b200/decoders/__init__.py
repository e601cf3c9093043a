"""Decoder factory (ref: src/iq_to_audio/decoders/__init__.py:9-24)."""
from __future__ import annotations

from .am import AMDecoder
from .base import Decoder, DecoderStats
from .nfm import NarrowbandFMDecoder
from .ssb import SSBDecoder


def create_decoder(mode: str, *, deemph_us: float, agc_enabled: bool) -> Decoder:
    key = mode.lower()
    if key in ("nfm", "fm"):
        return NarrowbandFMDecoder(deemph_us=deemph_us)
    if key == "am":
        return AMDecoder()
    if key in ("usb", "ssb", "lsb"):
        return SSBDecoder(sideband="lsb" if key == "lsb" else "usb", agc_enabled=agc_enabled)
    raise ValueError(f"Unsupported demod mode '{key}'.")


__all__ = ["Decoder", "DecoderStats", "create_decoder", "NarrowbandFMDecoder", "AMDecoder", "SSBDecoder"]
