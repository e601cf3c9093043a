"""Decoder plug-ins on the GPU.  `create_decoder` keeps the reference factory's contract
(src/iq_to_audio/decoders/__init__.py:9-24): case-insensitive mode names nfm/fm, am, usb/ssb, lsb;
`ValueError` for anything else; NFM and AM ignore `agc_enabled`."""
from __future__ import annotations

from .am import AMDecoder
from .base import Decoder, DecoderStats
from .nfm import NarrowbandFMDecoder
from .ssb import SSBDecoder

_BUILDERS = {
    "nfm": lambda deemph_us, agc: NarrowbandFMDecoder(deemph_us=deemph_us),
    "am": lambda deemph_us, agc: AMDecoder(),
    "usb": lambda deemph_us, agc: SSBDecoder(sideband="usb", agc_enabled=agc),
    "lsb": lambda deemph_us, agc: SSBDecoder(sideband="lsb", agc_enabled=agc),
}
_ALIASES = {"fm": "nfm", "ssb": "usb"}


def create_decoder(mode: str, *, deemph_us: float, agc_enabled: bool) -> Decoder:
    key = mode.lower()
    build = _BUILDERS.get(_ALIASES.get(key, key))
    if build is None:
        raise ValueError(f"Unsupported demod mode '{key}'.")
    return build(deemph_us, agc_enabled)


__all__ = ["Decoder", "DecoderStats", "create_decoder", "NarrowbandFMDecoder", "AMDecoder", "SSBDecoder"]
