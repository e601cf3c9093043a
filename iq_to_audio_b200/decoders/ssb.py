"""SSB product detector + DC blocker + per-chunk AGC (ref: src/iq_to_audio/decoders/ssb.py).

As in the reference, USB and LSB produce the same audio (real(conj(s)) == real(s),
ssb.py:42-43) and the AGC gain restarts at 1.0 on every `process` call (ssb.py:72)."""
from __future__ import annotations

import numpy as np

from .. import _lib
from .base import DecoderStats, _GpuChannelDecoder


class SSBDecoder(_GpuChannelDecoder):
    name = "ssb"

    def __init__(self, sideband: str, agc_enabled: bool, dc_radius: float = 0.995,
                 agc_target_dbfs: float = -12.0, agc_decay: float = 0.001):
        super().__init__()
        sideband = sideband.lower()
        if sideband not in {"usb", "lsb"}:
            raise ValueError("sideband must be 'usb' or 'lsb'")
        if dc_radius != 0.995 or agc_target_dbfs != -12.0 or agc_decay != 0.001:
            raise ValueError("the GPU path implements the reference's fixed DC/AGC constants")
        self._sideband = sideband
        self._mode_id = _lib.MODE_IDS[sideband]
        self._agc = bool(agc_enabled)

    def process(self, samples: np.ndarray) -> tuple[np.ndarray, DecoderStats | None]:
        audio, stats = self._run(samples)
        if samples.size:
            self._intermediates = {"audio": (audio.copy(), self._sample_rate)}
        return audio, stats
