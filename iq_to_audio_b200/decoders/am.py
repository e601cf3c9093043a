"""AM envelope detector with DC blocking (ref: src/iq_to_audio/decoders/am.py)."""
from __future__ import annotations

import numpy as np

from .. import _lib
from .base import DecoderStats, _GpuChannelDecoder


class AMDecoder(_GpuChannelDecoder):
    name = "am"
    _mode_id = _lib.MODE_IDS["am"]

    def __init__(self, dc_radius: float = 0.995):
        super().__init__()
        if dc_radius != 0.995:
            raise ValueError("the GPU path implements the reference's fixed DC radius 0.995")

    def process(self, samples: np.ndarray) -> tuple[np.ndarray, DecoderStats | None]:
        audio, stats = self._run(samples)
        if samples.size:
            self._intermediates = {"dc_block": (audio.copy(), self._sample_rate),
                                   "audio": (audio.copy(), self._sample_rate)}
        return audio, stats
