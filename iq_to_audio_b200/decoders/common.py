"""DC blocker (ref: src/iq_to_audio/decoders/common.py:6-30): y[n] = x[n] - x[n-1] + r*y[n-1]."""
from __future__ import annotations

import numpy as np

from .. import _lib
from .base import run_scan


class DCBlocker:
    def __init__(self, radius: float = 0.995):
        if not 0.0 < radius < 1.0:
            raise ValueError("radius must be between 0 and 1")
        if radius != 0.995:
            raise ValueError("the GPU path implements the reference's fixed radius 0.995")
        self.radius = radius
        self._st = _lib.ChannelState.fresh()

    @property
    def _x_prev(self) -> float:
        return float(self._st.dc_x)

    @property
    def _y_prev(self) -> float:
        return float(self._st.dc_y)

    def process(self, samples: np.ndarray) -> np.ndarray:
        if samples.size == 0:
            return samples
        return run_scan(1, 0.0, samples, self._st)
