"""Smallest run of the mirror-pair channel-bank kernel (one chunk, two channels) against generation 1."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from oracle import iq_oracle as orc
from iq_to_audio_b200.bank import ChannelBank, Target

fs, d = 10e6, 104
taps = orc.channel_taps(fs, 12_500.0, d)
rng = np.random.default_rng(1)
n = int(os.environ.get("PROBE_N", 6 * 449 * d))
raw = rng.integers(-20_000, 20_000, 2 * n, dtype=np.int16)
tg = [Target(1.0e6, taps, 1, "iq"), Target(-2.2e6, taps, -1, "iq")]
out = {}
for gen in ("v5", "v1"):
    os.environ["IQ2A_CHANNELIZER"] = gen
    with ChannelBank(fs, d, tg, ref_chunk=1 << 18, fft_size=512) as bank:
        out[gen] = (bank.kernel_generation, bank.process_chunk(raw, want_baseband=True).baseband.copy())
    print(gen, "generation", out[gen][0], "rows", out[gen][1].shape, flush=True)
err = np.abs(out["v5"][1] - out["v1"][1])
print("max |v5 - v1|", err.max(), "at", np.unravel_index(err.argmax(), err.shape), "scale", np.abs(out["v1"][1]).max())
