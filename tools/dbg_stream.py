import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from oracle import iq_oracle as orc
from tests import _cases
from iq_to_audio_b200.bank import ChannelBank, Target
m = _cases.manifest()["case_b_nfm_10M"]
fs = m["fs"]; d = 104
taps = orc.channel_taps(fs, 12500.0, d)
tg = [Target(t["f_off"], taps, 1, "nfm") for t in m["targets"]]
raw = _cases.raw_input("case_b_nfm_10M")
n = raw.size // 2
chunk = 100_000
views = [raw[2 * s:2 * min(s + chunk, n)] for s in range(0, n, chunk)]
with ChannelBank(fs, d, tg, ref_chunk=chunk) as bank:
    sync = [bank.process_chunk(v, want_baseband=True) for v in views]
    sync = [(r.audio.copy(), r.clipped.copy(), r.baseband.copy()) for r in sync]
    bank.reset()
    sync2 = [bank.process_chunk(v, want_baseband=True) for v in views]
    sync2 = [(r.audio.copy(), r.clipped.copy(), r.baseband.copy()) for r in sync2]
    bank.reset()
    piped = [(r.audio.copy(), r.clipped.copy(), r.baseband.copy()) for r in bank.stream(views, want_baseband=True)]
for k, (a, b, c) in enumerate(zip(sync, sync2, piped)):
    e2 = [float(np.abs(x - y).max()) for x, y in zip(a, b)]
    e3 = [float(np.abs(x - y).max()) for x, y in zip(a, c)]
    if max(e2) > 0 or max(e3) > 0:
        idx = np.argwhere(np.abs(a[2] - c[2]) > 0)
        print(k, "sync-vs-sync", e2, "sync-vs-piped", e3, "shape", a[2].shape, "first diff idx", idx[:3].tolist(), "count", len(idx))
print("done")
