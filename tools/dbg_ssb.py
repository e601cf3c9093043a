import sys, os
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from oracle import iq_oracle as orc
from tests import _cases
from iq_to_audio_b200.bank import ChannelBank, Target
m = _cases.manifest()["case_c_20M_am_ssb"]
fs = m["fs"]; d = 208
raw = _cases.raw_input("case_c_20M_am_ssb")
g = _cases.load("case_c_20M_usb")
t = m["targets"][1]
taps = orc.channel_taps(fs, 2800.0, d)
chunk = m["chunk"]
n = raw.size // 2
for prec in ("1", "0"):
    os.environ["IQ2A_PRECISE_SSB"] = prec
    with ChannelBank(fs, d, [Target(t["f_off"], taps, 1, "usb", 300.0, True)], ref_chunk=chunk) as bank:
        outs = [bank.process_chunk(raw[2*s:2*min(s+chunk, n)], want_baseband=True) for s in range(0, n, chunk)]
        a = np.concatenate([o.audio[0] for o in outs]); c = np.concatenate([o.clipped[0] for o in outs]); b = np.concatenate([o.baseband[0] for o in outs])
        print("precise", prec, "launches", bank.launches, "state", bank.get_state()[0])
    print(" bb match frac", np.mean(b == g["baseband"]), "bb max err", np.abs(b - g["baseband"]).max())
    print(" audio first", a[:5], "ref", g["audio"][:5])
    print(" audio max err", np.abs(a - g["audio"]).max(), "clipped max err", np.abs(c - g["clipped"]).max(), "n>1e-4:", int(np.sum(np.abs(c - g["clipped"]) > 1e-4)), "of", c.size)
    e = np.abs(a - g["audio"]); print(" first bad idx", int(np.argmax(e > 1e-4)) if (e > 1e-4).any() else None, "counts", [o.count for o in outs])
