#!/usr/bin/env python
"""Golden vectors for the spectrum previews from the UNMODIFIED reference module (build container only).

    python tests/golden/make_spectrum_golden.py

Imports ``iq_to_audio.spectrum`` from /root/reference/src, runs ``compute_psd`` and ``streaming_waterfall`` on
small seeded inputs (rebuilt by tests/_spectrum_cases.py from the same recipe) and stores what they return in
tests/golden/spectrum_vectors.npz.  Only outputs are stored, no reference code.
"""
from __future__ import annotations

import importlib.util
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
sys.path.insert(0, str(ROOT))
from tests._spectrum_cases import PSD_CASES, WATERFALL_CASES, psd_input, waterfall_chunks  # noqa: E402

REF = Path("/root/reference/src/iq_to_audio/spectrum.py")


def main() -> None:
    if not REF.exists():
        raise SystemExit("reference tree not present; goldens can only be regenerated in the build container")
    spec = importlib.util.spec_from_file_location("ref_spectrum", REF)
    ref = importlib.util.module_from_spec(spec)
    sys.modules["ref_spectrum"] = ref
    spec.loader.exec_module(ref)
    out = {}
    for name, case in PSD_CASES.items():
        freqs, psd = ref.compute_psd(psd_input(case), case["fs"], case["nfft"])
        out[f"psd_{name}_freqs_edge"] = np.concatenate([freqs[:4], freqs[-4:]])
        out[f"psd_{name}_db"] = psd
    for name, case in WATERFALL_CASES.items():
        freqs, avg, wf, frames = ref.streaming_waterfall(iter(waterfall_chunks(case)), case["fs"], nfft=case["nfft"],
                                                         hop=case["hop"], max_slices=case["max_slices"])
        out[f"wf_{name}_freqs_edge"] = np.concatenate([freqs[:4], freqs[-4:]])
        out[f"wf_{name}_avg"] = avg
        out[f"wf_{name}_times"] = wf.times
        out[f"wf_{name}_matrix"] = wf.matrix
        out[f"wf_{name}_frames"] = np.int64(frames)
        # the start indices the reference reports, one per window, straight from its generator
        starts = [s for s, _ in ref._sliding_windows(iter(waterfall_chunks(case)), nfft=case["nfft"],
                                                     hop=max(1, case["hop"] or case["nfft"] // 4))]
        out[f"wf_{name}_starts"] = np.asarray(starts, dtype=np.int64)
    np.savez_compressed(HERE / "spectrum_vectors.npz", **out)
    print("wrote", HERE / "spectrum_vectors.npz", sum(v.nbytes for v in out.values()), "bytes raw")


if __name__ == "__main__":
    main()
