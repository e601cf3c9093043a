"""Command line surface of the hot path (ref: src/iq_to_audio/cli.py:151-412 flags, :415-741 main).

Only the flags that reach the channelize-and-demodulate path are kept (`--in --ft --bw --fc --fs-ch
--demod --deemph --no-agc --out --dump-iq --chunk --filter-block --iq-order --mix-sign
--input-format --input-sample-rate --preview --probe-only --benchmark*`); the GUI, squelch and
digital-decoder flags belong to subsystems outside the scope table.  All `--ft` targets run as ONE pass over
the capture; the reference's limit of five per run (cli.py:514-515, there because every target is another full
pass) becomes MAX_TARGETS, the 256 channels of the wideband sweep configuration."""
from __future__ import annotations

import argparse
import logging
import math
import sys
from pathlib import Path

LOG = logging.getLogger("iq_to_audio_b200")


def positive_float(text: str) -> float:
    value = float(text)
    if value <= 0:
        raise argparse.ArgumentTypeError("must be positive")
    return value


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(prog="iq-to-audio-b200",
                                description="Extract and demodulate channels from SDR++ baseband recordings on a B200.")
    p.add_argument("--in", dest="input_path", type=Path, help="Input SDR++ baseband WAV / raw IQ file.")
    p.add_argument("--ft", dest="target_freqs", type=positive_float, action="append", default=None,
                   help="Target RF frequency in Hz; repeat up to five times (one pass for all).")
    p.add_argument("--bw", dest="bandwidth", type=positive_float, default=12_500.0)
    p.add_argument("--fc", dest="center_freq", type=positive_float)
    p.add_argument("--fs-ch", dest="fs_ch", type=positive_float, default=96_000.0)
    p.add_argument("--demod", dest="demod", choices=["nfm", "am", "usb", "lsb", "ssb", "none"], default="nfm")
    p.add_argument("--deemph", dest="deemph_us", type=positive_float, default=300.0)
    p.add_argument("--no-agc", dest="agc_enabled", action="store_false")
    p.add_argument("--out", dest="output_path", type=Path)
    p.add_argument("--dump-iq", dest="dump_iq", type=Path)
    p.add_argument("--chunk", dest="chunk_size", type=int, default=1_048_576)
    p.add_argument("--fft-workers", dest="fft_workers", type=int, help="accepted for compatibility; unused on the GPU")
    p.add_argument("--filter-block", dest="filter_block", type=int, default=65_536)
    p.add_argument("--iq-order", dest="iq_order", choices=["iq", "qi", "iq_inv", "qi_inv"], default="iq")
    p.add_argument("--mix-sign", dest="mix_sign", type=int, choices=[-1, 1])
    p.add_argument("--input-format", dest="input_format", type=str)
    p.add_argument("--input-sample-rate", dest="input_sample_rate", type=positive_float)
    p.add_argument("--preview", dest="preview_seconds", type=positive_float)
    p.add_argument("--probe-only", dest="probe_only", action="store_true")
    p.add_argument("--device", dest="device", type=int, default=0)
    p.add_argument("--benchmark", dest="benchmark", action="store_true")
    p.add_argument("--benchmark-seconds", dest="benchmark_seconds", type=positive_float, default=5.0)
    p.add_argument("--benchmark-sample-rate", dest="benchmark_sample_rate", type=positive_float, default=2_500_000.0)
    p.add_argument("--benchmark-offset", dest="benchmark_offset", type=float, default=25_000.0)
    p.add_argument("-v", "--verbose", action="store_true")
    return p


MAX_TARGETS = 256


def main(argv: list[str] | None = None) -> int:
    parser = build_parser()
    args = parser.parse_args(argv)
    logging.basicConfig(level=logging.DEBUG if args.verbose else logging.INFO, format="%(levelname)s %(message)s")
    freqs = list(args.target_freqs or [])
    if len(freqs) > MAX_TARGETS:
        parser.error(f"At most {MAX_TARGETS} target frequencies are supported per run.")
    for i, f in enumerate(freqs):
        if any(math.isclose(f, g, rel_tol=0.0, abs_tol=0.5) for g in freqs[:i]):
            parser.error("Duplicate target frequencies are not allowed.")
    shared = dict(bandwidth=args.bandwidth, center_freq=args.center_freq,
                  center_freq_source="cli" if args.center_freq is not None else None, demod_mode=args.demod,
                  fs_ch_target=args.fs_ch, deemph_us=args.deemph_us, agc_enabled=args.agc_enabled,
                  chunk_size=args.chunk_size, filter_block=args.filter_block, iq_order=args.iq_order,
                  probe_only=args.probe_only, mix_sign_override=args.mix_sign, fft_workers=args.fft_workers,
                  input_format=args.input_format, input_sample_rate=args.input_sample_rate, device=args.device)
    if args.benchmark:
        from .benchmark import run_benchmark
        try:
            return run_benchmark(seconds=args.benchmark_seconds, sample_rate=args.benchmark_sample_rate,
                                 freq_offset=args.benchmark_offset, center_freq=args.center_freq,
                                 target_freq=freqs[0] if freqs else None, base_kwargs=shared)
        except Exception as exc:
            LOG.error("Benchmark failed: %s", exc)
            return 1
    if args.input_path is None:
        parser.error("--in is required unless --benchmark is used.")
    if not freqs and not args.probe_only:
        parser.error("Provide at least one --ft target frequency.")
    from .pipeline import ProcessingCancelled, ProcessingConfig, ProcessingPipeline
    cfg = ProcessingConfig(in_path=args.input_path, target_freq=freqs[0] if freqs else 0.0, target_freqs=freqs or None,
                           output_path=args.output_path, dump_iq_path=args.dump_iq,
                           max_input_seconds=args.preview_seconds, **shared)
    try:
        for res in ProcessingPipeline(cfg).run_many():
            LOG.info("Target %.0f Hz -> %s (mix sign %d, peak %.2f dBFS)", res.target_freq, res.output_path,
                     res.mix_sign, 20.0 * math.log10(max(res.audio_peak, 1e-6)))
    except ProcessingCancelled:
        LOG.warning("Cancelled.")
        return 130
    except (ValueError, RuntimeError) as exc:
        LOG.error("%s", exc)
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
