"""ctypes binding of libiq2a_b200.so (include/iq2a_b200.h).

There is NO CPU fallback: if the shared library is missing it is built with nvcc
(``iq_to_audio_b200/build.py``); if that fails, or no CUDA device is present when a
compute entry point is called, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libiq2a_b200.so"

OK, ERR_INVALID, ERR_STATE, ERR_CUDA, ERR_NOMEM = 0, -1, -2, -3, -4
CODEC_IDS = {"pcm_s16le": 0, "pcm_u8": 1, "pcm_f32le": 2, "complex64": 2}
ORDER_IDS = {"iq": 0, "qi": 1, "iq_inv": 2, "qi_inv": 3}
MODE_IDS = {"nfm": 0, "fm": 0, "am": 1, "usb": 2, "ssb": 2, "lsb": 3, "iq": 4, "none": 4, "pass": 4}
FRAME_BYTES = {0: 4, 1: 2, 2: 8}

#: every symbol include/iq2a_b200.h declares (tests/test_cabi_symbols.py checks the header against this)
SYMBOLS = (
    "iq2a_last_error", "iq2a_version", "iq2a_device_count", "iq2a_host_alloc", "iq2a_host_free",
    "iq2a_bank_create", "iq2a_bank_destroy", "iq2a_bank_info_get", "iq2a_bank_reset",
    "iq2a_bank_get_state", "iq2a_bank_set_state", "iq2a_bank_process_chunk",
    "iq2a_bank_submit_chunk", "iq2a_bank_collect_chunk", "iq2a_bank_submit_chunks", "iq2a_bank_collect_chunks",
    "iq2a_bank_process_resident", "iq2a_bank_process_resident_async", "iq2a_bank_launch_count",
    "iq2a_bank_copy_gtable", "iq2a_bank_set_timing", "iq2a_bank_get_timing", "iq2a_bank_set_sm_reserve",
    "iq2a_unpack_mix", "iq2a_fir", "iq2a_decimate", "iq2a_demod", "iq2a_scan",
    "iq2a_resampler_create", "iq2a_resampler_destroy", "iq2a_resampler_process", "iq2a_resampler_flush",
    "iq2a_resampler_max_outputs",
    "iq2a_psd", "iq2a_spectrum_create", "iq2a_spectrum_destroy", "iq2a_spectrum_push", "iq2a_spectrum_counts",
    "iq2a_spectrum_result",
)


class ChannelDesc(C.Structure):
    _fields_ = [("freq_offset_hz", C.c_double), ("mix_sign", C.c_int32), ("mode", C.c_int32),
                ("taps", C.POINTER(C.c_double)), ("ntaps", C.c_int32), ("agc_enabled", C.c_int32),
                ("deemph_us", C.c_double)]


class BankConfig(C.Structure):
    _fields_ = [("sample_rate", C.c_double), ("decimation", C.c_int32), ("codec", C.c_int32),
                ("iq_order", C.c_int32), ("n_channels", C.c_int32), ("fft_size", C.c_int32),
                ("ref_chunk", C.c_int64), ("device", C.c_int32), ("reserved", C.c_int32)]


class BankInfo(C.Structure):
    _fields_ = [("fft_size", C.c_int32), ("overlap_rows", C.c_int32), ("rows_per_block", C.c_int32),
                ("n_channels", C.c_int32), ("hop", C.c_int64), ("halo", C.c_int64), ("fs_channel", C.c_double),
                ("kernel_generation", C.c_int32), ("reserved", C.c_int32)]


class ChannelState(C.Structure):
    _fields_ = [("prev_re", C.c_float), ("prev_im", C.c_float), ("dc_x", C.c_float), ("dc_y", C.c_float),
                ("deemph_z", C.c_double), ("peak", C.c_float), ("reserved", C.c_float)]

    @classmethod
    def fresh(cls) -> "ChannelState":
        return cls(1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0)   # decoders/nfm.py:15 -- prev = 1+0j


_lib = None


def load() -> C.CDLL:
    """Load (building first if needed) the shared library.  Raises if it cannot be had."""
    global _lib
    if _lib is not None:
        return _lib
    path = LIB_PATH
    if os.environ.get("IQ2A_LIB"):           # development aid: A/B a kernel variant built by tools/build_variant.py
        path = Path(os.environ["IQ2A_LIB"])
        if not path.exists():
            raise RuntimeError(f"IQ2A_LIB={path} does not exist")
    elif not LIB_PATH.exists():
        from . import build as _build
        _build.build()
    lib = C.CDLL(os.fspath(path))
    vp, i32, i64, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    fp = C.c_void_p          # float* passed as raw addresses (numpy .ctypes.data or device pointers)
    lib.iq2a_last_error.restype = C.c_char_p
    lib.iq2a_last_error.argtypes = []
    lib.iq2a_version.restype = C.c_int
    lib.iq2a_device_count.argtypes = [C.POINTER(i32)]
    lib.iq2a_host_alloc.argtypes = [C.POINTER(vp), i64]
    lib.iq2a_host_free.argtypes = [vp]
    lib.iq2a_bank_create.argtypes = [C.POINTER(BankConfig), C.POINTER(ChannelDesc), C.POINTER(vp)]
    lib.iq2a_bank_destroy.argtypes = [vp]
    lib.iq2a_bank_destroy.restype = None
    lib.iq2a_bank_info_get.argtypes = [vp, C.POINTER(BankInfo)]
    lib.iq2a_bank_reset.argtypes = [vp]
    lib.iq2a_bank_get_state.argtypes = [vp, C.POINTER(ChannelState), C.POINTER(i64)]
    lib.iq2a_bank_set_state.argtypes = [vp, C.POINTER(ChannelState)]
    lib.iq2a_bank_process_chunk.argtypes = [vp, vp, i64, fp, fp, fp, i64, C.POINTER(i64), C.POINTER(f64)]
    lib.iq2a_bank_submit_chunk.argtypes = [vp, vp, i64, i32]
    lib.iq2a_bank_collect_chunk.argtypes = [vp, fp, fp, fp, i64, C.POINTER(i64), C.POINTER(f64)]
    lib.iq2a_bank_submit_chunks.argtypes = [vp, vp, i64, i64, i32]
    lib.iq2a_bank_collect_chunks.argtypes = [vp, fp, fp, fp, i64, C.POINTER(i64), C.POINTER(f64), i64, C.POINTER(i64),
                                             C.POINTER(i64)]
    lib.iq2a_bank_process_resident.argtypes = [vp, vp, i64, i64, i64, i64, i32, fp, fp, fp, i64,
                                               C.POINTER(i64), C.POINTER(f64), i64]
    lib.iq2a_bank_process_resident_async.argtypes = [vp, vp, i64, i64, i64, i64, i32, fp, fp, fp, i64, vp]
    lib.iq2a_bank_launch_count.argtypes = [vp, C.POINTER(i64)]
    lib.iq2a_bank_copy_gtable.argtypes = [vp, fp, i64]
    lib.iq2a_bank_set_timing.argtypes = [vp, i32]
    lib.iq2a_bank_set_sm_reserve.argtypes = [vp, i32]
    lib.iq2a_bank_get_timing.argtypes = [vp, C.POINTER(f64), C.POINTER(f64), C.POINTER(f64), C.POINTER(i64)]
    lib.iq2a_unpack_mix.argtypes = [vp, i64, i32, i32, f64, f64, fp, i32]
    lib.iq2a_fir.argtypes = [fp, i64, fp, C.POINTER(f64), i32, fp, i32]
    lib.iq2a_decimate.argtypes = [fp, i64, i32, i64, fp, C.POINTER(i64), i32]
    lib.iq2a_demod.argtypes = [i32, i32, f64, fp, i64, C.POINTER(ChannelState), fp, C.POINTER(f64), i32]
    lib.iq2a_scan.argtypes = [i32, f64, fp, i64, C.POINTER(ChannelState), fp, i32]
    lib.iq2a_resampler_create.argtypes = [i32, i32, i32, i32, C.POINTER(vp)]
    lib.iq2a_resampler_destroy.argtypes = [vp]
    lib.iq2a_resampler_destroy.restype = None
    lib.iq2a_resampler_process.argtypes = [vp, fp, i64, i64, fp, i64, C.POINTER(i64)]
    lib.iq2a_resampler_flush.argtypes = [vp, fp, i64, C.POINTER(i64)]
    lib.iq2a_resampler_max_outputs.argtypes = [vp, i64, C.POINTER(i64)]
    lib.iq2a_psd.argtypes = [vp, i64, i32, i32, i32, f64, fp, i32]
    lib.iq2a_spectrum_create.argtypes = [i32, i32, i32, f64, i32, i32, i32, C.POINTER(vp)]
    lib.iq2a_spectrum_destroy.argtypes = [vp]
    lib.iq2a_spectrum_destroy.restype = None
    lib.iq2a_spectrum_push.argtypes = [vp, vp, i64]
    lib.iq2a_spectrum_counts.argtypes = [vp, C.POINTER(i64), C.POINTER(i32), C.POINTER(i64)]
    lib.iq2a_spectrum_result.argtypes = [vp, fp, fp, fp]
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if name not in ("iq2a_last_error", "iq2a_bank_destroy", "iq2a_resampler_destroy", "iq2a_spectrum_destroy"):
            fn.restype = C.c_int
    _lib = lib
    return lib


def check(rc: int) -> None:
    """Map a C status onto the exception types the reference raises for the same condition
    (ValueError for configuration, RuntimeError for environment/ordering)."""
    if rc == OK:
        return
    msg = load().iq2a_last_error().decode("utf-8", "replace")
    if rc == ERR_INVALID:
        raise ValueError(msg)
    if rc == ERR_NOMEM:
        raise MemoryError(msg)
    raise RuntimeError(msg)


def device_count() -> int:
    n = C.c_int32(0)
    rc = load().iq2a_device_count(C.byref(n))
    return int(n.value) if rc == OK else 0


def ptr(a: np.ndarray | None) -> int | None:
    return None if a is None else a.ctypes.data
