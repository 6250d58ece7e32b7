"""ctypes binding of libauditory_b200.so (C-ABI: include/auditory_b200.h).

The library is hand-written CUDA for sm_100a and there is no CPU fallback: if
the shared object is missing this module raises at import of the symbol table,
and every processing call fails with AUD_ERR_CUDA when no GPU is usable.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# AUD_B200_LIB: another build of the same library (the debug-check build, `make -C auditory_b200/csrc debug`)
LIB_PATH = os.environ.get("AUD_B200_LIB") or os.path.join(_HERE, "lib", "libauditory_b200.so")

AUD_OK = 0
AUD_ERR_INVALID = -1
AUD_ERR_UNSUPPORTED = -2
AUD_ERR_CUDA = -3
AUD_ERR_NOMEM = -4
AUD_ERR_PANIC = -5


class AudError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"auditory_b200 error {code}: {msg}")
        self.code = code
        self.msg = msg


class AudGaborSpec(C.Structure):
    _fields_ = [
        ("off", C.c_int32),
        ("wave_len", C.c_double), ("orientation", C.c_double), ("sigma_width", C.c_double),
        ("sigma_length", C.c_double), ("phase_offset", C.c_double),
        ("circle_edge", C.c_int32), ("circular", C.c_int32),
    ]


class AudParams(C.Structure):
    _fields_ = [
        ("sample_rate", C.c_int32),
        ("win_samples", C.c_int32), ("step_samples", C.c_int32), ("segment_samples", C.c_int32),
        ("stride_samples", C.c_int32), ("segment_steps", C.c_int32), ("border_steps", C.c_int32),
        ("comp_log_pow", C.c_int32),
        ("log_min", C.c_double), ("log_offset", C.c_double), ("prev_smooth", C.c_double), ("cur_smooth", C.c_double),
        ("n_mel", C.c_int32),
        ("mel_log_off", C.c_double), ("mel_log_min", C.c_double),
        ("renorm", C.c_int32),
        ("renorm_min", C.c_double), ("renorm_scale", C.c_double),
        ("mfcc", C.c_int32), ("n_coefs", C.c_int32), ("deltas", C.c_int32), ("mfcc_c0_energy", C.c_int32),
        ("gabor_nf", C.c_int32), ("gabor_size_x", C.c_int32), ("gabor_size_y", C.c_int32),
        ("gabor_stride_x", C.c_int32), ("gabor_stride_y", C.c_int32),
        ("gabor_gain", C.c_double),
        ("gabor_out_dims", C.c_int32),
        ("gabor_shape", C.c_int32 * 4),
        ("gabor_by_time", C.c_int32),
    ]


class AudDims(C.Structure):
    _fields_ = [("segment_steps", C.c_int32), ("n_bins", C.c_int32), ("n_mel", C.c_int32), ("n_coefs", C.c_int32),
                ("gabor_len", C.c_int64)]


class AudBatch(C.Structure):
    _fields_ = [("wave", C.c_void_p), ("utt_offset", C.c_void_p), ("utt_len", C.c_void_p),
                ("n_utt", C.c_int32), ("add_samples", C.c_int32)]


class AudDftParams(C.Structure):
    _fields_ = [("comp_log_pow", C.c_int32), ("log_min", C.c_double), ("log_offset", C.c_double),
                ("prev_smooth", C.c_double), ("cur_smooth", C.c_double)]


class AudMelParams(C.Structure):
    _fields_ = [("n_filters", C.c_int32), ("log_off", C.c_double), ("log_min", C.c_double), ("renorm", C.c_int32),
                ("renorm_min", C.c_double), ("renorm_scale", C.c_double)]


class AudFffbParams(C.Structure):
    _fields_ = [("on", C.c_int32)] + [(n, C.c_float) for n in ("gi", "ff", "fb", "fb_tau", "max_vs_avg", "ff0")]


class AudKwtaParams(C.Structure):
    _fields_ = [("on", C.c_int32), ("iters", C.c_int32), ("del_act_thr", C.c_float),
                ("lay_fffb", AudFffbParams), ("pool_fffb", AudFffbParams)] + \
               [(n, C.c_float) for n in ("xx1_thr", "xx1_gain", "xx1_nvar", "xx1_vm_act_thr", "xx1_sig_mult", "xx1_sig_mult_pow",
                                         "xx1_sig_gain", "xx1_interp_range", "xx1_gain_cor_range", "xx1_gain_cor", "act_tau",
                                         "gbar_e", "gbar_l", "gbar_i", "gbar_k", "erev_e", "erev_l", "erev_i", "erev_k")] + \
               [("pool_mode", C.c_int32), ("neigh_on", C.c_int32), ("neigh_gi", C.c_float)]


OUTPUT_NAMES = ("mel", "mfcc", "deltas", "delta_deltas", "energy", "gabor", "power", "logpower")


class AudOutputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in OUTPUT_NAMES]


# every symbol include/auditory_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "aud_msec_to_samples": (C.c_int32, [C.c_double, C.c_int32]),
    "aud_freq_to_mel": (C.c_double, [C.c_double]),
    "aud_mel_to_freq": (C.c_double, [C.c_double]),
    "aud_freq_to_bin": (C.c_int32, [C.c_double, C.c_double, C.c_double]),
    "aud_mel_init_filters": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_double,
                                         C.c_void_p, C.c_void_p, C.c_void_p]),
    "aud_gabor_to_tensor": (C.c_int32, [C.POINTER(AudGaborSpec), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "aud_dct1_matrix": (None, [C.c_int32, C.c_int32, C.c_void_p]),
    "aud_params_defaults": (C.c_int32, [C.POINTER(AudParams), C.c_int32, C.c_double, C.c_double, C.c_double,
                                        C.c_double, C.c_int32]),
    "aud_create": (C.c_int32, [C.POINTER(AudParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                               C.POINTER(C.c_void_p)]),
    "aud_destroy": (None, [C.c_void_p]),
    "aud_get_dims": (C.c_int32, [C.c_void_p, C.POINTER(AudDims)]),
    "aud_seg_count": (C.c_int32, [C.c_void_p, C.c_int32]),
    "aud_total_segments": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "aud_process_host": (C.c_int32, [C.c_void_p, C.POINTER(AudBatch), C.POINTER(AudOutputs)]),
    "aud_process_host_multi": (C.c_int32, [C.POINTER(C.c_void_p), C.c_int32, C.POINTER(AudBatch), C.POINTER(AudOutputs)]),
    "aud_process_host_multi_i16": (C.c_int32, [C.POINTER(C.c_void_p), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                               C.c_int32, C.POINTER(AudOutputs)]),
    "aud_process_device": (C.c_int32, [C.c_void_p, C.POINTER(AudBatch), C.POINTER(AudOutputs), C.c_void_p]),
    "aud_process_host_i16": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                         C.POINTER(AudOutputs)]),
    "aud_process_device_i16": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                           C.POINTER(AudOutputs), C.c_void_p]),
    "aud_gabor_convolve": (C.c_int32, [C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                       C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_int32, C.c_void_p,
                                       C.c_int32, C.c_void_p]),
    "aud_dft_filter": (C.c_int32, [C.c_int32, C.POINTER(AudDftParams), C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "aud_mel_filter_dft": (C.c_int32, [C.c_int32, C.POINTER(AudMelParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                       C.c_int32, C.c_void_p]),
    "aud_cepstrum_dct": (C.c_int32, [C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "aud_kwta_defaults": (None, [C.POINTER(AudKwtaParams)]),
    "aud_apply_kwta": (C.c_int32, [C.c_int32, C.POINTER(AudKwtaParams), C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                   C.c_int32, C.c_void_p, C.c_void_p]),
    "aud_host_alloc": (C.c_void_p, [C.c_uint64]),
    "aud_host_free": (None, [C.c_void_p]),
    "aud_launch_count": (C.c_int64, [C.c_void_p]),
    "aud_set_option": (C.c_int32, [C.c_void_p, C.c_char_p, C.c_int64]),
    "aud_measure_fp32": (C.c_int32, [C.c_int32, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "aud_last_error": (C.c_char_p, []),
    "aud_version": (C.c_int32, []),
}

_lib = None


def lib():
    """Load the shared library (once).  Raises if it has not been built:
    run `python -c 'import __graft_entry__ as g; g.build()'` or
    `make -C auditory_b200/csrc`."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built and there is no fallback path")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)       # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error() -> str:
    return (lib().aud_last_error() or b"").decode("utf-8", "replace")


def check(rc: int) -> int:
    if rc < 0:
        raise AudError(rc, last_error())
    return rc
