"""
c_oracle.py -- ctypes binding of oracle/liboracle.so (auditory_oracle.c).

TEST INFRASTRUCTURE ONLY (see np_oracle.py header): used by tests/, smoke()
and bench.py's cpu_baseline / --impl reference legs; never by the product.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


class OrcParams(C.Structure):
    _fields_ = [
        ("sample_rate", C.c_int32),
        ("win_ms", C.c_double), ("step_ms", C.c_double), ("segment_ms", C.c_double), ("stride_ms", C.c_double),
        ("border_steps", C.c_int32),
        ("comp_log_pow", C.c_int32),
        ("log_min", C.c_double), ("log_offset", C.c_double), ("prev_smooth", C.c_double), ("cur_smooth", C.c_double),
        ("n_filters", C.c_int32),
        ("lo_hz", C.c_double), ("hi_hz", C.c_double), ("mel_log_off", C.c_double), ("mel_log_min", C.c_double),
        ("renorm", C.c_int32),
        ("renorm_min", C.c_double), ("renorm_max", C.c_double),
        ("mfcc", C.c_int32), ("deltas", C.c_int32), ("n_coefs", C.c_int32),
        ("size_x", C.c_int32), ("size_y", C.c_int32), ("stride_x", C.c_int32), ("stride_y", C.c_int32),
        ("gain", C.c_double),
        ("distribute", C.c_int32),
        ("pools_y", C.c_int32), ("pools_x", C.c_int32), ("units_y", C.c_int32), ("units_x", C.c_int32),
        ("by_time", C.c_int32),
        ("rebuild_plan", C.c_int32),
    ]


class OrcGaborSpec(C.Structure):
    _fields_ = [
        ("off", C.c_int32),
        ("wave_len", C.c_double), ("orientation", C.c_double), ("sigma_width", C.c_double),
        ("sigma_length", C.c_double), ("phase_offset", C.c_double),
        ("circle_edge", C.c_int32), ("circular", C.c_int32),
    ]


class OrcOutputs(C.Structure):
    _fields_ = [(n, C.POINTER(C.c_double)) for n in
                ("mel", "mfcc", "deltas", "delta_deltas", "energy", "power", "logpower")] + \
               [("gabor", C.POINTER(C.c_float))]


class OrcOutputsF32(C.Structure):
    _fields_ = [(n, C.POINTER(C.c_float)) for n in ("mel", "mfcc", "energy", "gabor")]


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile recipe if missing."""
    src = os.path.join(_HERE, "auditory_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_msec_to_samples.restype = C.c_int32
        L.orc_msec_to_samples.argtypes = [C.c_double, C.c_int32]
        L.orc_fft.restype = None
        L.orc_fft.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        L.orc_dct1.restype = None
        L.orc_dct1.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.orc_mel_init.restype = C.c_int32
        L.orc_mel_init.argtypes = [C.POINTER(OrcParams), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_gabor_to_tensor.restype = C.c_int32
        L.orc_gabor_to_tensor.argtypes = [C.POINTER(OrcParams), C.POINTER(OrcGaborSpec), C.c_int32, C.c_void_p]
        L.orc_gabor_convolve.restype = C.c_int32
        L.orc_gabor_convolve.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                         C.c_int32, C.c_int32, C.c_double, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]
        L.orc_env_create.restype = C.c_int32
        L.orc_env_create.argtypes = [C.POINTER(OrcParams), C.POINTER(OrcGaborSpec), C.c_int32, C.POINTER(C.c_void_p)]
        L.orc_env_destroy.restype = None
        L.orc_env_destroy.argtypes = [C.c_void_p]
        L.orc_env_dims.restype = C.c_int32
        L.orc_env_dims.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_env_seg_count.restype = C.c_int32
        L.orc_env_seg_count.argtypes = [C.c_void_p, C.c_int32]
        L.orc_env_binpts.restype = C.POINTER(C.c_int32)
        L.orc_env_binpts.argtypes = [C.c_void_p]
        L.orc_env_mel_filters.restype = C.POINTER(C.c_double)
        L.orc_env_mel_filters.argtypes = [C.c_void_p]
        L.orc_env_gabor.restype = C.POINTER(C.c_double)
        L.orc_env_gabor.argtypes = [C.c_void_p]
        L.orc_env_process.restype = C.c_int32
        L.orc_env_process.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(OrcOutputs)]
        L.orc_batch_process_f32.restype = C.c_int64
        L.orc_batch_process_f32.argtypes = [C.POINTER(OrcParams), C.POINTER(OrcGaborSpec), C.c_int32, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                            C.POINTER(OrcOutputsF32), C.POINTER(C.c_double)]
        L.orc_online_cpus.restype = C.c_int32
        _lib = L
    return _lib


def default_params(**kw) -> OrcParams:
    """SndEnv.Defaults() + DFT.Defaults() + Mel.Defaults() values
    (sound/sndenv.go:64-71, dft/dft.go:33-39, mel/mel.go:69-74,171-180;
    Renorm forced off by mel.go:80)."""
    p = OrcParams(
        sample_rate=16000, win_ms=25.0, step_ms=10.0, segment_ms=100.0, stride_ms=100.0, border_steps=2,
        comp_log_pow=1, log_min=-100.0, log_offset=1.0, prev_smooth=0.0, cur_smooth=1.0,
        n_filters=32, lo_hz=0.0, hi_hz=8000.0, mel_log_off=0.0, mel_log_min=-10.0,
        renorm=0, renorm_min=-6.0, renorm_max=4.0, mfcc=1, deltas=1, n_coefs=13,
        size_x=1, size_y=1, stride_x=1, stride_y=1, gain=1.0, distribute=0,
        pools_y=0, pools_x=0, units_y=1, units_x=1, by_time=0, rebuild_plan=0)
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def processspeech_specs() -> List[OrcGaborSpec]:
    """examples/processspeech/processspeech.go:226-253 parameter values."""
    specs = []
    for orient in (0.0, 45.0, 90.0, 135.0):
        for ph in (0.0, 1.5708):
            specs.append(OrcGaborSpec(off=0, wave_len=2.0, orientation=orient, sigma_width=0.5, sigma_length=0.5,
                                      phase_offset=ph, circle_edge=1, circular=0))
    return specs


def with_processspeech_gabor(p: OrcParams, out4d: bool = True, by_time: bool = False) -> OrcParams:
    p.size_x = p.size_y = 9
    p.stride_x = p.stride_y = 3
    p.gain = 2.0
    p.distribute = 0
    p.by_time = int(by_time)
    if out4d:
        p.pools_y, p.pools_x, p.units_y, p.units_x = 8, 2, 2, 8
    else:
        p.pools_y = p.pools_x = 0
        p.units_y, p.units_x = 16, 16
    return p


def _spec_array(specs: Sequence[OrcGaborSpec]):
    arr = (OrcGaborSpec * max(len(specs), 1))()
    for i, s in enumerate(specs):
        arr[i] = s
    return arr


class Env:
    """One SndEnv-equivalent (C twin)."""

    def __init__(self, params: OrcParams, specs: Sequence[OrcGaborSpec] = ()):
        self.L = lib()
        self.p = params
        self.specs = list(specs)
        self._h = C.c_void_p()
        rc = self.L.orc_env_create(C.byref(params), _spec_array(specs), len(specs), C.byref(self._h))
        if rc != 0:
            raise ValueError(f"orc_env_create failed: {rc}")
        d = np.zeros(10, dtype=np.int32)
        self.L.orc_env_dims(self._h, d.ctypes.data)
        (self.win, self.step, self.stride, self.S, self.B, self.nf, self.ncoef,
         self.gabor_nf, self.gabor_len, self.seg_samples) = [int(x) for x in d]

    def __del__(self):
        if getattr(self, "_h", None):
            self.L.orc_env_destroy(self._h)
            self._h = None

    def seg_count(self, n: int) -> int:
        return int(self.L.orc_env_seg_count(self._h, n))

    @property
    def binpts(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.L.orc_env_binpts(self._h), shape=(self.nf + 2,)).copy()

    @property
    def mel_filters(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.L.orc_env_mel_filters(self._h), shape=(self.nf, self.nf + 2)).copy()

    @property
    def gabor_filters(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.L.orc_env_gabor(self._h),
                                     shape=(self.gabor_nf, self.p.size_y, self.p.size_x)).copy()

    def process(self, signal: np.ndarray, add_ms: int = 0, want_power: bool = False) -> dict:
        sig = np.ascontiguousarray(signal, dtype=np.float64)
        nseg = max(self.seg_count(len(sig)), 0)
        S, B, nf, nc = self.S, self.B, self.nf, self.ncoef
        out = {"mel": np.zeros((nseg, nf, S)), "energy": np.zeros((nseg, S))}
        if self.p.mfcc:
            out["mfcc"] = np.zeros((nseg, nc, S))
            if self.p.deltas:
                out["deltas"] = np.zeros((nseg, nc, S))
                out["delta_deltas"] = np.zeros((nseg, nc, S))
        if want_power:
            out["power"] = np.zeros((nseg, B, S))
            out["logpower"] = np.zeros((nseg, B, S))
        if self.gabor_nf > 0:
            out["gabor"] = np.zeros((nseg, self.gabor_len), dtype=np.float32)
        o = OrcOutputs()
        for k, v in out.items():
            ptr_t = C.POINTER(C.c_float) if k == "gabor" else C.POINTER(C.c_double)
            setattr(o, k, v.ctypes.data_as(ptr_t))
        rc = self.L.orc_env_process(self._h, sig.ctypes.data, len(sig), add_ms, C.byref(o))
        # SegCnt goes negative for very short signals (sndenv.go:263-265) and is returned as is: no segments, no error
        if rc < 0 and self.seg_count(len(sig)) > 0:
            raise RuntimeError(f"orc_env_process failed: {rc}")   # -5: agabor.Convolve index out of range (the reference panics)
        return out


def batch_process_f32(params: OrcParams, specs: Sequence[OrcGaborSpec], wave: np.ndarray, utt_off: np.ndarray,
                      utt_len: np.ndarray, nthreads: int = 1, add_ms: int = 0, want: Sequence[str] = ()):
    """Timed CPU-baseline driver.  Returns (total_segments, outputs dict, checksum)."""
    L = lib()
    wave = np.ascontiguousarray(wave, dtype=np.float32)
    utt_off = np.ascontiguousarray(utt_off, dtype=np.int64)
    utt_len = np.ascontiguousarray(utt_len, dtype=np.int32)
    n = len(utt_len)
    outs = {}
    o_ptr = None
    seg_base_ptr = None
    if want:
        env = Env(params, specs)
        segs = np.array([max(env.seg_count(int(x)), 0) for x in utt_len], dtype=np.int64)
        seg_base = np.concatenate([[0], np.cumsum(segs)]).astype(np.int64)
        tot = int(seg_base[-1])
        shapes = {"mel": (tot, env.nf, env.S), "mfcc": (tot, env.ncoef, env.S), "energy": (tot, env.S),
                  "gabor": (tot, max(env.gabor_len, 1))}
        o = OrcOutputsF32()
        for k in want:
            outs[k] = np.zeros(shapes[k], dtype=np.float32)
            setattr(o, k, outs[k].ctypes.data_as(C.POINTER(C.c_float)))
        o_ptr = C.byref(o)
        seg_base_ptr = seg_base.ctypes.data
    cs = C.c_double(0.0)
    tot = L.orc_batch_process_f32(C.byref(params), _spec_array(specs), len(specs), wave.ctypes.data,
                                  utt_off.ctypes.data, utt_len.ctypes.data, n, add_ms, nthreads,
                                  seg_base_ptr, o_ptr, C.byref(cs))
    if tot < 0:
        raise RuntimeError(f"orc_batch_process_f32 failed: {tot}")
    return int(tot), outs, float(cs.value)


def fft(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=np.complex128)
    re = np.ascontiguousarray(x.real)
    im = np.ascontiguousarray(x.imag)
    ore = np.zeros_like(re)
    oim = np.zeros_like(im)
    lib().orc_fft(re.ctypes.data, im.ctypes.data, len(re), ore.ctypes.data, oim.ctypes.data)
    return ore + 1j * oim


def dct1(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.zeros_like(x)
    lib().orc_dct1(x.ctypes.data, len(x), y.ctypes.data)
    return y


def online_cpus() -> int:
    return int(lib().orc_online_cpus())
