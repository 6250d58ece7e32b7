"""Pipeline: a thin Python owner of one aud_handle (include/auditory_b200.h).

Host arrays go through aud_process_host (pinned staging + copies inside the
library); torch CUDA tensors go through aud_process_device on torch's current
stream.  No compute happens in Python."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import AudBatch, AudDims, AudOutputs, AudParams, OUTPUT_NAMES


class Pipeline:
    def __init__(self, params: AudParams, mel_bin_pts: np.ndarray, mel_filters: np.ndarray,
                 gabor_filters: Optional[np.ndarray] = None, dct: Optional[np.ndarray] = None, device: int = 0):
        L = _lib.lib()
        self._L = L
        self.params = params
        bp = np.ascontiguousarray(mel_bin_pts, dtype=np.int32)
        mf = np.ascontiguousarray(mel_filters, dtype=np.float64)
        gf = None if gabor_filters is None or params.gabor_nf == 0 else np.ascontiguousarray(gabor_filters, dtype=np.float64)
        dm = None if dct is None else np.ascontiguousarray(dct, dtype=np.float64)
        self._h = C.c_void_p()
        _lib.check(L.aud_create(C.byref(params), bp.ctypes.data, mf.ctypes.data,
                                None if gf is None else gf.ctypes.data,
                                None if dm is None else dm.ctypes.data, device, C.byref(self._h)))
        d = AudDims()
        _lib.check(L.aud_get_dims(self._h, C.byref(d)))
        self.S, self.n_bins, self.n_mel, self.n_coefs, self.gabor_len = (
            d.segment_steps, d.n_bins, d.n_mel, d.n_coefs, int(d.gabor_len))
        self.device = device

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._L.aud_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ shapes
    def seg_count(self, n_samples: int) -> int:
        return int(self._L.aud_seg_count(self._h, int(n_samples)))

    def seg_base(self, utt_len: np.ndarray) -> np.ndarray:
        ul = np.ascontiguousarray(utt_len, dtype=np.int32)
        base = np.zeros(len(ul) + 1, dtype=np.int64)
        _lib.check(int(self._L.aud_total_segments(self._h, ul.ctypes.data, len(ul), base.ctypes.data)))
        return base

    def out_shape(self, name: str, nseg: int):
        S = self.S
        return {
            "mel": (nseg, self.n_mel, S), "mfcc": (nseg, self.n_coefs, S), "deltas": (nseg, self.n_coefs, S),
            "delta_deltas": (nseg, self.n_coefs, S), "energy": (nseg, S), "gabor": (nseg, self.gabor_len),
            "power": (nseg, self.n_bins, S), "logpower": (nseg, self.n_bins, S),
        }[name]

    def set_option(self, name: str, value: int) -> None:
        _lib.check(self._L.aud_set_option(self._h, name.encode(), int(value)))

    @property
    def launch_count(self) -> int:
        return int(self._L.aud_launch_count(self._h))

    # -------------------------------------------------------------- host path
    def process_host(self, wave: np.ndarray, utt_offset: Sequence[int], utt_len: Sequence[int],
                     want: Sequence[str] = ("mel",), add_samples: int = 0,
                     out: Optional[Dict[str, np.ndarray]] = None) -> Dict[str, np.ndarray]:
        """wave: 1-D float32 (or int16 PCM) host array holding every utterance.  Ordinary numpy memory is fine: the
        library page-locks large buffers for the call; arrays over aud_host_alloc memory skip that step."""
        off, ln, res, o = self._host_args(wave, utt_offset, utt_len, want, out)
        if wave.dtype == np.int16:
            _lib.check(self._L.aud_process_host_i16(self._h, wave.ctypes.data, off.ctypes.data, ln.ctypes.data, len(ln),
                                                    int(add_samples), C.byref(o)))
        else:
            b = AudBatch(wave.ctypes.data, off.ctypes.data, ln.ctypes.data, len(ln), int(add_samples))
            _lib.check(self._L.aud_process_host(self._h, C.byref(b), C.byref(o)))
        return res

    def _host_args(self, wave, utt_offset, utt_len, want, out):
        if wave.dtype not in (np.float32, np.int16) or not wave.flags.c_contiguous:
            raise TypeError("wave must be a C-contiguous float32 array (etensor.Float32-style input) or int16 PCM")
        off = np.ascontiguousarray(utt_offset, dtype=np.int64)
        ln = np.ascontiguousarray(utt_len, dtype=np.int32)
        if len(off) != len(ln):
            raise ValueError("utt_offset and utt_len differ in length")
        if len(ln) and (off.min() < 0 or int((off + ln).max()) > wave.size):
            raise ValueError("utterance extents fall outside the wave buffer")
        nseg = int(self.seg_base(ln)[-1])
        res = {} if out is None else out
        o = AudOutputs()
        for name in want:
            if name not in OUTPUT_NAMES:
                raise KeyError(name)
            if name not in res:
                res[name] = np.zeros(self.out_shape(name, nseg), dtype=np.float32)
            a = res[name]
            if a.dtype != np.float32 or not a.flags.c_contiguous or a.size != int(np.prod(self.out_shape(name, nseg))):
                raise ValueError(f"output buffer '{name}' has the wrong dtype / layout / size")
            setattr(o, name, a.ctypes.data)
        return off, ln, res, o

    # ------------------------------------------------------------ device path
    def process_device(self, wave, utt_offset: np.ndarray, utt_len: np.ndarray, outputs: Dict[str, "object"],
                       add_samples: int = 0, stream: Optional[int] = None) -> None:
        """wave and outputs[...] are torch CUDA float32 tensors (or anything with
        data_ptr()); utt_offset / utt_len are host int64 / int32 numpy arrays.
        Enqueues on `stream` (a raw cudaStream_t; default: torch's current
        stream) and returns without synchronising."""
        if stream is None:
            import torch
            stream = torch.cuda.current_stream().cuda_stream
        if utt_offset.dtype != np.int64 or utt_len.dtype != np.int32:
            raise TypeError("utt_offset must be int64 and utt_len int32 numpy arrays")
        o = AudOutputs()
        for name, t in outputs.items():
            if name not in OUTPUT_NAMES:
                raise KeyError(name)
            setattr(o, name, t.data_ptr())
        if getattr(wave, "element_size", lambda: 4)() == 2:      # int16 PCM tensor
            _lib.check(self._L.aud_process_device_i16(self._h, wave.data_ptr(), utt_offset.ctypes.data, utt_len.ctypes.data,
                                                      len(utt_len), int(add_samples), C.byref(o), C.c_void_p(stream)))
            return
        b = AudBatch(wave.data_ptr(), utt_offset.ctypes.data, utt_len.ctypes.data, len(utt_len), int(add_samples))
        _lib.check(self._L.aud_process_device(self._h, C.byref(b), C.byref(o), C.c_void_p(stream)))


def process_host_multi(pipes: Sequence[Pipeline], wave: np.ndarray, utt_offset: Sequence[int], utt_len: Sequence[int],
                       want: Sequence[str] = ("mel",), add_samples: int = 0,
                       out: Optional[Dict[str, np.ndarray]] = None) -> Dict[str, np.ndarray]:
    """One batch over several GPUs of the box (aud_process_host_multi): `pipes` are Pipelines created with the same
    parameters on different devices.  Utterances are cut into contiguous blocks, one host thread per GPU inside the
    library, every GPU writing its own range of the output arrays; no collective."""
    if not pipes:
        raise ValueError("no pipelines")
    first = pipes[0]
    off, ln, res, o = first._host_args(wave, utt_offset, utt_len, want, out)
    hs = (C.c_void_p * len(pipes))(*[p._h for p in pipes])
    L = first._L
    if wave.dtype == np.int16:
        _lib.check(L.aud_process_host_multi_i16(hs, len(pipes), wave.ctypes.data, off.ctypes.data, ln.ctypes.data, len(ln),
                                                int(add_samples), C.byref(o)))
    else:
        b = AudBatch(wave.ctypes.data, off.ctypes.data, ln.ctypes.data, len(ln), int(add_samples))
        _lib.check(L.aud_process_host_multi(hs, len(pipes), C.byref(b), C.byref(o)))
    return res
