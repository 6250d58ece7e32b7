"""agabor.Filter / agabor.FilterSet mirror (reference agabor/gabor.go:17-70).
ToTensor is the C-ABI's aud_gabor_to_tensor; Convolve (gabor.go:225-315) runs
inside the fused CUDA kernel."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import ctypes as C

import numpy as np

from . import _lib


@dataclass
class Filter:
    Off: bool = False
    WaveLen: float = 0.0
    Orientation: float = 0.0
    SigmaWidth: float = 0.0
    SigmaLength: float = 0.0
    PhaseOffset: float = 0.0
    CircleEdge: bool = False
    Circular: bool = False


@dataclass
class FilterSet:
    SizeX: int = 0
    SizeY: int = 0
    StrideX: int = 0
    StrideY: int = 0
    Gain: float = 0.0
    Distribute: bool = False
    Filters: Optional[np.ndarray] = None     # float64 [n_active, SizeY, SizeX]


def Active(specs: Sequence[Filter]) -> List[Filter]:
    """agabor/gabor.go:329-336."""
    return [s for s in specs if not s.Off]


def ToTensor(specs: Sequence[Filter], fs: FilterSet) -> None:
    """agabor/gabor.go:89-222: fills fs.Filters for the active specs."""
    n = len(specs)
    arr = (_lib.AudGaborSpec * max(n, 1))()
    for i, s in enumerate(specs):
        arr[i] = _lib.AudGaborSpec(int(s.Off), s.WaveLen, s.Orientation, s.SigmaWidth, s.SigmaLength,
                                   s.PhaseOffset, int(s.CircleEdge), int(s.Circular))
    n_act = len(Active(specs))
    out = np.zeros((n_act, max(fs.SizeY, 0), max(fs.SizeX, 0)), dtype=np.float64)
    if n_act:
        got = _lib.check(_lib.lib().aud_gabor_to_tensor(arr, n, fs.SizeX, fs.SizeY, int(fs.Distribute), out.ctypes.data))
        assert got == n_act
    fs.Filters = out


def Convolve(melData: np.ndarray, filters: FilterSet, rawOut: np.ndarray, byTime: bool, device: int = 0) -> None:
    """agabor/gabor.go:225-315 on the GPU (aud_gabor_convolve).  melData: float [NFilters, steps] or a stack
    [n, NFilters, steps]; rawOut: float32 array with 2 or 4 dims (or a stack of them), written in place --
    cells the convolution does not reach keep their values, and nothing is written when the filter is
    wider than the input, as in the reference."""
    mel = np.ascontiguousarray(melData, dtype=np.float32)
    stacked = mel.ndim == 3
    if mel.ndim not in (2, 3):
        raise ValueError("melData must be [NFilters, steps] or [n, NFilters, steps]")
    n = mel.shape[0] if stacked else 1
    shape = rawOut.shape[1:] if stacked else rawOut.shape
    if rawOut.dtype != np.float32 or not rawOut.flags.c_contiguous or (stacked and rawOut.shape[0] != n):
        raise ValueError("rawOut must be a C-contiguous float32 array (one output tensor per input tensor)")
    if filters.Filters is None or filters.Filters.size == 0:
        raise ValueError("FilterSet has no filters (call ToTensor first)")
    w = np.ascontiguousarray(filters.Filters, dtype=np.float64)
    shp = (C.c_int32 * 4)(*([int(d) for d in shape] + [0] * (4 - len(shape)))[:4])
    _lib.check(_lib.lib().aud_gabor_convolve(device, mel.ctypes.data, n, mel.shape[-2], mel.shape[-1], w.ctypes.data,
                                             w.shape[0], filters.SizeX, filters.SizeY, filters.StrideX, filters.StrideY,
                                             float(filters.Gain), len(shape), shp, int(bool(byTime)), rawOut.ctypes.data))
