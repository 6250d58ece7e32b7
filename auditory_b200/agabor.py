"""agabor.Filter / agabor.FilterSet mirror (reference agabor/gabor.go:17-70).
ToTensor is the C-ABI's aud_gabor_to_tensor; Convolve (gabor.go:225-315) runs
inside the fused CUDA kernel."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

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
