"""dft.Params mirror (reference dft/dft.go:15-39).  The arithmetic of
dft.Filter / dft.Power (dft.go:42-85) runs inside the fused CUDA kernel."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib


@dataclass
class Params:
    CompLogPow: bool = True
    LogMin: float = -100.0
    LogOffSet: float = 1.0
    PrevSmooth: float = 0.0
    CurSmooth: float = 1.0

    def Defaults(self) -> None:
        """dft/dft.go:33-39."""
        self.PrevSmooth = 0.0
        self.CurSmooth = 1.0 - self.PrevSmooth
        self.CompLogPow = True
        self.LogOffSet = 1.0
        self.LogMin = -100.0

    def FilterSegment(self, windows: np.ndarray, device: int = 0, log_power: bool = True):
        """dft.Params.Filter (dft/dft.go:42-85) for every step of one segment: windows[steps][winSamples] are the
        frames a per-step caller cut out itself (examples/gaborview/gbv.go:627-634).  Returns (PowerSegment,
        LogPowerSegment) shaped [winSamples/2+1][steps]; the second is None when CompLogPow is off or not wanted."""
        w = np.ascontiguousarray(windows, dtype=np.float32)
        if w.ndim != 2:
            raise ValueError("windows must be [steps][winSamples]")
        steps, n = w.shape
        dp = _lib.AudDftParams(int(self.CompLogPow), self.LogMin, self.LogOffSet, self.PrevSmooth, self.CurSmooth)
        power = np.zeros((n // 2 + 1, steps), dtype=np.float32)
        logp = np.zeros_like(power) if (log_power and self.CompLogPow) else None
        _lib.check(_lib.lib().aud_dft_filter(device, C.byref(dp), w.ctypes.data, steps, n, power.ctypes.data,
                                             None if logp is None else logp.ctypes.data))
        return power, logp
