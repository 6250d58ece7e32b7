"""dft.Params mirror (reference dft/dft.go:15-39).  The arithmetic of
dft.Filter / dft.Power (dft.go:42-85) runs inside the fused CUDA kernel."""
from __future__ import annotations

from dataclasses import dataclass


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
