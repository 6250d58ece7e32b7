"""mel.FilterBank / mel.Params mirror (reference mel/mel.go:16-180).  Table
construction is the C-ABI's aud_mel_init_filters; FilterDft / CepstrumDct
(mel.go:120-153, 192-212) run inside the fused CUDA kernel."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import _lib


def FreqToMel(freq: float) -> float:
    return float(_lib.lib().aud_freq_to_mel(freq))


def MelToFreq(mel: float) -> float:
    return float(_lib.lib().aud_mel_to_freq(mel))


def FreqToBin(freq: float, n_fft: float, sample_rate: float) -> int:
    return int(_lib.lib().aud_freq_to_bin(freq, n_fft, sample_rate))


@dataclass
class FilterBank:
    NFilters: int = 32
    LoHz: float = 0.0
    HiHz: float = 8000.0
    LogOff: float = 0.0
    LogMin: float = -10.0
    Renorm: bool = True
    RenormMin: float = -6.0
    RenormMax: float = 4.0
    RenormScale: float = 0.0

    def Defaults(self) -> None:
        """mel/mel.go:171-180."""
        self.LoHz = 0.0
        self.HiHz = 8000.0
        self.NFilters = 32
        self.LogOff = 0.0
        self.LogMin = -10.0
        self.Renorm = True
        self.RenormMin = -6.0
        self.RenormMax = 4.0


@dataclass
class Params:
    FBank: FilterBank = field(default_factory=FilterBank)
    BinPts: Optional[np.ndarray] = None
    HzPts: Optional[np.ndarray] = None
    MFCC: bool = False
    Deltas: bool = False
    NCoefs: int = 13

    def Defaults(self) -> None:
        """mel/mel.go:69-74: MFCC and Deltas are switched ON."""
        self.FBank.Defaults()
        self.MFCC = True
        self.NCoefs = 13
        self.Deltas = True

    def InitFilters(self, dft_size: int, sample_rate: int) -> np.ndarray:
        """mel/mel.go:77-117; returns the [NFilters, NFilters+2] float64 table
        (the `filters` tensor of the reference signature)."""
        nf = self.FBank.NFilters
        self.BinPts = np.zeros(nf + 2, dtype=np.int32)
        self.HzPts = np.zeros(nf + 2, dtype=np.float64)
        self.FBank.Renorm = False           # mel.go:80
        filters = np.zeros((nf, nf + 2), dtype=np.float64)
        _lib.check(_lib.lib().aud_mel_init_filters(dft_size, sample_rate, nf, self.FBank.LoHz, self.FBank.HiHz,
                                                   self.BinPts.ctypes.data, self.HzPts.ctypes.data,
                                                   filters.ctypes.data))
        return filters
