"""mel.FilterBank / mel.Params mirror (reference mel/mel.go:16-180).  Table
construction is the C-ABI's aud_mel_init_filters; FilterDft / CepstrumDct
(mel.go:120-153, 192-212) run inside the fused CUDA kernel."""
from __future__ import annotations

import ctypes as C
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

    def FilterDftSegment(self, power_segment: np.ndarray, filters: np.ndarray, device: int = 0) -> np.ndarray:
        """mel.Params.FilterDft (mel/mel.go:120-153) for every step of one segment: PowerSegment [bins][steps] ->
        MelFBankSegment [NFilters][steps]."""
        pw = np.ascontiguousarray(power_segment, dtype=np.float32)
        fb = self.FBank
        mp = _lib.AudMelParams(fb.NFilters, fb.LogOff, fb.LogMin, int(fb.Renorm), fb.RenormMin, fb.RenormScale)
        bp = np.ascontiguousarray(self.BinPts, dtype=np.int32)
        ft = np.ascontiguousarray(filters, dtype=np.float64)
        out = np.zeros((fb.NFilters, pw.shape[1]), dtype=np.float32)
        _lib.check(_lib.lib().aud_mel_filter_dft(device, C.byref(mp), bp.ctypes.data, ft.ctypes.data, pw.ctypes.data,
                                                 pw.shape[0], pw.shape[1], out.ctypes.data))
        return out

    def CepstrumDctSegment(self, mel_segment: np.ndarray, device: int = 0) -> np.ndarray:
        """mel.Params.CepstrumDct (mel/mel.go:192-212) for every step: MelFBankSegment [NFilters][steps] ->
        MFCCSegment [NCoefs][steps] with coefficient 0 = ln(1 + y0^2)."""
        ms = np.ascontiguousarray(mel_segment, dtype=np.float32)
        out = np.zeros((self.NCoefs, ms.shape[1]), dtype=np.float32)
        _lib.check(_lib.lib().aud_cepstrum_dct(device, ms.ctypes.data, ms.shape[0], ms.shape[1], self.NCoefs, None,
                                               out.ctypes.data))
        return out
