"""sound.SndEnv mirror (reference sound/sndenv.go) over the B200 C-ABI.

Same field and method names as the Go type for the speech-feature path
(Params, DFT, Mel, GaborSpecs, GaborFilters, GborOut*, Defaults, Init,
ProcessSegment, ApplyGabor, Tail, Pad, MSecToSamples).  Tensors are float32
numpy arrays (the drop-in narrows etensor.Float64 to Float32).

ProcessSegment keeps the reference's call shape -- one segment per call, results
left in MelFBankSegment / MFCCSegment / Energy / ... -- but the GPU work is
batched: the first call for a given `add` runs every segment of the signal
through the fused kernel in one launch and later calls are served from that
result.  ProcessBatch is the batched entry point for many utterances.

Kwta / NeighInhib (sndenv.go:303-323) are outside this path: with
both off, ApplyGabor returns GborOutput exactly as the reference does.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _lib, agabor, dft, mel
from .pipeline import Pipeline


def MSecToSamples(ms: float, rate: int) -> int:
    """sound/sndenv.go:522-524."""
    return int(_lib.lib().aud_msec_to_samples(float(ms), int(rate)))


def SamplesToMSec(samples: int, rate: int) -> float:
    """sound/sndenv.go:527-529."""
    return 1000.0 * float(samples) / float(rate)


def _go_round(x: float) -> float:
    return math.floor(x + 0.5) if x >= 0 else -math.floor(-x + 0.5)


@dataclass
class Params:
    """sound/sndenv.go:24-61."""
    WinMs: float = 25.0
    StepMs: float = 10.0
    SegmentMs: float = 100.0
    StrideMs: float = 100.0
    BorderSteps: int = 2
    Channel: int = 0
    WinSamples: int = 0
    StepSamples: int = 0
    SegmentSamples: int = 0
    StrideSamples: int = 0
    SegmentSteps: int = 0
    Steps: List[int] = field(default_factory=list)


class SndEnv:
    def __init__(self, device: int = 0):
        self.Nm = ""
        self.Dsc = ""
        self.On = True
        self.Params = Params()
        self.DFT = dft.Params()
        self.Mel = mel.Params()
        self.GaborSpecs: List[agabor.Filter] = []
        self.GaborFilters = agabor.FilterSet()
        self.GborOutPoolsX = 0
        self.GborOutPoolsY = 0
        self.GborOutUnitsX = 0
        self.GborOutUnitsY = 0
        self.ByTime = False
        self.SampleRate = 0          # stands in for Sound.SampleRate()
        self.Channels = 1            # stands in for Sound.Channels()
        self.Signal = np.zeros(0, dtype=np.float32)
        self.SegCnt = 0
        self.device = device
        self._pipe: Optional[Pipeline] = None
        self._pipe_key = None
        self._cache: Optional[Dict[str, np.ndarray]] = None
        self._cache_key = None
        self._gabor_shape = None

    # ---------------------------------------------------------------- set-up
    def ParamDefaults(self) -> None:
        """sound/sndenv.go:64-71."""
        self.Params.WinMs = 25.0
        self.Params.StepMs = 10.0
        self.Params.SegmentMs = 100.0
        self.Params.Channel = 0
        self.Params.StrideMs = 100.0
        self.Params.BorderSteps = 2

    def Defaults(self) -> None:
        """sound/sndenv.go:185-192."""
        self.ParamDefaults()
        self.On = True
        self.Mel.Defaults()
        self.ByTime = False

    def SetSignal(self, samples: np.ndarray, sample_rate: int) -> None:
        """Stands in for Sound.Load + ToTensor (sndenv.go:297-300): mono float samples."""
        self.Signal = np.ascontiguousarray(samples, dtype=np.float32).reshape(-1)
        self.SampleRate = int(sample_rate)
        self._cache = None

    def Init(self) -> None:
        """sound/sndenv.go:195-267."""
        sr = self.SampleRate
        if sr <= 0:
            raise ValueError("sample rate <= 0")
        p = self.Params
        p.WinSamples = MSecToSamples(p.WinMs, sr)
        p.StepSamples = MSecToSamples(p.StepMs, sr)
        p.SegmentSamples = MSecToSamples(p.SegmentMs, sr)
        steps = int(_go_round(p.SegmentMs / p.StepMs))
        p.SegmentSteps = steps + 2 * p.BorderSteps
        p.StrideSamples = MSecToSamples(p.StrideMs, sr)

        specs = agabor.Active(self.GaborSpecs)
        nfilters = len(specs)
        agabor.ToTensor(specs, self.GaborFilters)
        if self.GborOutPoolsX == 0 and self.GborOutPoolsY == 0:
            self._gabor_shape = (self.GborOutUnitsY, self.GborOutUnitsX)
        elif self.GborOutPoolsX > 0 and self.GborOutPoolsY > 0:
            self._gabor_shape = (self.GborOutPoolsY, self.GborOutPoolsX, self.GborOutUnitsY, self.GborOutUnitsX)
        else:
            raise ValueError("GborOutPoolsX & GborOutPoolsY must both be == 0 or > 0 (i.e. 2D or 4D)")
        self.GborOutput = np.zeros(self._gabor_shape, dtype=np.float32)
        self._n_gabor = nfilters

        half = p.WinSamples // 2 + 1
        self.DFT.Defaults()                                    # wipes smoothing set before Init
        self.MelFilters = self.Mel.InitFilters(p.WinSamples, sr)
        S = p.SegmentSteps
        self.PowerSegment = np.zeros((half, S), dtype=np.float32)
        self.LogPowerSegment = np.zeros((half, S), dtype=np.float32)
        p.Steps = [p.StepSamples * (i - p.BorderSteps) for i in range(S)]
        nf = self.Mel.FBank.NFilters
        self.MelFBankSegment = np.zeros((nf, S), dtype=np.float32)
        self.Energy = np.zeros(S, dtype=np.float32)
        if self.Mel.MFCC:
            self.MFCCSegment = np.zeros((self.Mel.NCoefs, S), dtype=np.float32)
            self.MFCCDeltas = np.zeros((self.Mel.NCoefs, S), dtype=np.float32)
            self.MFCCDeltaDeltas = np.zeros((self.Mel.NCoefs, S), dtype=np.float32)
        siglen = len(self.Signal) - p.SegmentSamples * self.Channels
        siglen = int(siglen / self.Channels)                  # Go integer division truncates toward zero
        self.SegCnt = int(siglen / p.StrideSamples) + 1
        self._cache = None
        self._pipe_key = None

    # ------------------------------------------------------------ GPU plumbing
    def aud_params(self) -> _lib.AudParams:
        p, fb = self.Params, self.Mel.FBank
        ap = _lib.AudParams()
        ap.sample_rate = self.SampleRate
        ap.win_samples, ap.step_samples = p.WinSamples, p.StepSamples
        ap.segment_samples, ap.stride_samples = p.SegmentSamples, p.StrideSamples
        ap.segment_steps, ap.border_steps = p.SegmentSteps, p.BorderSteps
        ap.comp_log_pow = int(self.DFT.CompLogPow)
        ap.log_min, ap.log_offset = self.DFT.LogMin, self.DFT.LogOffSet
        ap.prev_smooth, ap.cur_smooth = self.DFT.PrevSmooth, self.DFT.CurSmooth
        ap.n_mel = fb.NFilters
        ap.mel_log_off, ap.mel_log_min = fb.LogOff, fb.LogMin
        ap.renorm = int(fb.Renorm)
        ap.renorm_min, ap.renorm_scale = fb.RenormMin, fb.RenormScale
        ap.mfcc, ap.n_coefs, ap.deltas = int(self.Mel.MFCC), self.Mel.NCoefs, int(self.Mel.MFCC and self.Mel.Deltas)
        ap.mfcc_c0_energy = 1
        ap.gabor_nf = self._n_gabor
        gf = self.GaborFilters
        ap.gabor_size_x, ap.gabor_size_y = gf.SizeX, gf.SizeY
        ap.gabor_stride_x, ap.gabor_stride_y = gf.StrideX, gf.StrideY
        ap.gabor_gain = gf.Gain
        ap.gabor_out_dims = len(self._gabor_shape)
        for i, d in enumerate(self._gabor_shape):
            ap.gabor_shape[i] = d
        ap.gabor_by_time = int(self.ByTime)
        return ap

    def pipeline(self) -> Pipeline:
        ap = self.aud_params()
        key = bytes(ap) + self.MelFilters.tobytes() + self.Mel.BinPts.tobytes() + \
            (self.GaborFilters.Filters.tobytes() if self._n_gabor else b"")
        if self._pipe is None or key != self._pipe_key:
            if self._pipe is not None:
                self._pipe.close()
            self._pipe = Pipeline(ap, self.Mel.BinPts, self.MelFilters,
                                  self.GaborFilters.Filters if self._n_gabor else None, device=self.device)
            self._pipe_key = key
            self._cache = None
        return self._pipe

    def _wanted(self, power: bool = False) -> List[str]:
        want = ["mel", "energy"]
        if self.Mel.MFCC:
            want.append("mfcc")
            if self.Mel.Deltas:
                want += ["deltas", "delta_deltas"]
        if self._n_gabor:
            want.append("gabor")
        if power:
            want.append("power")
            if self.DFT.CompLogPow:
                want.append("logpower")
        return want

    def ProcessBatch(self, wave: np.ndarray, utt_offset: Sequence[int], utt_len: Sequence[int], add: int = 0,
                     want: Optional[Sequence[str]] = None) -> Dict[str, np.ndarray]:
        """Every segment of every utterance in one fused launch.  Outputs are
        [total_segments, ...] float32 in the reference's per-segment layouts."""
        pipe = self.pipeline()
        return pipe.process_host(np.ascontiguousarray(wave, dtype=np.float32), utt_offset, utt_len,
                                 want=self._wanted() if want is None else want,
                                 add_samples=MSecToSamples(float(add), self.SampleRate))

    # ----------------------------------------------------- reference call shape
    def ProcessSegment(self, segment: int, add: int = 0, power: bool = True) -> None:
        """sound/sndenv.go:342-433 for one segment (see module docstring)."""
        pipe = self.pipeline()
        key = (int(add), bool(power), self.Signal.ctypes.data, self.Signal.size)
        if self._cache is None or self._cache_key != key:
            self._cache = pipe.process_host(self.Signal, [0], [self.Signal.size], want=self._wanted(power),
                                            add_samples=MSecToSamples(float(add), self.SampleRate))
            self._cache_key = key
        c = self._cache
        nseg = c["mel"].shape[0]
        if not 0 <= segment < nseg:
            raise IndexError(f"segment {segment} outside 0..{nseg - 1}")
        self.MelFBankSegment = c["mel"][segment]
        self.Energy = c["energy"][segment]
        if "mfcc" in c:
            self.MFCCSegment = c["mfcc"][segment]
        if "deltas" in c:
            self.MFCCDeltas = c["deltas"][segment]
            self.MFCCDeltaDeltas = c["delta_deltas"][segment]
        if "power" in c:
            self.PowerSegment = c["power"][segment]
        if "logpower" in c:
            self.LogPowerSegment = c["logpower"][segment]
        self._segment = segment

    def ApplyGabor(self) -> np.ndarray:
        """sound/sndenv.go:481-497 with Kwta.On = NeighInhib.On = false."""
        if not self._n_gabor:
            return self.GborOutput
        self.GborOutput = self._cache["gabor"][self._segment].reshape(self._gabor_shape)
        return self.GborOutput

    # ------------------------------------------------------------ host helpers
    def Tail(self, signal: np.ndarray) -> int:
        """sound/sndenv.go:503-507 (Go % truncates toward zero)."""
        temp = len(signal) - self.Params.SegmentSamples
        return int(math.fmod(temp, self.Params.StrideSamples))

    def Pad(self, signal: np.ndarray, value: float = 0.0) -> np.ndarray:
        """sound/sndenv.go:510-519."""
        tail = self.Tail(signal)
        pad_len = self.Params.SegmentSamples - self.Params.StepSamples - int(math.fmod(tail, self.Params.StepSamples))
        return np.concatenate([np.asarray(signal), np.full(pad_len, value, dtype=np.asarray(signal).dtype)])

    def Name(self) -> str:
        return self.Nm

    def Desc(self) -> str:
        return self.Dsc
