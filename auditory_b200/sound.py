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
import zlib
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _lib, agabor, dft, kwta, mel
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


class Wave:
    """sound.Wave (sound/sound.go:32-141): a decoded WAV file.  Load stands in for go-audio/wav's
    Decoder.FullPCMBuffer (third-party, unpinned): integer PCM, little endian, 8 (unsigned, as go-audio
    reads it) / 16 / 24 / 32 bits, channels interleaved in `Data`."""

    def __init__(self):
        self.Data = np.zeros(0, dtype=np.int32)   # audio.IntBuffer.Data
        self.SourceBitDepth = 16
        self.NumChannels = 1
        self.Rate = 0

    def Load(self, fn: str) -> None:
        with open(fn, "rb") as f:
            raw = f.read()
        if len(raw) < 12 or raw[:4] != b"RIFF" or raw[8:12] != b"WAVE":
            raise ValueError(f"sound.Load: {fn} is not a RIFF/WAVE file")
        pos, fmt, data = 12, None, None
        while pos + 8 <= len(raw):
            cid, size = raw[pos:pos + 4], int.from_bytes(raw[pos + 4:pos + 8], "little")
            body = raw[pos + 8:pos + 8 + size]
            if cid == b"fmt ":
                fmt = body
            elif cid == b"data":
                data = body
                break
            pos += 8 + size + (size & 1)
        if fmt is None or data is None or len(fmt) < 16:
            raise ValueError(f"sound.Load: {fn} has no fmt / data chunk")
        tag = int.from_bytes(fmt[0:2], "little")
        self.NumChannels = int.from_bytes(fmt[2:4], "little")
        self.Rate = int.from_bytes(fmt[4:8], "little")
        self.SourceBitDepth = int.from_bytes(fmt[14:16], "little")
        if tag not in (1, 0xFFFE):
            raise ValueError(f"sound.Load: only integer PCM is supported (format tag {tag})")
        b = np.frombuffer(data, dtype=np.uint8)
        bd = self.SourceBitDepth
        if bd == 8:
            self.Data = b.astype(np.int32)
        elif bd == 16:
            self.Data = np.frombuffer(data[:len(data) // 2 * 2], dtype="<i2").astype(np.int32)
        elif bd == 24:
            t = b[:len(b) // 3 * 3].reshape(-1, 3).astype(np.int32)
            v = t[:, 0] | (t[:, 1] << 8) | (t[:, 2] << 16)
            self.Data = np.where(v & 0x800000, v - (1 << 24), v).astype(np.int32)
        elif bd == 32:
            self.Data = np.frombuffer(data[:len(data) // 4 * 4], dtype="<i4").astype(np.int32)
        else:
            raise ValueError(f"sound.Load: unsupported bit depth {bd}")

    def SampleRate(self) -> int:
        return self.Rate

    def Channels(self) -> int:
        return self.NumChannels

    def NumFrames(self) -> int:
        return self.Data.size // max(1, self.NumChannels)

    def GetFloatAtIdx(self, idx: int) -> float:
        """sound/sound.go:130-141."""
        scale = {32: 0x7FFFFFFF, 24: 0x7FFFFF, 16: 0x7FFF, 8: 0x7F}.get(self.SourceBitDepth)
        return float(self.Data[idx]) / float(scale) if scale else 0.0

    def SoundToTensor(self) -> np.ndarray:
        """sound/sound.go:116-127: Data[i] for i < NumFrames -- for interleaved multi-channel data that is
        the first NumFrames interleaved samples, as in the reference."""
        scale = {32: 0x7FFFFFFF, 24: 0x7FFFFF, 16: 0x7FFF, 8: 0x7F}.get(self.SourceBitDepth)
        n = self.NumFrames()
        if not scale:
            return np.zeros(n, dtype=np.float64)
        return self.Data[:n].astype(np.float64) / float(scale)

    def pcm16(self) -> Optional[np.ndarray]:
        """The int16 samples SoundToTensor would normalise, when the file is 16-bit: they can go to the GPU
        as they are (aud_process_host_i16 applies the /0x7FFF there)."""
        if self.SourceBitDepth != 16:
            return None
        return self.Data[:self.NumFrames()].astype(np.int16)


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
        self.Kwta = kwta.KWTA()
        self.NeighInhib = kwta.NeighInhib()
        self.KwtaPool = False
        self.Sound = Wave()
        self.SampleRate = 0          # stands in for Sound.SampleRate()
        self.Channels = 1            # stands in for Sound.Channels()
        self._signal = np.zeros(0, dtype=np.float32)
        self.SegCnt = 0
        self.device = device
        self._pipe: Optional[Pipeline] = None
        self._pipe_key = None
        self._cache: Optional[Dict[str, np.ndarray]] = None
        self._cache_key = None
        self._gabor_shape = None

    # The ProcessSegment cache (one batched GPU call per signal) must never outlive the signal it was computed from:
    # assigning Signal drops it, and the key carries a checksum of the samples for edits made in place.
    @property
    def Signal(self) -> np.ndarray:
        return self._signal

    @Signal.setter
    def Signal(self, value) -> None:
        self._signal = value
        self._cache = None

    def Invalidate(self) -> None:
        """Drop cached features (after editing Signal in place, for example)."""
        self._cache = None

    def _signal_key(self):
        sig = self._signal
        n = sig.size
        # every sample for short signals, an even spread of 64 K samples plus both ends for long ones
        probe = sig if n <= (1 << 16) else np.concatenate([sig[:: max(1, n >> 16)], sig[:256], sig[-256:]])
        return (id(sig), sig.ctypes.data, n, zlib.crc32(np.ascontiguousarray(probe).view(np.uint8)))

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
        self.Kwta.Defaults()                 # kwta ON and pool mode, as in the reference (sndenv.go:189-190)
        self.KwtaPool = True
        self.ByTime = False

    def SetSignal(self, samples: np.ndarray, sample_rate: int) -> None:
        """Stands in for Sound.Load + ToTensor (sndenv.go:297-300): mono float samples."""
        self.Signal = np.ascontiguousarray(samples, dtype=np.float32).reshape(-1)
        self.SampleRate = int(sample_rate)
        self._cache = None

    def ToTensor(self) -> bool:
        """sound/sndenv.go:297-300: Sound -> Signal (normalised samples)."""
        self.Signal = self.Sound.SoundToTensor().astype(np.float32)
        self.SampleRate = self.Sound.SampleRate()
        self.Channels = self.Sound.Channels()
        self._cache = None
        return True

    def AdjustForSilence(self, add: float, existing: float) -> int:
        """sound/sndenv.go:274-294: trim or prepend leading silence (milliseconds); returns the offset."""
        sr = self.SampleRate
        if sr <= 0:
            return -1
        offset = 0
        if add >= 0:
            if add < existing:
                offset = int(existing - add)
                self.Signal = self.Signal[MSecToSamples(float(offset), sr):]
            elif add > existing:
                offset = int(add - existing)
                self.Signal = np.concatenate([np.zeros(MSecToSamples(float(offset), sr), dtype=self.Signal.dtype), self.Signal])
        self._cache = None
        return offset

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
        key = (int(add), bool(power)) + self._signal_key()
        if self._cache is None or self._cache_key != key:
            sig = np.ascontiguousarray(self._signal, dtype=np.float32).reshape(-1)
            self._cache = pipe.process_host(sig, [0], [sig.size], want=self._wanted(power),
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
        """sound/sndenv.go:481-497: GborOutput of the segment last processed, then ApplyNeighInhib / ApplyKwta
        (:303-323).  Returns GborKwta when Kwta.On, else GborOutput.  The kwta step runs once for all segments of the
        signal, in segment order (KWTAPool keeps per-pool state from call to call, as se.Inhibs does)."""
        if not self._n_gabor:
            return self.GborOutput
        self.GborOutput = self._cache["gabor"][self._segment].reshape(self._gabor_shape)
        self.ExtGi = np.zeros(self._gabor_shape, dtype=np.float32)
        if not (self.Kwta.On or self.NeighInhib.On):
            return self.GborOutput
        key = repr((self.Kwta, self.NeighInhib, self.KwtaPool))
        if self._cache.get("kwta_key") != key:
            self._cache["kwta"], self._cache["ext_gi"] = kwta.Apply(self.Kwta, self.NeighInhib, self.KwtaPool,
                                                                    self._cache["gabor"], self._gabor_shape, device=self.device)
            self._cache["kwta_key"] = key
        self.ExtGi = self._cache["ext_gi"][self._segment].reshape(self._gabor_shape)
        if not self.Kwta.On:
            return self.GborOutput
        self.GborKwta = self._cache["kwta"][self._segment].reshape(self._gabor_shape)
        return self.GborKwta

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
